"""Ad-hoc: first GPU numbers (parity summary + kernel time) for the fused fbank launch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from torchaudio.compliance import kaldi

dev = "cuda:0"
rng = np.random.default_rng(0)
B, N = 16, 160000
w = rng.uniform(-0.5, 0.5, (B, N)).astype(np.float32)
wav = torch.from_numpy(w).to(dev)
n = np.full(B, N, dtype=np.int64)
fe = lasr_b200.GpuFbankFrontend()
f, fl = fe(wav, n)
torch.cuda.synchronize()
ref = kaldi.fbank(torch.from_numpy(w[0:1]) * 32768.0, num_mel_bins=80, dither=0.0, energy_floor=1.0).numpy()
g = f[0].cpu().numpy()
d = np.abs(g - ref)
print("shape", tuple(f.shape), "max abs diff", d.max(), "violations", int((d > 1e-5 + 1e-4 * np.abs(ref)).sum()), "nan", int(np.isnan(g).sum()))
print("gpu[0,:5]", g[0, :5], "ref", ref[0, :5])
# timing on a big batch (C2-like total size)
B2, N2 = 256, 16000 * 18
wav2 = (torch.randn((B2, N2), device=dev) * 0.1).clamp_(-1, 1)
n2 = np.full(B2, N2, dtype=np.int64)
for _ in range(3):
    fe(wav2, n2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 10
for _ in range(K):
    fe(wav2, n2)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
hours = B2 * N2 / 16000 / 3600
frames = B2 * (1 + (N2 - 400) // 160)
print("ms/step %.3f  audio-h/s %.1f  GB/s(alg) %.1f  ns/frame %.3f" % (ms, hours / (ms * 1e-3), (4 * B2 * N2 + 320 * frames) / (ms * 1e-3) / 1e9, ms * 1e6 / frames))

# raw C-ABI launch timing (no Python allocation in the loop)
import ctypes as C
from importlib import import_module
L = import_module("lighting-asr_b200._lib")
plan = fe.plan(dev)
T2 = int(1 + (N2 - 400) // 160)
out = torch.empty((B2, T2, 80), device=dev)
nd = torch.from_numpy(n2).to(dev)
a = L.FbankArgs()
a.d_wav = wav2.data_ptr(); a.wav_stride = wav2.stride(0); a.d_nsamp = nd.data_ptr(); a.batch = B2
a.d_out = out.data_ptr(); a.max_frames = T2
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3): L.check(plan.lib.b200fe_fbank_fused(plan.handle, C.byref(a), st), "x")
torch.cuda.synchronize()
e0.record()
for _ in range(20): plan.lib.b200fe_fbank_fused(plan.handle, C.byref(a), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("RAW kernel: ms %.4f  audio-h/s %.1f  GB/s(alg) %.1f  ns/frame %.3f  static_mel=%d" % (ms, hours / (ms * 1e-3), (4 * B2 * N2 + 320 * frames) / (ms * 1e-3) / 1e9, ms * 1e6 / frames, plan.lib.b200fe_plan_info(plan.handle, 0)))
