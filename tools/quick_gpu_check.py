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
