"""Error statistics of the GPU kernel, torchaudio fp32 and the numpy fp32 oracle against the fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from torchaudio.compliance import kaldi
from oracle import lasr_frontend

def viol(a, b): return int((np.abs(a - b) > 1e-5 + 1e-4 * np.abs(b)).sum())
rng = np.random.default_rng(0)
wavs = [rng.uniform(-0.5, 0.5, 160000) for _ in range(16)]
fe = lasr_b200.GpuFbankFrontend()
w = np.stack(wavs).astype(np.float32)
f, _ = fe(torch.from_numpy(w).cuda(), np.full(16, 160000))
g = f.cpu().numpy()
tot = dict(g_ta=0, g_np=0, g_64=0, ta_64=0, np_64=0, ta_np=0)
eg = et = en = 0.0
for i in range(16):
    ta = kaldi.fbank(torch.from_numpy(w[i:i+1]) * 32768.0, num_mel_bins=80, dither=0.0, energy_floor=1.0).numpy()
    n32 = lasr_frontend.wav_to_kaldi_fbank(wavs[i])
    r64 = lasr_frontend.wav_to_kaldi_fbank(wavs[i], dtype=np.float64)
    tot["g_ta"] += viol(g[i], ta); tot["g_np"] += viol(g[i], n32); tot["g_64"] += viol(g[i], r64)
    tot["ta_64"] += viol(ta, r64); tot["np_64"] += viol(n32, r64); tot["ta_np"] += viol(ta, n32)
    eg = max(eg, np.abs(g[i] - r64).max()); et = max(et, np.abs(ta - r64).max()); en = max(en, np.abs(n32 - r64).max())
    if i == 0:
        for name, a in (("gpu", g[i]), ("ta32", ta), ("np32", n32)):
            d = np.abs(a - r64)
            print(name, "rms err %.3e  p99.99 %.3e  max %.3e  max-in-bins>=10 %.3e" % (np.sqrt((d**2).mean()), np.quantile(d, 0.9999), d.max(), d[:, 10:].max()))
print("violations (of %d):" % (16 * 998 * 80), tot)
print("max abs err vs fp64: gpu %.3e  torchaudio %.3e  numpy32 %.3e" % (eg, et, en))
