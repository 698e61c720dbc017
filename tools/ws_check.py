"""Ad-hoc A/B: parity dump + timings of the fused launch for one library build (B200FE_LIB selects the build, B200FE_WS=1 the experimental warp-specialised kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
tag = sys.argv[1]
dev = "cuda:0"
rng = np.random.default_rng(7)
B = 24
n = np.round(rng.uniform(0.3, 6.0, B) * 16000).astype(np.int64)
n[0] = 401; n[1] = 400 + 160 * 23; n[2] = 400 + 160 * 24; n[3] = 559
nmax = int((n.max() + 3) // 4 * 4)
w = np.zeros((B, nmax), dtype=np.float32)
for i in range(B):
    w[i, : n[i]] = rng.uniform(-0.5, 0.5, n[i])
wav = torch.from_numpy(w).to(dev)
res = {}
fe = lasr_b200.GpuFbankFrontend()
f, fl = fe(wav, n); res["plain"] = f.cpu().numpy(); res["len"] = fl.cpu().numpy()
fe2 = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
f, fl = fe2(wav, n); res["utt"] = f.cpu().numpy()
fe3 = lasr_b200.GpuFbankFrontend(peak_norm=True)
f, fl = fe3(wav, n); res["peak"] = f.cpu().numpy()
st = fe.accumulate_stats(wav, n); res["stats"] = st.cpu().numpy()
fe4 = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=res["stats"])
f, fl = fe4(wav, n); res["glob"] = f.cpu().numpy()
wi = torch.from_numpy(np.round(w * 32767).astype(np.int16)).to(dev)
f, fl = fe(wi, n); res["i16"] = f.cpu().numpy()
f, fl = fe(wav[:, 1:], n - 1); res["unaligned"] = f.cpu().numpy()       # generic (non-TMA) producer path
torch.cuda.synchronize()
np.savez("/tmp/ws_%s.npz" % tag, **res)
print(tag, "dump ok; nan:", {k: int(np.isnan(v).sum()) for k, v in res.items()})

def timeit(fn, K=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

B2, N2 = 256, 16000 * 18
wav2 = (torch.randn((B2, N2), device=dev) * 0.1).clamp_(-1, 1)
n2 = np.full(B2, N2, dtype=np.int64)
T2 = 1 + (N2 - 400) // 160
out = torch.empty((B2, T2, 80), device=dev)
frames = B2 * T2
ms = timeit(lambda: fe(wav2, n2, out=out))
print(tag, "uniform plain: ms %.4f ns/frame %.3f GB/s %.1f frac %.3f" % (ms, ms * 1e6 / frames, (4 * B2 * N2 + 320 * frames) / ms / 1e6, (4 * B2 * N2 + 320 * frames) / ms / 1e6 / 6544.3))
# C2-shaped ragged batch, utterance CMVN (statistics mode + post pass) and plain
rng = np.random.default_rng(1)
n3 = np.round(rng.uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
nm3 = int((n3.max() + 3) // 4 * 4)
wav3 = (torch.randn((256, nm3), device=dev) * 0.1).clamp_(-1, 1)
T3 = 1 + (n3 - 400) // 160
out3 = torch.empty((256, int(T3.max()), 80), device=dev)
alg = 4 * n3.sum() + 320 * T3.sum()
for name, f_ in (("C2 plain", fe), ("C2 utt_meanvar", fe2)):
    ms = timeit(lambda: f_(wav3, n3, out=out3))
    print(tag, "%s: ms %.4f audio-h/s %.1f GB/s(step) %.1f" % (name, ms, n3.sum() / 16000 / 3600 / (ms * 1e-3), alg / ms / 1e6))
ms = timeit(lambda: fe.accumulate_stats(wav3, n3))
print(tag, "C2 stats only: ms %.4f" % ms)
