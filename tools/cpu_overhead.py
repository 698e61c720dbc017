"""Ad-hoc: CPU time of one GpuFbankFrontend.forward call (enqueue only) vs device time, C2 batch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = "cuda:0"
rng = np.random.default_rng(1)
n = np.round(rng.uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
nm = int((n.max() + 3) // 4 * 4)
wav = (torch.randn((256, nm), device=dev) * 0.1).clamp_(-1, 1)
T = 1 + (n - 400) // 160
out = torch.empty((256, int(T.max()), 80), device=dev)
for cm in ("none", "utt_meanvar"):
    fe = lasr_b200.GpuFbankFrontend(cmvn=cm)
    for _ in range(5): fe(wav, n, out=out)
    torch.cuda.synchronize()
    K = 50
    t0 = time.perf_counter()
    for _ in range(K): fe(wav, n, out=out)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("cmvn=%s: cpu enqueue %.3f ms/call, total %.3f ms/call" % (cm, (t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3))
import cProfile, pstats
fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
for _ in range(5): fe(wav, n, out=out)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): fe(wav, n, out=out)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(40)
