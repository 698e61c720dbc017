"""Ad-hoc A/B timing of the fused launch for one library build (B200FE_LIB): uniform batch and the C2 batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
tag = sys.argv[1]
dev = "cuda:0"
def timeit(fn, K=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
fe = lasr_b200.GpuFbankFrontend()
fe2 = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
if os.environ.get("PAD_TILES") == "0":
    fe.pad_tiles = fe2.pad_tiles = False
B2, N2 = 256, 16000 * 18
wav2 = (torch.randn((B2, N2), device=dev) * 0.1).clamp_(-1, 1)
n2 = np.full(B2, N2, dtype=np.int64)
T2 = 1 + (N2 - 400) // 160
out = torch.empty((B2, T2, 80), device=dev)
r = [timeit(lambda: fe(wav2, n2, out=out))]
rng = np.random.default_rng(1)
n3 = np.round(rng.uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
nm3 = int((n3.max() + 3) // 4 * 4)
wav3 = (torch.randn((256, nm3), device=dev) * 0.1).clamp_(-1, 1)
T3 = 1 + (n3 - 400) // 160
out3 = torch.empty((256, int(T3.max()), 80), device=dev)
r.append(timeit(lambda: fe(wav3, n3, out=out3)))
r.append(timeit(lambda: fe2(wav3, n3, out=out3)))
r.append(timeit(lambda: fe.accumulate_stats(wav3, n3)))
print("%-8s uniform plain %.4f | C2 plain %.4f | C2 utt_meanvar %.4f | C2 stats only %.4f  (ms)" % (tag, *r), flush=True)
