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
fe2.inlaunch_cmvn = False          # finalize + post-pass launches
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
fe2.profile_events = []
t_post = timeit(lambda: fe2(wav3, n3, out=out3))
torch.cuda.synchronize()
ev = fe2.profile_events[-30:]
print("%-8s C2 utt_meanvar post-pass path %.4f ms, of which the fused launch (statistics + features) %.4f ms" % (tag, t_post, sum(a.elapsed_time(b) for a, b in ev) / len(ev)), flush=True)
fe2.profile_events = None
for lag in (int(x) for x in os.environ.get("APPLY_LAGS", "16").split(",")):
    fe3 = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
    fe3.inlaunch_cmvn = True
    fe3.apply_lag = lag
    fe3.pad_tiles = fe.pad_tiles
    fe3.profile_events = []
    t = timeit(lambda: fe3(wav3, n3, out=out3))
    torch.cuda.synchronize()
    ev = fe3.profile_events[-30:]
    print("%-8s   fused launch with apply tiles alone %.4f ms" % (tag, sum(a.elapsed_time(b) for a, b in ev) / len(ev)), flush=True)
    work, idx = fe3.last["apply_flags"]
    print("%-8s C2 utt_meanvar in-launch (apply tiles, lag %d) %.4f ms  error flag %d" % (tag, lag, t, int(work[idx].item())), flush=True)
print("%-8s uniform plain %.4f | C2 plain %.4f | C2 utt_meanvar %.4f | C2 stats only %.4f  (ms)" % (tag, *r), flush=True)

# the work-list builders alone
import ctypes as C
plan = fe.plan(torch.device(dev)); lib = plan.lib
nd = torch.from_numpy(n3).to(dev)
Tm = int(T3.max())
cap = lib.b200fe_tile_table_capacity(plan.handle, 256, Tm, 1) + 256 * ((Tm + 239) // 240)
work = torch.empty((2 * cap + 2 + 257,), dtype=torch.int32, device=dev)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p0 = work.data_ptr()
t_old = timeit(lambda: lib.b200fe_build_tile_table_device(plan.handle, nd.data_ptr(), 256, Tm, 1, p0, cap, p0 + 8 * cap, p0 + 8 * cap + 4, st))
z = torch.empty((256, 2, 80), dtype=torch.float64, device=dev)
t_new = timeit(lambda: lib.b200fe_build_work_list_device(plan.handle, nd.data_ptr(), 256, Tm, 1, 0, p0, cap, p0 + 8 * cap, p0 + 8 * cap + 4, None, z.data_ptr(), z.numel() * 8, st))
t_z = timeit(lambda: torch.zeros((256, 2, 80), dtype=torch.float64, device=dev))
print("%-8s builders alone: frame+padding list %.4f ms, the same + statistics zero fill %.4f ms; torch.zeros of the statistics %.4f ms" % (tag, t_old, t_new, t_z), flush=True)
