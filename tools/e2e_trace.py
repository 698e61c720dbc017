"""Ad-hoc: where does extract_host spend its time? CPU enqueue time vs device time, per variant."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = "cuda:0"
rng = np.random.default_rng(1)
n = np.round(rng.uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
nm = int((n.max() + 3) // 4 * 4)
w = np.zeros((256, nm), dtype=np.float32)
for i in range(256):
    w[i, : n[i]] = np.clip(rng.normal(0, 0.1, n[i]), -1, 1)
wp = torch.from_numpy(w).pin_memory()
for kc in (True, False):
    for cm in ("utt_meanvar", "none"):
        fe = lasr_b200.GpuFbankFrontend(cmvn=cm)
        fe.kernel_copies = kc
        for rh in (True, False):
            for _ in range(3): fe.extract_host(wp, n, device=dev, return_host=rh)
            torch.cuda.synchronize()
            K = 8
            cpu = 0.0
            t0 = time.perf_counter()
            for _ in range(K):
                c0 = time.perf_counter()
                fe.extract_host(wp, n, device=dev, return_host=rh)
                cpu += time.perf_counter() - c0
            torch.cuda.synchronize()
            tot = time.perf_counter() - t0
            print("kernel_copies=%s cmvn=%s return_host=%s: wall %.3f ms/step, cpu enqueue %.3f ms/step" % (kc, cm, rh, tot / K * 1e3, cpu / K * 1e3), flush=True)
# single-call latency with a sync before (no cross-step pipelining)
fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
for _ in range(3): fe.extract_host(wp, n, device=dev)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fe.extract_host(wp, n, device=dev); torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("single call with syncs: %s ms" % ["%.3f" % t for t in ts])
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fe.extract_host(wp, n, device=dev); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
t00 = ev[0].time_range.start
for e in ev[:60]:
    print("%9.1f us +%8.1f us  %s" % (e.time_range.start - t00, e.time_range.end - e.time_range.start, e.name[:60]))
