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
fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
fe.kernel_h2d = sys.argv[1][0] == "k"; fe.kernel_d2h = sys.argv[1][1] == "k"
for _ in range(3): fe.extract_host(wp, n, device=dev)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fe.extract_host(wp, n, device=dev); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == "CUDA"]
ev.sort(key=lambda e: e.time_range.start)
t00 = ev[0].time_range.start
agg = {}
for e in ev:
    nm_ = e.name[:40]
    d = e.time_range.end - e.time_range.start
    if d > 30 or "ragged" in nm_:
        print("%9.1f us +%8.1f us  %s" % (e.time_range.start - t00, d, nm_))
print("end", max(e.time_range.end for e in ev) - t00, "n events", len(ev))
