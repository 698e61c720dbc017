"""Where a B200Collate call spends its wall time (host stamps of HostPipeline): float64 and int16 C2 lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from bench import make_list

lists = [make_list(s)[0] for s in (1, 101)]
col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=True, cmvn="utt_meanvar")
import time
for kind in ("float64", "int16", "int16 dma", "int16 to_host=False", "int16 bf16"):
    ls = lists if kind == "float64" else [[np.round(w * 32767).astype(np.int16) for w in l] for l in lists]
    col.to_host = "False" not in kind
    col.pipeline.d2h_mode = "dma" if "dma" in kind else "kernel"
    if "bf16" in kind:
        col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=True, cmvn="utt_meanvar", out_dtype=torch.bfloat16)
    for i in range(4):
        col(ls[i % 2])
    t0 = time.perf_counter()
    for i in range(10):
        col(ls[i % 2])
    print("== %s: %.3f ms per synchronous call" % (kind, (time.perf_counter() - t0) * 100))
    for i in range(4):
        col(ls[i % 2])
    col.pipeline.trace = []
    col.pipeline.dev_trace = []
    col(ls[0])
    torch.cuda.synchronize()
    dtr = col.pipeline.dev_trace
    print("== %s: device timeline of one call (ms after the call's first stream operation)" % kind)
    for lab, ev in dtr[1:]:
        print("  %-18s %8.3f" % (lab, dtr[0][1].elapsed_time(ev)))
    col.pipeline.trace = []
    col.pipeline.dev_trace = []
    col(ls[0]); col(ls[1])
    tr = col.pipeline.trace
    col.pipeline.trace = None
    t0 = tr[0][1]
    print("==", kind)
    prev = t0
    for lab, t in tr:
        print("  %-12s %8.3f ms  (+%.3f)" % (lab, (t - t0) * 1e3, (t - prev) * 1e3))
        prev = t
