"""Second ablation of the B200Collate call (int16 lists): compute removed, copy-kernel CTA count, D2H by DMA."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from bench import make_list
lists = [make_list(s)[0] for s in (1, 101, 201, 301)]
l16 = [[np.round(w * 32767).astype(np.int16) for w in l] for l in lists]
def bench_col(col, ls, n=12):
    for i in range(5):
        col(ls[i % 4])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        col(ls[i % 4])
    return (time.perf_counter() - t0) / n * 1e3
mode = sys.argv[1]
col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=True, cmvn="utt_meanvar")
if mode == "nocompute":
    fe = col.pipeline.fe
    fe.forward = lambda *a, **k: None
if mode == "dma":
    col.pipeline.d2h_mode = "dma"
print("%-12s copy CTAs %-4s %.3f ms per call" % (mode, os.environ.get("B200FE_COPY_CTAS", "16"), bench_col(col, l16)), flush=True)
if mode == "trace":
    col.pipeline.trace = []; col.pipeline.dev_trace = []
    col(l16[0]); torch.cuda.synchronize()
    d = col.pipeline.dev_trace
    print("  ".join("%s %.2f" % (lab, d[0][1].elapsed_time(ev)) for lab, ev in d[1:]))
