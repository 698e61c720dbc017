"""B200Collate call time against the number of packing threads (16-core box: the calling thread and the CUDA driver's threads need
cores too), float64 and int16 C2 lists, 4 rotating batches; median and worst of 30 calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from bench import make_list
lists = [make_list(s)[0] for s in (1, 101, 201, 301)]
l16 = [[np.round(w * 32767).astype(np.int16) for w in l] for l in lists]
for kind, ls in (("float64", lists), ("int16", l16)):
    for nt in [int(x) for x in os.environ.get("THREADS", "10,12,13,14,15,16").split(",")]:
        col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=True, cmvn="utt_meanvar", threads=nt)
        for i in range(5):
            col(ls[i % 4])
        torch.cuda.synchronize()
        ts = []
        for i in range(30):
            t0 = time.perf_counter(); col(ls[i % 4]); ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        print("%-8s threads %2d: median %.2f ms, mean %.2f, p90 %.2f, max %.2f" % (kind, nt, ts[15], sum(ts) / 30, ts[27], ts[-1]), flush=True)
        del col
