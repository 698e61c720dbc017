"""Third ablation (int16 and float64 lists, 15 packing threads): D2H deferred to one copy kernel at the end, zero fill skipped."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from lasr_b200 import host_pipeline as hp
from bench import make_list
lists = [make_list(s)[0] for s in (1, 101, 201, 301)]
l16 = [[np.round(w * 32767).astype(np.int16) for w in l] for l in lists]
orig = hp.stale_ranges
for kind, ls in (("int16", l16), ("float64", lists)):
    for name in ("baseline", "d2h at the end", "no zero fill", "d2h at the end + no zero fill", "to_host=False"):
        hp.stale_ranges = (lambda dirty, valid, nbytes: ([], dirty)) if "no zero" in name else orig
        col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=name != "to_host=False", cmvn="utt_meanvar")
        if "at the end" in name:
            col.pipeline.d2h_mode = "kernel_end"
        for i in range(5):
            col(ls[i % 4])
        torch.cuda.synchronize()
        ts = []
        for i in range(24):
            t0 = time.perf_counter(); col(ls[i % 4]); ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        print("%-8s %-32s median %.2f ms  p90 %.2f" % (kind, name, ts[12], ts[21]), flush=True)
        del col
hp.stale_ranges = orig
