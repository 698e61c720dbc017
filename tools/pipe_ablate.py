"""Which host-side activity slows the PCIe copies of a B200Collate call?  int16 and float64 C2 lists; ablations change timing only
(the zero fill of stale padding is skipped in one of them, so its output is not valid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from lasr_b200 import host_pipeline as hp
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

orig_stale = hp.stale_ranges
for kind, ls in (("int16", l16), ("float64", lists)):
    for name, kw, patch in (("baseline", {}, None), ("threads=8", {"threads": 8}, None), ("threads=4", {"threads": 4}, None),
                            ("no zero fill (invalid output)", {}, "nozero"), ("group_bytes 16 MB", {"group_bytes": 16 << 20}, None),
                            ("group_bytes 64 MB", {"group_bytes": 64 << 20}, None), ("to_host=False", {"to_host": False}, None)):
        hp.stale_ranges = (lambda dirty, valid, nbytes: ([], dirty)) if patch == "nozero" else orig_stale
        kw2 = dict(kw)
        to_host = kw2.pop("to_host", True)
        gb = kw2.pop("group_bytes", None)
        col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=to_host, cmvn="utt_meanvar", **kw2)
        if gb:
            col.pipeline.group_bytes = gb
        ms = bench_col(col, ls)
        print("%-8s %-32s %.3f ms per call, zero fill %.1f MB per call" % (kind, name, ms, getattr(col.pipeline, "zero_bytes", 0) / 1e6), flush=True)
        del col
hp.stale_ranges = orig_stale
