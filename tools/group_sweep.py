"""Synchronous B200Collate call time against the utterance-group size of the host pipeline (float64 and int16 C2 lists)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from bench import make_list
lists = [make_list(s)[0] for s in (1, 101)]
l16 = [[np.round(w * 32767).astype(np.int16) for w in l] for l in lists]
for mb in (8, 16, 32, 64, 128):
    col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=True, cmvn="utt_meanvar")
    col.pipeline.group_bytes = mb << 20
    out = []
    for name, ls in (("float64", lists), ("int16", l16)):
        for i in range(4):
            col(ls[i % 2])
        t0 = time.perf_counter()
        for i in range(10):
            col(ls[i % 2])
        sync_ms = (time.perf_counter() - t0) * 100
        t0 = time.perf_counter()
        for _ in col.prefetch(ls[i % 2] for i in range(10)):
            pass
        pf_ms = (time.perf_counter() - t0) * 100
        out.append("%s sync %.2f ms prefetch %.2f ms" % (name, sync_ms, pf_ms))
    print("group %3d MB: %s" % (mb, " | ".join(out)), flush=True)
    del col
