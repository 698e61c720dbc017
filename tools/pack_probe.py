"""Bare host packing rate (float64 list -> pinned float32 staging, b200fe_host_pack_begin) against the number of pool threads,
on the C2-shaped list of seed 1: is the plug-in call's packing stage bound by cores or by the memory system?"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from importlib import import_module
_lib = import_module("lighting-asr_b200._lib")
lib = _lib.load()
wavs, n = bench.make_list(bench.BATCH_SEEDS[0], 0, 1, np.float64)
B = len(wavs)
offs = np.zeros(B, dtype=np.int64)
np.cumsum((n[:-1] + 3) // 4 * 4, out=offs[1:])
total = int(offs[-1] + (n[-1] + 3) // 4 * 4)
dst = torch.empty((total + 64,), dtype=torch.float32, pin_memory=True)
ptrs = (C.c_void_p * B)(*[a.__array_interface__["data"][0] for a in wavs])
src_bytes, dst_bytes = int(n.sum()) * 8, int(n.sum()) * 4
info = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "src_MB": src_bytes / 1e6, "dst_MB": dst_bytes / 1e6}
try:
    info["model"] = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    pass
res = {}
for mode in os.environ.get("PACK_MODES", "0").split(","):
    os.environ["B200FE_CVT_MODE"] = mode
    for nt in (2, 4, 8, 12, 16, 24, 32, 48):
        pool = C.c_void_p()
        _lib.check(lib.b200fe_host_pool_create(nt, C.byref(pool)), "pool")
        ts = []
        for rep in range(6):
            t0 = time.perf_counter()
            tk = lib.b200fe_host_pack_begin(pool, ptrs, C.c_void_p(n.ctypes.data), B, 2, C.c_void_p(dst.data_ptr()), C.c_void_p(offs.ctypes.data), dst.numel())
            assert tk > 0
            _lib.check(lib.b200fe_host_wait(pool, tk), "wait")
            ts.append(time.perf_counter() - t0)
        lib.b200fe_host_pool_destroy(pool)
        best = min(ts[1:])
        res["mode%s_threads%d" % (mode, nt)] = {"ms": round(best * 1e3, 3), "GBps_read_plus_write": round((src_bytes + dst_bytes) / best / 1e9, 1)}
print(json.dumps({"info": info, "results": res}, indent=1))
