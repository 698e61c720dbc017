"""The chunked duplex copy pattern of tools/duplex_probe.py while the host thread pool packs a C2 list (memory traffic of the
packing stage): does host-memory contention explain the slow copies inside a B200Collate call?"""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200, bench
dev = torch.device("cuda:0")
_lib = lasr_b200._lib
lib = _lib.load()
G = 5
wavs, n = bench.make_list(bench.BATCH_SEEDS[0], 0, 1, np.float64)
w16 = [np.round(w * 32767).astype(np.int16) for w in wavs]
B = len(wavs)
def pack_setup(arrs, code, esz):
    al = 16 // esz
    offs = np.zeros(B, dtype=np.int64)
    np.cumsum((n[:-1] + al - 1) // al * al, out=offs[1:])
    total = int(offs[-1] + (n[-1] + al - 1) // al * al)
    dst = torch.empty((total + 64,), dtype=torch.int16 if code == 1 else torch.float32, pin_memory=True)
    ptrs = (C.c_void_p * B)(*[a.__array_interface__["data"][0] for a in arrs])
    return dict(ptrs=ptrs, offs=offs, dst=dst, code=code)
packs = {"none": None, "int16 memcpy": pack_setup(w16, 1, 2), "float64->float32": pack_setup(wavs, 2, 4)}
up = int(147.3e6 / G) // 16 * 16
dn = int(141.0e6 / G) // 16 * 16
h_in = torch.zeros((up * G,), dtype=torch.uint8, pin_memory=True)
d_in = torch.zeros((up * G,), dtype=torch.uint8, device=dev)
d_out = torch.zeros((dn * G,), dtype=torch.uint8, device=dev)
h_out = torch.zeros((dn * G,), dtype=torch.uint8, pin_memory=True)
tab = torch.from_numpy(np.stack([np.arange(G, dtype=np.int64) * dn, np.full(G, dn, dtype=np.int64)])).to(dev)
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def run(pool, pk, reps_pack):
    evs = []
    tks = []
    if pk is not None:
        for _ in range(reps_pack):
            tks.append(lib.b200fe_host_pack_begin(pool, pk["ptrs"], C.c_void_p(n.ctypes.data), B, pk["code"], C.c_void_p(pk["dst"].data_ptr()), C.c_void_p(pk["offs"].ctypes.data), pk["dst"].numel()))
    t0 = torch.cuda.Event(enable_timing=True); t0.record(s_in)
    s_out.wait_event(t0)
    for g in range(G):
        with torch.cuda.stream(s_in):
            d_in[g * up:(g + 1) * up].copy_(h_in[g * up:(g + 1) * up], non_blocking=True)
        e = torch.cuda.Event(enable_timing=True); e.record(s_in)
        s_out.wait_event(e)
        _lib.check(lib.b200fe_copy_ragged(C.c_void_p(d_out.data_ptr()), C.c_void_p(tab.data_ptr() + 8 * g), C.c_void_p(h_out.data_ptr()),
                                          C.c_void_p(tab.data_ptr() + 8 * g), C.c_void_p(tab.data_ptr() + 8 * (G + g)), 1, dn, C.c_void_p(s_out.cuda_stream)), "copy")
        e2 = torch.cuda.Event(enable_timing=True); e2.record(s_out)
        evs.append((e, e2))
    torch.cuda.synchronize()
    tp = time.perf_counter()
    for tk in tks:
        _lib.check(lib.b200fe_host_wait(pool, tk), "wait")
    return [(round(t0.elapsed_time(a), 2), round(t0.elapsed_time(b), 2)) for a, b in evs], round((time.perf_counter() - tp) * 1e3, 2)

for nt in (16, 8):
    pool = C.c_void_p()
    _lib.check(lib.b200fe_host_pool_create(nt, C.byref(pool)), "pool")
    for name, pk in packs.items():
        reps = 3 if name.startswith("int16") else 1
        run(pool, pk, reps)
        tl, rest = run(pool, pk, reps)
        print("threads %2d, concurrent packing: %-18s copies done after %.2f ms  h2d %s  d2h %s  (packing went on for %.2f ms after the copies)" %
              (nt, name, max(b for _, b in tl), [a for a, _ in tl], [b for _, b in tl], rest), flush=True)
    lib.b200fe_host_pool_destroy(pool)
