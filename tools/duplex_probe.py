"""PCIe in both directions at once, in the chunked pattern of the host pipeline: G groups, H2D(g) -> [compute g] -> D2H(g), with H2D(g+1)
overlapping D2H(g).  No compute here: only the copies, with the byte counts of the int16 C2 call (147 MB up, 141 MB down) and of
the float32 staging of the float64 call (295 MB up, 141 MB down)."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = torch.device("cuda:0")
lib = lasr_b200._lib.load()
G = int(os.environ.get("GROUPS", "5"))
res = {}
for label, up_mb in (("int16", 147.3), ("float32", 294.6)):
    up = int(up_mb * 1e6 / G) // 16 * 16
    dn = int(141.0e6 / G) // 16 * 16
    h_in = torch.zeros((up * G,), dtype=torch.uint8, pin_memory=True)
    d_in = torch.zeros((up * G,), dtype=torch.uint8, device=dev)
    d_out = torch.zeros((dn * G,), dtype=torch.uint8, device=dev)
    h_out = torch.zeros((dn * G,), dtype=torch.uint8, pin_memory=True)
    tab = torch.from_numpy(np.stack([np.arange(G, dtype=np.int64) * dn, np.full(G, dn, dtype=np.int64)])).to(dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(mode):
        evs = []
        t0 = torch.cuda.Event(enable_timing=True); t0.record(s_in)
        s_out.wait_event(t0)
        for g in range(G):
            if mode != "d2h_only":
                with torch.cuda.stream(s_in):
                    d_in[g * up:(g + 1) * up].copy_(h_in[g * up:(g + 1) * up], non_blocking=True)
            e = torch.cuda.Event(enable_timing=True); e.record(s_in)
            if mode == "h2d_only":
                evs.append((e, e)); continue
            if mode != "serial":
                s_out.wait_event(e)
            elif g == 0:
                pass
            if mode == "serial" and g == 0:
                pass
            if mode in ("kernel", "serial_kernel", "d2h_only"):
                if mode == "serial_kernel" and g == 0:
                    last = torch.cuda.Event(); 
                lasr_b200._lib.check(lib.b200fe_copy_ragged(C.c_void_p(d_out.data_ptr()), C.c_void_p(tab.data_ptr() + 8 * g), C.c_void_p(h_out.data_ptr()),
                                                             C.c_void_p(tab.data_ptr() + 8 * g), C.c_void_p(tab.data_ptr() + 8 * (G + g)), 1, dn, C.c_void_p(s_out.cuda_stream)), "copy")
            else:
                with torch.cuda.stream(s_out):
                    h_out[g * dn:(g + 1) * dn].copy_(d_out[g * dn:(g + 1) * dn], non_blocking=True)
            e2 = torch.cuda.Event(enable_timing=True); e2.record(s_out)
            evs.append((e, e2))
        torch.cuda.synchronize()
        return [(round(t0.elapsed_time(a), 2), round(t0.elapsed_time(b), 2)) for a, b in evs]

    for mode in ("h2d_only", "d2h_only", "kernel", "dma"):
        run(mode)
        tl = run(mode)
        res["%s %s" % (label, mode)] = {"h2d_done_ms": [a for a, _ in tl], "d2h_done_ms": [b for _, b in tl], "total_ms": max(max(a, b) for a, b in tl)}
print(json.dumps(res, indent=1))
