"""D2H of a C2 feature batch (ragged rows into the padded pinned host tensor): copy engine vs the copy kernel, alone and while an
H2D DMA of the waveforms runs in the other direction."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = torch.device("cuda:0")
lib = lasr_b200._lib.load()
n = np.round(np.random.default_rng(1).uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
T = 1 + (n - 400) // 160
Tmax, D, B = int(T.max()), 80, 256
feats = torch.randn((B, Tmax, D), device=dev)
host = torch.zeros((B, Tmax, D), pin_memory=True)
wav_h = torch.zeros((int(n.sum()),), pin_memory=True)
wav_d = torch.zeros((int(n.sum()),), device=dev)
tab = torch.from_numpy(np.stack([np.arange(B, dtype=np.int64) * (Tmax * D * 4), T.astype(np.int64) * (D * 4)])).to(dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
valid_bytes = int(T.sum()) * D * 4

def d2h_kernel():
    lasr_b200._lib.check(lib.b200fe_copy_ragged(C.c_void_p(feats.data_ptr()), C.c_void_p(tab.data_ptr()), C.c_void_p(host.data_ptr()), C.c_void_p(tab.data_ptr()),
                                                 C.c_void_p(tab.data_ptr() + 8 * B), B, int(T.max()) * D * 4, C.c_void_p(s2.cuda_stream)), "copy")
def d2h_dma_rows():
    rows = np.ascontiguousarray(T.astype(np.int64))
    lib.b200fe_d2h_ragged(C.c_void_p(feats.data_ptr()), D, Tmax, C.c_void_p(rows.ctypes.data), B, C.c_void_p(host.data_ptr()), C.c_void_p(s2.cuda_stream))
flat_d = torch.randn((valid_bytes // 4,), device=dev); flat_h = torch.zeros((valid_bytes // 4,), pin_memory=True)
def d2h_one():
    with torch.cuda.stream(s2):
        flat_h.copy_(flat_d, non_blocking=True)
def h2d():
    with torch.cuda.stream(s1):
        wav_d.copy_(wav_h, non_blocking=True)

def timeit(fns, reps=5):
    for f in fns: f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for f in fns: f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

print("valid feature bytes %.1f MB, waveform bytes %.1f MB, copy CTAs env=%s" % (valid_bytes / 1e6, wav_h.numel() * 4 / 1e6, os.environ.get("B200FE_COPY_CTAS")))
for name, fns in (("D2H one contiguous DMA", [d2h_one]), ("D2H copy kernel (ragged -> padded pinned)", [d2h_kernel]), ("D2H one DMA per utterance", [d2h_dma_rows]),
                  ("H2D one DMA", [h2d]), ("H2D DMA + D2H one DMA", [h2d, d2h_one]), ("H2D DMA + D2H copy kernel", [h2d, d2h_kernel]),
                  ("H2D DMA + D2H DMA per utterance", [h2d, d2h_dma_rows])):
    print("%-45s %.3f ms" % (name, timeit(fns)))
