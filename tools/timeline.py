"""Phase timeline of the fused kernel (debug build with -DB200FE_TIMELINE, see fbank_kernel.cuh): where a CTA's tile time goes.

    B200FE_NVCC_EXTRA=-DB200FE_TIMELINE B200FE_LIB=/tmp/libtl.so python tools/timeline.py

Stamps per (CTA, warp, tile): 0 loop top, 1 tile data landed (mbarrier), 2 phase A done, 3 past barrier B1, 4 phase B done,
5 past barrier B2, 6 phase C done.  Prints mean SM cycles per segment over steady-state tiles, per warp, and the
per-tile CTA period."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from bench import make_batch

dev = torch.device("cuda:0")
wav_np, n = make_batch()
fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
wav = torch.from_numpy(wav_np).to(dev)
T, _ = fe.frame_counts(n)
out = torch.empty((len(n), int(T.max()), 80), device=dev)
nd = torch.from_numpy(n).to(dev)
lib = lasr_b200._lib.load()
for _ in range(3):
    fe(wav, nd, max_frames=int(T.max()), out=out)
torch.cuda.synchronize()
lib.b200fe_debug_timeline.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int]
lib.b200fe_debug_timeline(None, 0, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fe(wav, nd, max_frames=int(T.max()), out=out); e1.record()
torch.cuda.synchronize()
buf = np.zeros((296, 8, 48, 8), dtype=np.int64)
assert lib.b200fe_debug_timeline(buf.ctypes.data, buf.nbytes, 0) == 0
ms = e0.elapsed_time(e1)
nvalid = buf[..., 7] & 0xffffffff
full = (nvalid == 32) & (buf[..., 6] > 0)
full[:, :, :2] = False                       # skip the first tiles (cold start)
names = ["wait tile data (0->1)", "phase A (1->2)", "barrier B1 wait (2->3)", "phase B (3->4)", "barrier B2 wait (4->5)", "phase C (5->6)"]
seg = np.diff(buf[..., :7], axis=-1).astype(np.float64)
res = {"step_ms_instrumented": ms, "tiles_sampled": int(full[:, 0].sum())}
print("step %.4f ms; full tiles sampled per warp: %d" % (ms, full[:, 0].sum()))
for k, nm in enumerate(names):
    per_warp = [float(seg[:, w, :, k][full[:, w]].mean()) for w in range(8)]
    res[nm] = {"mean_cycles": float(np.mean(per_warp)), "per_warp": per_warp}
    print("%-26s mean %7.0f  per warp %s" % (nm, np.mean(per_warp), " ".join("%6.0f" % x for x in per_warp)))
# CTA tile period: loop top to loop top of consecutive full tiles (warp 0)
top = buf[:, 0, :, 0].astype(np.float64)
per = np.diff(top, axis=1)
ok = full[:, 0, :-1] & full[:, 0, 1:]
res["cta_tile_period_cycles"] = float(per[ok].mean())
print("CTA tile period %.0f cycles (= %.1f cycles per frame per SM with 2 CTAs/SM)" % (per[ok].mean(), per[ok].mean() / 64))
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "timeline.json"), "w"), indent=1)
