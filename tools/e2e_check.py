"""Ad-hoc: extract_host with the copy kernel vs per-utterance DMA: equality + timing on the C2 batch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = "cuda:0"
rng = np.random.default_rng(1)
n = np.round(rng.uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
nm = int((n.max() + 3) // 4 * 4)
w = np.zeros((256, nm), dtype=np.float32)
for i in range(256):
    w[i, : n[i]] = np.clip(rng.normal(0, 0.1, n[i]), -1, 1)
wp = torch.from_numpy(w).pin_memory()
wi = torch.from_numpy(np.round(w * 32767).astype(np.int16)).pin_memory()
hours = n.sum() / 16000 / 3600
res = {}
for gb in (32 << 20, 64 << 20, 16 << 20):
    for kc in (True, False):
        fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
        fe.kernel_copies = kc
        for src, name in ((wp, "f32"), (wi, "i16")):
            for _ in range(2): hf, hl = fe.extract_host(src, n, device=dev, group_bytes=gb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K = 10
            for _ in range(K): hf, hl = fe.extract_host(src, n, device=dev, group_bytes=gb)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / K
            key = (name, gb)
            if key in res: print("   equal to kernel-copy result:", bool(torch.equal(res[key], hf)))
            else: res[key] = hf.clone()
            for _ in range(2): fe.extract_host(src, n, device=dev, group_bytes=gb, return_host=False)
            torch.cuda.synchronize(); e0.record()
            for _ in range(K): fe.extract_host(src, n, device=dev, group_bytes=gb, return_host=False)
            e1.record(); torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / K
            print("group %3d MB kernel_copies=%s %s: e2e %.3f ms (%.1f audio-h/s)   features stay on device %.3f ms (%.1f)" % (gb >> 20, kc, name, ms, hours / ms * 1e3, ms2, hours / ms2 * 1e3), flush=True)
