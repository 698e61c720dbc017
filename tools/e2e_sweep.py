"""Ad-hoc: extract_host variants on the C2 batch: host layout (padded / packed), H2D and D2H mechanism (copy kernel / DMA)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
dev = "cuda:0"
GB = int(os.environ.get('E2E_GB', '32')) << 20
rng = np.random.default_rng(1)
n = np.round(rng.uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
nm = int((n.max() + 3) // 4 * 4)
w = np.zeros((256, nm), dtype=np.float32)
for i in range(256):
    w[i, : n[i]] = np.clip(rng.normal(0, 0.1, n[i]), -1, 1)
hours = n.sum() / 16000 / 3600
ref = None
for dt in (torch.float32, torch.int16):
    src = w if dt == torch.float32 else np.round(w * 32767).astype(np.int16)
    wp = torch.from_numpy(src).pin_memory()
    pk, lens, offs = lasr_b200.GpuFbankFrontend.pack_host([src[i, : n[i]] for i in range(256)], dtype=dt)
    for mode in sys.argv[1:]:
        fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
        fe.kernel_h2d = mode[0] == "k"; fe.kernel_d2h = mode[1] == "k"
        packed = mode[0] == "p"; fe.overlap_calls = "n" not in mode[2:]
        call = (lambda rh: fe.extract_host(pk, lens, device=dev, return_host=rh, wav_offsets=offs, group_bytes=GB)) if packed else (lambda rh: fe.extract_host(wp, n, device=dev, return_host=rh, group_bytes=GB))
        out = []
        for rh in (True, False):
            for _ in range(3): hf, hl = call(rh)
            torch.cuda.synchronize()
            K = 10
            t0 = time.perf_counter()
            for _ in range(K): hf, hl = call(rh)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / K * 1e3
            out.append("%s %.3f ms %.1f audio-h/s" % ("host->host" if rh else "host->device", ms, hours / ms * 1e3))
            if rh and dt == torch.float32:
                if ref is None: ref = hf.clone()
                else: out.append("equal=%s" % bool(torch.equal(ref, hf)))
        print("GB", GB >> 20, str(dt), "mode", mode, " | ".join(out), flush=True)
