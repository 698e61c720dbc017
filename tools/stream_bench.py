"""BASELINE config 5: 40 ms chunked streaming front end at 8 kHz and 16 kHz -- per-push latency
(launch -> features ready, host-synchronised) and aggregate throughput for S concurrent streams."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200

dev = "cuda:0"
res = []
for sf, chunk in ((16000.0, 640), (8000.0, 320)):
    for S in (1, 64, 4096):
        st = lasr_b200.StreamingFbank(S, device=dev, sample_frequency=sf)
        audio = (torch.rand((S, chunk), device=dev) - 0.5)
        for _ in range(10):
            st.push(audio)
        torch.cuda.synchronize()
        lat = []
        n = 200
        t_all = time.perf_counter()
        for _ in range(n):
            t0 = time.perf_counter()
            f = st.push(audio)
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e6)
        wall = time.perf_counter() - t_all
        # throughput without per-push synchronisation
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            st.push(audio)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        hours = S * chunk * n / sf / 3600.0
        r = dict(sample_rate=sf, streams=S, chunk_ms=40, frames_per_push=int(f.shape[1]), latency_us_p50=float(np.percentile(lat, 50)),
                 latency_us_p99=float(np.percentile(lat, 99)), audio_hours_per_s=hours / (ms * 1e-3), realtime_factor=S * chunk * n / sf / (ms * 1e-3))
        res.append(r)
        print(json.dumps(r))
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "stream_bench.json"), "w"), indent=1)
