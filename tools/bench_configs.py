"""BASELINE.json configs 1, 3, 4 and 5 as functions (bench.py calls them for `--config cN` and, in short form, for the
`extra` block of the default C2 line).  Every function times the device-resident path with CUDA events on the launching
stream after warm-up and returns a dict; shapes and seeds follow SURVEY.md 8(d)."""
import os
import random
import time

import numpy as np
import torch

SR = 16000.0


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _timed(fn, steps, warmup, dev):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = _events()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def c1_inputs():
    """seed 0; 16 x uniform(-0.5, 0.5) float64 (160000,) -- what soundfile.read returns."""
    rng = np.random.default_rng(0)
    return [rng.uniform(-0.5, 0.5, 160000) for _ in range(16)]


def run_c1(dev, steps=50, warmup=5):
    """C1: 16 utterances x 10 s, fbank:80 only (dither 0).  16 tiles of work cannot fill 148 SMs: latency, not throughput."""
    import lasr_b200
    wavs = c1_inputs()
    fe = lasr_b200.GpuFbankFrontend()
    wav = torch.from_numpy(np.stack(wavs).astype(np.float32)).to(dev)
    n = torch.full((16,), 160000, dtype=torch.int64, device=dev)
    T = 998
    out = torch.empty((16, T, 80), dtype=torch.float32, device=dev)
    ol = torch.empty((16,), dtype=torch.int64, device=dev)
    ms = _timed(lambda: fe(wav, n, max_frames=T, out=out, out_len=ol), steps, warmup, dev)
    hours = 16 * 10.0 / 3600.0
    alg = 4 * 16 * 160000 + 320 * 16 * T + 8 * 16
    col = lasr_b200.lasr_plugin.B200Collate(dev, to_host=True)
    for _ in range(3):
        col(wavs)
    t0 = time.perf_counter()
    for _ in range(steps):
        col(wavs)
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    # A10: the registry transform ASRProcess.frontend calls once per file (asrprocess.py:49-56): numpy in -> features out
    lat = {}
    for key, kw in (("ndarray_out", dict(return_tensor=False)), ("cuda_tensor_out", dict(return_tensor=True))):
        tr = lasr_b200.lasr_plugin.GpuTransform(dev, **kw)
        for _ in range(5):
            tr(wavs[0])
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(50):
            t0 = time.perf_counter()
            tr(wavs[0])
            torch.cuda.synchronize(dev)
            ts.append((time.perf_counter() - t0) * 1e6)
        lat[key] = {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99))}
    return {"workload": "C1: 16 utts x 10 s @16 kHz, uniform(-0.5,0.5), fbank:80, dither 0", "ms_per_step": ms, "value": hours / (ms * 1e-3),
            "gpu_transform_latency_10s_utterance": lat,
            "unit": "audio-h/s", "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6,
            "e2e": {"value": hours / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "api": "B200Collate(list of 16 float64 ndarrays) -> pinned host batch"},
            "note": "32 frame tiles x 16 utterances = 512 tiles on 296 resident CTAs: a latency-sized launch (the reference's CPU-runnable case)"}


def c3_inputs(dev, B=512, N=160000):
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    return (torch.randn((B, N), device=dev, generator=g) * 0.1).clamp_(-1, 1), np.full(B, N, dtype=np.int64)


def run_c3(dev, steps=20, warmup=3, variants=("global", "zero", "mean", "full")):
    """C3: 512 x 10 s, fbank + global CMVN (stats pass over the same batch) + SpecAugment 2F/2T."""
    import lasr_b200
    B, N = 512, 160000
    wav, n = c3_inputs(dev, B, N)
    T = 1 + (N - 400) // 160
    hours = B * N / SR / 3600.0
    alg = 4 * B * N + 320 * B * T + 8 * B
    stats = lasr_b200.GpuFbankFrontend().accumulate_stats(wav, n).cpu().numpy()
    out = torch.empty((B, T, 80), device=dev)
    table = {"global": ("fbank + global CMVN", dict()),
             "zero": ("+ SpecAugment masks 2F/2T, zero fill (one fused launch)", dict(specaug=True, replace_with_zero=True)),
             "mean": ("+ SpecAugment masks 2F/2T, mean fill (reference default; north_star scope)", dict(specaug=True)),
             "full": ("+ time warp + masks, mean fill (registry transform `specaug`, PIL-exact)", dict(specaug=True, time_warp=True))}
    res = {}
    for key in variants:
        name, kw = table[key]
        fe = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=stats, **kw)
        random.seed(2)
        np.random.seed(2)
        use_out = None if kw.get("time_warp") else out
        fe.launch_count = 0
        ms = _timed(lambda: fe(wav, n, out=use_out), steps, warmup, dev)
        res[key] = {"variant": name, "ms_per_step": ms, "value": hours / (ms * 1e-3), "unit": "audio-h/s", "achieved_gbs": alg / ms / 1e6,
                    "launches_per_step": fe.launch_count / (steps + warmup)}
    return {"workload": "C3: 512 utts x 10 s @16 kHz N(0,0.1^2), fbank + global CMVN + SpecAugment (2 freq F=27 / 2 time T=40)",
            "algorithmic_bytes": alg, "audio_hours_per_step": hours, "variants": res, "global_stats": stats}


def run_c5(dev, pushes=200, streams=(1, 64, 4096)):
    """C5: 40 ms chunks at 16 kHz (640 samples) and 8 kHz (320 samples), S concurrent streams.  Latency = host-synchronised
    push (launch -> features ready); throughput without per-push synchronisation."""
    import lasr_b200
    res = []
    for sf, chunk in ((16000.0, 640), (8000.0, 320)):
        for S in streams:
            st = lasr_b200.StreamingFbank(S, device=dev, sample_frequency=sf)
            audio = torch.rand((S, chunk), device=dev) - 0.5
            for _ in range(10):
                st.push(audio)
            torch.cuda.synchronize(dev)
            lat = []
            for _ in range(pushes):
                t0 = time.perf_counter()
                f = st.push(audio)
                torch.cuda.synchronize(dev)
                lat.append((time.perf_counter() - t0) * 1e6)
            e0, e1 = _events()
            e0.record()
            for _ in range(pushes):
                st.push(audio)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            hours = S * chunk * pushes / sf / 3600.0
            res.append({"sample_rate": sf, "streams": S, "chunk_ms": 40, "frames_per_push": int(f.shape[1]),
                        "latency_us_p50": float(np.percentile(lat, 50)), "latency_us_p99": float(np.percentile(lat, 99)),
                        "value": hours / (ms * 1e-3), "unit": "audio-h/s", "us_per_push_async": ms * 1e3 / pushes})
    # independent streams behind the C ABI's stream handle: every push feeds a random half of the streams with chunks of 20-60 ms
    # from HOST memory (pinned staging + H2D + graph replay of the three kernels + D2H of the frame counts, host-synchronised)
    ind = []
    rng = np.random.default_rng(5)
    for S in streams:
        st = lasr_b200.IndependentStreams(S, device=dev, max_chunk=960)
        ids_all = np.arange(S)
        def one_push():
            k = max(1, S // 2)
            ids = rng.choice(ids_all, size=k, replace=False) if S > 1 else ids_all
            lens = rng.integers(320, 961, size=len(ids))
            return ids.tolist(), [np.zeros(int(n), dtype=np.float32) for n in lens], int(lens.sum())
        for _ in range(5):
            i_, c_, _n = one_push()
            st.push(i_, c_)
        lat, samples = [], 0
        t_all = time.perf_counter()
        for _ in range(max(20, pushes // 4)):
            i_, c_, n_ = one_push()
            t0 = time.perf_counter()
            st.push(i_, c_)
            lat.append((time.perf_counter() - t0) * 1e6)
            samples += n_
        wall = time.perf_counter() - t_all
        # device-resident chunks (ids / lengths / samples already in HBM): the graph replay alone, all S streams, 40 ms each
        st.d_meta[0].copy_(torch.arange(S, dtype=torch.int32, device=dev))
        st.d_meta[1].fill_(640)
        for _ in range(5):
            st._launch(S)
        torch.cuda.synchronize(dev)
        e0, e1 = _events()
        e0.record()
        for _ in range(pushes):
            st._launch(S)
        e1.record()
        torch.cuda.synchronize(dev)
        dev_us = e0.elapsed_time(e1) * 1e3 / pushes
        ind.append({"streams": S, "streams_per_push": max(1, S // 2), "chunk_ms": "20-60 (ragged)", "latency_us_p50": float(np.percentile(lat, 50)),
                    "device_resident_push_us": dev_us, "device_resident_value": S * 640 / 16000.0 / 3600.0 / (dev_us * 1e-6),
                    "latency_us_p99": float(np.percentile(lat, 99)), "value": samples / 16000.0 / 3600.0 / (sum(lat) * 1e-6), "unit": "audio-h/s",
                    "note": "host lists in, CUDA features out; includes building the random test chunks' staging copy"})
        del st
    return {"workload": "C5: 40 ms chunked streaming fbank, S lock-step streams, 16 kHz (640-sample chunks) and 8 kHz (320)", "rows": res,
            "independent_streams_16k": ind}


def run_c4(dev, rank, world, total_hours=1000.0, pool_hours=10.0, verify=False):
    """C4: resident pool of `pool_hours` of LibriSpeech-shaped audio per GPU (seed 3 + rank), swept until the job covered
    `total_hours`; pass 1 = global CMVN statistics, one all-reduce, pass 2 = fbank + global CMVN."""
    import torch.distributed as dist
    import lasr_b200
    B = 256
    g = torch.Generator(device=dev)
    g.manual_seed(3 + rank)
    rng = np.random.default_rng(3 + rank)
    pool, got = [], 0.0
    while got < pool_hours:
        n = np.round(rng.uniform(1.0, 35.0, B) * SR).astype(np.int64)
        nmax = int((n.max() + 3) // 4 * 4) + 64              # room for the rotating offset
        pool.append(((torch.randn((B, nmax), device=dev, generator=g) * 0.1).clamp_(-1, 1), n))
        got += float(n.sum()) / SR / 3600.0
    sweeps = max(1, int(round(total_hours / world / got)))
    fe = lasr_b200.GpuFbankFrontend()
    stats = torch.zeros((2, 81), dtype=torch.float64, device=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def views(s):
        off = 4 * (s % 16)                                      # rotating, 16-byte aligned sample offset
        return [(w[:, off:], n) for w, n in pool]

    for w, n in views(0)[:2]:
        fe.accumulate_stats(w, n)
    e0, e1 = _events()
    sync()
    e0.record()
    for s in range(sweeps):
        for w, n in views(s):
            fe.accumulate_stats(w, n, stats)
    e1.record()
    sync()
    t1 = e0.elapsed_time(e1)
    sync()
    e0.record()
    lasr_b200.cmvn.allreduce_stats(stats)
    e1.record()
    sync()
    t_ar = e0.elapsed_time(e1)
    fe2 = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=stats.cpu().numpy())
    Tm = max(int(fe.frame_counts(n)[0].max()) for _, n in pool)
    out = torch.empty((B, Tm, 80), device=dev)
    for w, n in views(0)[:2]:
        fe2(w, n, max_frames=Tm, out=out)
    sync()
    e0.record()
    for s in range(sweeps):
        for w, n in views(s):
            fe2(w, n, max_frames=Tm, out=out)
    e1.record()
    sync()
    t2 = e0.elapsed_time(e1)
    red = torch.tensor([t1, t2, t_ar], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    t1, t2, t_ar = (float(x) for x in red.cpu())
    hours = got * sweeps * world
    return {"workload": "C4: %.0f h corpus sweep, %.1f h resident pool per GPU (seed 3+rank), per-utterance sharding, stats pass -> all-reduce(2x81 f64) -> fbank + global CMVN pass" % (hours, got),
            "n_gpus": world, "audio_hours": hours, "pool_hours_per_gpu": got, "sweeps": sweeps,
            "pass1_stats_value": hours / (t1 * 1e-3), "pass2_fbank_global_cmvn_value": hours / (t2 * 1e-3), "unit": "audio-h/s",
            "allreduce_us": t_ar * 1e3, "frames": float(stats[0, 80]), "ms_pass1": t1, "ms_pass2": t2}
