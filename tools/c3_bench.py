"""BASELINE config 3: the training front end -- fbank + global CMVN + SpecAugment (2 frequency / 2 time masks) at batch 512
(512 x 10 s, seed 2) on one B200.  Reports, per variant, the device time of a step, audio-h/s, the algorithmic GB/s of the whole
step and the host time spent replaying the reference's RNG draws for the masks."""
import sys, os, json, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200

dev = "cuda:0"
B, N = 512, 160000
g = torch.Generator(device=dev); g.manual_seed(2)
wav = (torch.randn((B, N), device=dev, generator=g) * 0.1).clamp_(-1, 1)
n = np.full(B, N, dtype=np.int64)
T = 1 + (N - 400) // 160
hours = B * N / 16000 / 3600
alg = 4 * B * N + 320 * B * T + 8 * B
stats = lasr_b200.GpuFbankFrontend().accumulate_stats(wav, n).cpu().numpy()
out = torch.empty((B, T, 80), device=dev)
res = []
for name, kw in (("fbank + global CMVN", dict()),
                 ("+ SpecAugment masks, zero fill (one fused launch)", dict(specaug=True, replace_with_zero=True)),
                 ("+ SpecAugment masks, mean fill (reference default)", dict(specaug=True)),
                 ("+ time warp + masks, mean fill (registry transform `specaug`)", dict(specaug=True, time_warp=True))):
    fe = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=stats, **kw)
    random.seed(2); np.random.seed(2)
    use_out = None if kw.get("time_warp") else out
    for _ in range(3): fe(wav, n, out=use_out)
    torch.cuda.synchronize()
    K = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K): fe(wav, n, out=use_out)
    e1.record()
    cpu = (time.perf_counter() - t0) / K * 1e3
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    plan_ms = 0.0
    if kw.get("specaug"):
        t0 = time.perf_counter()
        for _ in range(K): lasr_b200.specaug.plan_batch(np.full(B, T), 80, return_warp=bool(kw.get("time_warp")), **fe.sa)
        plan_ms = (time.perf_counter() - t0) / K * 1e3
    r = dict(config="C3 512 x 10 s, " + name, ms_per_step=ms, audio_hours_per_s=hours / ms * 1e3, algorithmic_gb_per_s=alg / ms / 1e6,
             host_enqueue_ms=cpu, host_mask_planning_ms=plan_ms)
    res.append(r)
    print(json.dumps(r), flush=True)
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "c3_bench.json"), "w"), indent=1)
