"""A/B of kernel variants (alternative builds selected with B200FE_LIB): quick parity against live torchaudio + device time of
the plain fused launch and of the C2 step (utterance CMVN) on C2-shaped lengths.  One JSON line per run."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200
from torchaudio.compliance import kaldi

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("B200FE_LIB", "default"))
n = np.round(np.random.default_rng(1).uniform(1.0, 35.0, 256) * 16000).astype(np.int64)
nmax = int((n.max() + 3) // 4 * 4)
g = torch.Generator(device=dev); g.manual_seed(1)
wav = (torch.randn((256, nmax), device=dev, generator=g) * 0.1).clamp_(-1, 1)
fe = lasr_b200.GpuFbankFrontend()
fc = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
T, _ = fe.frame_counts(n)
Tmax = int(T.max())
out = torch.empty((256, Tmax, 80), device=dev)
nd = torch.from_numpy(n).to(dev)
bad = cells = 0
f, _ = fe(wav[:3], n[:3])
for i in range(3):
    ref = kaldi.fbank(wav[i:i + 1, : n[i]].cpu() * 32768.0, num_mel_bins=80, dither=0.0, energy_floor=1.0).numpy()
    d = np.abs(f[i, : ref.shape[0]].cpu().numpy() - ref)
    bad += int((d > 1e-5 + 1e-4 * np.abs(ref)).sum()); cells += ref.size
res = {"variant": name, "violations": bad, "cells": cells}
plan = fe.plan(dev)
res["static_mel"] = plan.lib.b200fe_plan_info(plan.handle, 0)
res["ctas_per_sm"] = plan.lib.b200fe_plan_info(plan.handle, 3)
res["warps_per_cta"] = plan.lib.b200fe_plan_info(plan.handle, 9)
res["smem_per_cta"] = plan.lib.b200fe_plan_info(plan.handle, 2)
fa = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
fa.inlaunch_cmvn = True
ref = fc(wav, n, max_frames=Tmax)[0].clone()
got = fa(wav, n, max_frames=Tmax)[0]
torch.cuda.synchronize()
res["inlaunch_ran"] = fa.last["apply_flags"] is not None
if res["inlaunch_ran"]:
    w_, i_ = fa.last["apply_flags"]
    res["inlaunch_error_flag"] = int(w_[i_].item())
    res["inlaunch_max_diff_vs_post_pass"] = float((got - ref).abs().max())
del ref, got
for key, fn in (("plain_ms", lambda: fe(wav, nd, max_frames=Tmax, out=out)), ("c2_step_ms", lambda: fc(wav, nd, max_frames=Tmax, out=out)),
                ("c2_inlaunch_ms", lambda: fa(wav, n, max_frames=Tmax, out=out))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / 20)
    res[key] = min(best)
print(json.dumps(res), flush=True)
