#!/usr/bin/env python
"""Benchmark of the B200 acoustic front end (see DESIGN.md section 6 for the definitions).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference [--steps K --warmup W]  # the reference's CPU path on the host cores

A "step" is one pass of the hot path over one synthetic batch of BASELINE.json configs[1]:
256 utterances, durations uniform(1, 35) s at 16 kHz, N(0, 0.1^2) clipped to +-1, zero padded, ->
80-dim Kaldi fbank + utterance CMVN (mean, variance) + zero-padded (B, Tmax, 80) / length tensors.
Under torchrun every rank processes its own batch of that shape (weak scaling, no data-path
collective: utterance CMVN needs none); the only collective of the path, the all-reduce of the
2 x 81 global-CMVN statistics, is timed separately and reported under "extra".
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "C2: 256 utts, 1-35 s @16 kHz (LibriSpeech-shaped), 80-dim kaldi fbank + utterance CMVN (mean,var), zero-padded (B,Tmax,80) + lengths"
SR = 16000.0


def make_batch(rank=0, world=1, B=256):
    """SURVEY 8(d) C2 inputs: durations uniform(1,35) s from seed 1, N(0,0.1^2) clipped to +-1.  With N ranks the global
    batch is N x 256 utterances sharded per utterance, length-balanced (cmvn.shard_utterances): every rank
    gets the same amount of audio to within one utterance, so the max-over-ranks time measures the machine, not the draw."""
    import lasr_b200
    n_all = np.round(np.random.default_rng(1).uniform(1.0, 35.0, B * world) * SR).astype(np.int64)
    mine = lasr_b200.cmvn.shard_utterances(n_all, world)[rank] if world > 1 else np.arange(B)
    n = n_all[mine]
    nmax = int((n.max() + 3) // 4 * 4)
    wav = np.zeros((len(n), nmax), dtype=np.float32)
    for i, u in enumerate(mine):
        wav[i, : n[i]] = np.clip(np.random.default_rng([1, int(u)]).normal(0.0, 0.1, n[i]), -1.0, 1.0).astype(np.float32)
    return wav, n


def algorithmic_bytes(n, T, B):
    """4 B per sample read once + 80*4 B per frame written once + 8 B per length."""
    return 4 * int(n.sum()) + 320 * int(T.sum()) + 8 * B


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # the median over the samples with the highest load (the timed loop keeps the GPU busy)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def reference_arm(args, rank, world, out):
    """The reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from oracle import cpu_baseline
    wav, n = make_batch()
    wavs = [wav[i, : n[i]].astype(np.float64) for i in range(len(n))]      # soundfile.read hands float64 to the transforms
    cores = os.cpu_count() or 1
    hours = float(n.sum()) / SR / 3600.0
    pool = cpu_baseline.make_pool(wavs, cores)
    try:
        for _ in range(max(args.warmup, 1)):
            cpu_baseline.run_chain(wavs[: 2 * cores], "utt_meanvar", False, cores, pool)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_baseline.run_chain(wavs, "utt_meanvar", False, cores, pool)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
        pool.join()
    val = hours * args.steps / dt
    line = {"impl": "reference", "metric": "audio-hours/sec", "value": val, "unit": "audio-h/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "inputs": "host float64 waveforms (what soundfile.read returns)"},
            "cpu_baseline": {"value": val, "unit": "audio-h/s", "cores": cores, "kind": "port",
                             "sample": "the full 256-utterance C2 batch per step (%.3f audio-h): torchaudio.compliance.kaldi.fbank via the oracle's "
                                       "restatement of WavToKaldiFbank + fp64 utterance CMVN + batch_list, %d single-threaded worker processes" % (hours, cores)},
            "e2e": {"value": val, "unit": "audio-h/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


def _claim_stdout():
    """Everything that libraries print to stdout (e.g. NCCL's version banner) is sent to stderr; the
    returned file object is the real stdout, used for the single JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="device-resident loop only (for ncu launch lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world, real_stdout)
        return

    import torch
    import torch.distributed as dist
    import lasr_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the front end has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wav_np, n = make_batch(rank, world)
    B = len(n)
    fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
    T, _ = fe.frame_counts(n)
    hours = float(n.sum()) / SR / 3600.0
    alg_bytes = algorithmic_bytes(n, T, B)
    wav_dev = torch.from_numpy(wav_np).to(dev)
    wav_pin = torch.from_numpy(wav_np).pin_memory()
    out = torch.empty((B, int(T.max()), 80), dtype=torch.float32, device=dev)
    out_len = torch.empty((B,), dtype=torch.int64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value, roofline): waveforms AND sample counts live in HBM (no upload, no host sync per step) ----
    n_host = n
    Tmax = int(T.max())
    n = torch.from_numpy(n_host).to(dev)

    def step():
        fe(wav_dev, n, max_frames=Tmax, out=out, out_len=out_len)

    for _ in range(args.warmup):
        step()
    barrier()
    fe.launch_count = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = fe.launch_count
    # The same K steps again with a CUDA event pair around every fused launch (roofline.achieved).  Kept out of the loop above:
    # an event record between the work-list builder and the fused launch costs ~5 us per step and keeps the fused kernel's
    # prologue from overlapping the builder (programmatic dependent launch); the instrumented step time is reported next to it.
    fe.profile_events = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_instrumented = e0.elapsed_time(e1)
    n = n_host
    fused_ms = [a.elapsed_time(b) for a, b in fe.profile_events]
    fe.profile_events = None

    if args.profile_only:
        print(json.dumps({"profile_only": True, "ms_per_step": ms_total / args.steps, "fused_ms_per_step": sum(fused_ms) / args.steps,
                          "gpu_launches": launches}), file=real_stdout, flush=True)
        return
    # ---- end to end through the public host API: pinned host waveforms in, host features out ----
    # The step's input is the batch as a data loader hands it over: the utterances back to back in ONE pinned buffer
    # (GpuFbankFrontend.pack_host; the reference's collate receives them as a list, dataset.py:190-206).  Output: the
    # reference's padded (B, Tmax, 80) float32 batch + frame counts in pinned host memory.
    pk_pin, pk_len, pk_off = lasr_b200.GpuFbankFrontend.pack_host([wav_np[i, : n[i]] for i in range(B)])
    for _ in range(max(2, args.warmup)):
        fe.extract_host(pk_pin, pk_len, device=dev, wav_offsets=pk_off)
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        hf, hl = fe.extract_host(pk_pin, pk_len, device=dev, wav_offsets=pk_off)
    g1.record()
    barrier()
    e2e_ms = g0.elapsed_time(g1)
    h2d, d2h = fe.h2d_bytes, fe.d2h_bytes
    # the same with the zero-padded (B, Nmax) host tensor batch_list builds (one copy kernel over the valid samples)
    for _ in range(2):
        fe.extract_host(wav_pin, n, device=dev)
    barrier()
    g0.record()
    for _ in range(args.steps):
        fe.extract_host(wav_pin, n, device=dev)
    g1.record()
    barrier()
    e2e_pad_ms = g0.elapsed_time(g1)
    # H2D only, features stay on the device for the encoder (the training-loop case)
    for _ in range(2):
        fe.extract_host(pk_pin, pk_len, device=dev, wav_offsets=pk_off, return_host=False)
    barrier()
    g0.record()
    for _ in range(args.steps):
        fe.extract_host(pk_pin, pk_len, device=dev, wav_offsets=pk_off, return_host=False)
    g1.record()
    barrier()
    e2e_dev_ms = g0.elapsed_time(g1)
    # int16 PCM host input (what the audio files hold): half the H2D bytes
    pcm_pin, pcm_len, pcm_off = lasr_b200.GpuFbankFrontend.pack_host([np.round(wav_np[i, : n[i]] * 32767.0).astype(np.int16) for i in range(B)],
                                                                      dtype=torch.int16)
    for _ in range(2):
        fe.extract_host(pcm_pin, pcm_len, device=dev, wav_offsets=pcm_off)
    barrier()
    g0.record()
    for _ in range(args.steps):
        fe.extract_host(pcm_pin, pcm_len, device=dev, wav_offsets=pcm_off)
    g1.record()
    barrier()
    e2e_i16_ms = g0.elapsed_time(g1)
    h2d_i16, d2h_i16 = fe.h2d_bytes, fe.d2h_bytes
    # packed feature output: (sum T, 80) without padding rows, one DMA per group in both directions
    for _ in range(2):
        fe.extract_host(pk_pin, pk_len, device=dev, wav_offsets=pk_off, packed_out=True)
    barrier()
    g0.record()
    for _ in range(args.steps):
        fe.extract_host(pk_pin, pk_len, device=dev, wav_offsets=pk_off, packed_out=True)
    g1.record()
    barrier()
    e2e_pk_ms = g0.elapsed_time(g1)
    for _ in range(2):
        fe.extract_host(pcm_pin, pcm_len, device=dev, wav_offsets=pcm_off, packed_out=True)
    barrier()
    g0.record()
    for _ in range(args.steps):
        fe.extract_host(pcm_pin, pcm_len, device=dev, wav_offsets=pcm_off, packed_out=True)
    g1.record()
    barrier()
    e2e_pk16_ms = g0.elapsed_time(g1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- the path's only collective: all-reduce of the global CMVN statistics ----
    ar_us = None
    if world > 1:
        st = torch.zeros((2, 81), dtype=torch.float64, device=dev)
        for _ in range(5):
            lasr_b200.cmvn.allreduce_stats(st)
        barrier()
        g0.record()
        for _ in range(20):
            lasr_b200.cmvn.allreduce_stats(st)
        g1.record()
        barrier()
        ar_us = g0.elapsed_time(g1) / 20 * 1e3

    # max over ranks of the device time, sum over ranks of the work
    red = torch.tensor([ms_total, e2e_ms, e2e_dev_ms, e2e_i16_ms, e2e_pad_ms, e2e_pk_ms, e2e_pk16_ms], dtype=torch.float64, device=dev)
    work = torch.tensor([hours, float(alg_bytes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(work, op=dist.ReduceOp.SUM)
    ms_total, e2e_ms, e2e_dev_ms, e2e_i16_ms, e2e_pad_ms, e2e_pk_ms, e2e_pk16_ms = (float(x) for x in red.cpu())
    hours_all = float(work[0])

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fused_per_step_ms = sum(fused_ms) / args.steps           # the dominant kernel: all fused launches of one step
        achieved = alg_bytes / (fused_per_step_ms * 1e-3) / 1e9
        traffic = None
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "fbank_fused_summary.json")))
            traffic = prof.get("dram_bytes_per_step")
        except Exception:  # noqa: BLE001
            pass
        step_ms = ms_total / args.steps
        line = {
            "metric": "audio-hours/sec", "value": hours_all / (step_ms * 1e-3), "unit": "audio-h/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "audio_hours_per_gpu_step": hours,
                       "inputs": "value: float32 waveforms (B, Nmax) and int64 sample counts resident in HBM, features + frame counts written to HBM; e2e: pinned host buffers",
                       "l2": "inputs (%.0f MB/step/GPU) exceed the 126 MB L2; no flush needed" % (wav_np.nbytes / 1e6),
                       "parallelism": "global batch of %d utterances sharded per utterance (length-balanced) x%d, no data-path collective" % (256 * world, world)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs,
                         "traffic": traffic, "kernel": "fbank_fused_kernel<13,true,false,false,false,false,true> = lean instantiation of the default option set (every fused launch of one step, incl. the zero fill of the padded rows by its padding tiles)",
                         "algorithmic_bytes_per_step": alg_bytes, "kernel_ms_per_step": fused_per_step_ms,
                         "kernel_share_of_step": fused_per_step_ms / (ms_instrumented / args.steps), "peak_source": peak_src,
                         "timing": "CUDA event pair around every fused launch in an instrumented repeat of the K timed steps (%.4f ms per step with the events, %.4f without)" % (ms_instrumented / args.steps, step_ms)},
            "e2e": {"value": hours_all / (e2e_ms / args.steps * 1e-3), "unit": "audio-h/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "api": "GpuFbankFrontend.extract_host(pinned float32 utterances packed back to back (pack_host), lengths, offsets) -> "
                           "pinned host (B, Tmax, 80) features + frame counts; one DMA per 32 MB utterance group in, one copy kernel per group out"},
            "gpu_launches": launches,
            "clocks": clocks,
            "extra": {"e2e_features_stay_on_device": {"value": hours_all / (e2e_dev_ms / args.steps * 1e-3), "unit": "audio-h/s",
                                                      "ms_per_step": e2e_dev_ms / args.steps},
                      "e2e_padded_host_input": {"value": hours_all / (e2e_pad_ms / args.steps * 1e-3), "unit": "audio-h/s",
                                                "ms_per_step": e2e_pad_ms / args.steps,
                                                "api": "extract_host(zero-padded pinned (B, Nmax) float32 batch): valid samples gathered by one copy kernel per group"},
                      "e2e_int16_pcm_host_input": {"value": hours_all / (e2e_i16_ms / args.steps * 1e-3), "unit": "audio-h/s",
                                                   "ms_per_step": e2e_i16_ms / args.steps, "h2d_bytes_per_step": h2d_i16,
                                                   "d2h_bytes_per_step": d2h_i16},
                      "e2e_packed_feature_output": {"value": hours_all / (e2e_pk_ms / args.steps * 1e-3), "unit": "audio-h/s",
                                                    "ms_per_step": e2e_pk_ms / args.steps,
                                                    "api": "extract_host(..., packed_out=True): (sum T, 80) pinned features + lengths + row offsets, one DMA per group each way"},
                      "e2e_int16_in_packed_out": {"value": hours_all / (e2e_pk16_ms / args.steps * 1e-3), "unit": "audio-h/s",
                                                  "ms_per_step": e2e_pk16_ms / args.steps},
                      "global_cmvn_stats_allreduce_us": ar_us},
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline
            wavs = [wav_np[i, : n[i]].astype(np.float64) for i in range(B)]
            cores = os.cpu_count() or 1
            r = cpu_baseline.time_chain(wavs, SR, "utt_meanvar", False, cores, min_seconds=10.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "audio-h/s", "cores": cores, "kind": "port",
                                    "sample": "%d x the full C2 batch (%.3f audio-h each) in %.1f s: torchaudio.compliance.kaldi.fbank via the oracle's "
                                              "restatement of WavToKaldiFbank + fp64 utterance CMVN + batch_list, %d single-threaded worker processes"
                                              % (r["reps"], r["audio_hours_per_rep"], r["seconds"], cores)}
        print(json.dumps(line), file=real_stdout, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
