"""BASELINE config 4: synthetic corpus sweep sharded per utterance across the GPUs of one box.

    python tools/corpus_sweep.py --total-hours 1000                                    # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/corpus_sweep.py       # N GPUs

Every rank keeps a resident pool of LibriSpeech-shaped synthetic utterances (seed 3 + rank)
and sweeps it until the job has covered ``--total-hours`` of audio (rotating the sample offset per sweep):
pass 1 accumulates the global CMVN statistics (fused kernel in statistics-only mode: no feature output),
ONE all-reduce(sum) of the 2 x 81 float64 statistics over NCCL follows, pass 2 computes fbank + global CMVN.
Prints one JSON line on rank 0: audio-h/s of each pass (max over ranks of the device time), all-reduce time,
and, with --verify, the comparison of the all-reduced statistics with a single-process recomputation.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import lasr_b200

SR = 16000.0


def make_pool(seed, n_batches, dev, B=256):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    pool = []
    for _ in range(n_batches):
        n = np.round(rng.uniform(1.0, 35.0, B) * SR).astype(np.int64)
        nmax = int((n.max() + 3) // 4 * 4) + 64              # room for the rotating offset
        wav = (torch.randn((B, nmax), device=dev, generator=g) * 0.1).clamp_(-1, 1)
        pool.append((wav, n))
    return pool


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-hours", type=float, default=1000.0)
    ap.add_argument("--pool-batches", type=int, default=2)
    ap.add_argument("--verify", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pool = make_pool(3 + rank, args.pool_batches, dev)
    pool_hours = sum(float(n.sum()) for _, n in pool) / SR / 3600.0
    sweeps = max(1, int(round(args.total_hours / world / pool_hours)))
    fe = lasr_b200.GpuFbankFrontend()
    stats = torch.zeros((2, 81), dtype=torch.float64, device=dev)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def views(s):
        off = 4 * (s % 16)                                      # rotating, 16-byte aligned sample offset
        return [(w[:, off:], n) for w, n in pool]

    for w, n in views(0):                                       # warm-up
        fe.accumulate_stats(w, n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync(); e0.record()
    for s in range(sweeps):
        for w, n in views(s):
            fe.accumulate_stats(w, n, stats)
    e1.record(); sync()
    t1 = e0.elapsed_time(e1)
    local_stats = stats.clone()
    sync(); e0.record()
    lasr_b200.cmvn.allreduce_stats(stats)
    e1.record(); sync()
    t_ar = e0.elapsed_time(e1)
    fe2 = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=stats.cpu().numpy())
    Tm = max(int(fe.frame_counts(n)[0].max()) for _, n in pool)
    out = torch.empty((256, Tm, 80), device=dev)
    for w, n in views(0):
        fe2(w, n, max_frames=Tm, out=out)
    sync(); e0.record()
    for s in range(sweeps):
        for w, n in views(s):
            fe2(w, n, max_frames=Tm, out=out)
    e1.record(); sync()
    t2 = e0.elapsed_time(e1)
    red = torch.tensor([t1, t2, t_ar], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    t1, t2, t_ar = (float(x) for x in red.cpu())
    hours = pool_hours * sweeps * world
    line = {"config": "C4 corpus sweep", "n_gpus": world, "audio_hours": hours, "pool_hours_per_gpu": pool_hours, "sweeps": sweeps,
            "pass1_stats_audio_h_per_s": hours / (t1 * 1e-3), "pass2_fbank_global_cmvn_audio_h_per_s": hours / (t2 * 1e-3),
            "allreduce_us": t_ar * 1e3, "frames": float(stats[0, 80])}
    if args.verify:
        # single-process recomputation of every rank's contribution (same seeds), fp64 on the device features
        ref = torch.zeros((2, 81), dtype=torch.float64, device=dev)
        for r in range(world):
            p = make_pool(3 + r, args.pool_batches, dev)
            for s in range(sweeps):
                off = 4 * (s % 16)
                for w, n in p:
                    f, fl = fe(w[:, off:], n)
                    T, _ = fe.frame_counts(n)
                    mask = (torch.arange(f.shape[1], device=dev)[None, :] < torch.from_numpy(T).to(dev)[:, None]).unsqueeze(-1)
                    fd = f.double() * mask
                    ref[0, :80] += fd.sum((0, 1)); ref[1, :80] += (fd * fd).sum((0, 1)); ref[0, 80] += float(T.sum())
        line["verify_max_rel_err"] = float(((stats - ref).abs() / ref.abs().clamp_min(1e-30)).max())
        line["verify_count_equal"] = bool(stats[0, 80] == ref[0, 80])
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
