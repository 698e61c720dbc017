#!/usr/bin/env python
"""Global CMVN statistics of a corpus (Kaldi `compute-cmvn-stats` semantics, the file `apply-cmvn` reads).

    python tools/compute_cmvn_stats.py --list wavs.txt --out cmvn.stats [--batch-seconds 2000] [--peak-norm]
    python -m torch.distributed.run --nproc-per-node N ... tools/compute_cmvn_stats.py --list wavs.txt --out cmvn.stats

``wavs.txt``: one path per line (optionally ``id path`` like a Kaldi wav.scp): ``.npy`` (1-D float or int16 array) or PCM-16
``.wav`` (read with the standard library; the reference reads audio with soundfile, R/lasr/data/reader.py:15-24, absent here).
Utterances are sharded per rank by length (cmvn.shard_utterances), every rank accumulates [sum | count ; sumsq | 0] with the
fused kernel in statistics-only mode (no feature output), ONE all-reduce of the 2 x 81 float64 matrix follows and
rank 0 writes the Kaldi text matrix (lasr_b200.cmvn.save_stats).  Use it with GpuFbankFrontend(cmvn="global", cmvn_stats=...)."""
import argparse
import os
import sys
import wave

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def read_audio(path):
    if path.endswith(".npy"):
        a = np.load(path)
    elif path.endswith(".wav"):
        with wave.open(path, "rb") as w:
            if w.getsampwidth() != 2:
                raise ValueError("only PCM-16 .wav files are read here: " + path)
            a = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
            if w.getnchannels() > 1:
                a = np.average(a.reshape(-1, w.getnchannels()).astype(np.float64) / 32768.0, axis=1)     # avgchannel, datatrans.py:10-14
    else:
        raise ValueError("unknown audio type: " + path)
    if a.dtype == np.int16:
        return a
    return np.asarray(a, dtype=np.float32).reshape(-1)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--list", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--batch-seconds", type=float, default=2000.0)
    ap.add_argument("--sample-rate", type=float, default=16000.0)
    ap.add_argument("--peak-norm", action="store_true", help="apply VoiceNorm first (audio_trans: [norm, fbank:80])")
    args = ap.parse_args(argv)
    import torch.distributed as dist
    import lasr_b200
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    paths = [ln.split()[-1] for ln in open(args.list).read().splitlines() if ln.strip()]
    sizes = np.array([os.path.getsize(p) for p in paths], dtype=np.int64)          # shard by file size: no audio is read twice
    mine = lasr_b200.cmvn.shard_utterances(sizes, world)[rank] if world > 1 else np.arange(len(paths))
    fe = lasr_b200.GpuFbankFrontend(sample_frequency=args.sample_rate, peak_norm=args.peak_norm)
    stats = torch.zeros((2, fe.num_mel_bins + 1), dtype=torch.float64, device=dev)
    win = int(args.sample_rate * 0.025)
    batch, acc = [], 0.0

    def flush():
        nonlocal batch, acc
        if not batch:
            return
        dt = torch.int16 if batch[0].dtype == np.int16 else torch.float32
        n = np.array([len(w) for w in batch], dtype=np.int64)
        buf = torch.zeros((len(batch), int((n.max() + 7) // 8 * 8)), dtype=dt)
        for i, w in enumerate(batch):
            buf[i, : len(w)] = torch.from_numpy(np.ascontiguousarray(w))
        fe.accumulate_stats(buf.to(dev), n, stats)
        batch, acc = [], 0.0

    skipped = 0
    for i in mine:
        w = read_audio(paths[int(i)])
        if len(w) < win:
            skipped += 1                        # shorter than one window: no frames (LASR filters min_duration, dataset.py:272)
            continue
        if batch and batch[0].dtype != w.dtype:
            flush()
        batch.append(w)
        acc += len(w) / args.sample_rate
        if acc >= args.batch_seconds:
            flush()
    flush()
    lasr_b200.cmvn.allreduce_stats(stats)
    if rank == 0:
        lasr_b200.cmvn.save_stats(args.out, stats)
        print("frames %d  utterances %d (skipped %d on rank 0)  -> %s" % (int(stats[0, -1]), len(paths), skipped, args.out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
