"""Batch formation: the reference's length-sorted, filtered, duration- or size-bounded batches
(``BatchAudioDataSet.check_dataset`` / ``make_batch_size`` / ``make_batch_duration``, lasr/data/dataset.py:260-305) as index
lists, plus the packed layout of every batch (16-byte aligned utterance starts, prefix-sum offsets) that
``GpuFbankFrontend.forward(wav, lens, wav_offsets=...)`` and ``HostPipeline`` read, and the per-rank sharding of batches.

The plan is bit-identical to the reference's for the same ``random.seed`` (it shuffles with the global ``random`` before the
stable sort, dataset.py:262-265, so that ties keep a random order)."""
import random

import numpy as np


def plan_batches(wav_len, token_len, batch_sort=True, batch_size=32, batch_duration=320, batch_type="size", max_duration=30,
                 min_duration=0.3, text_freq=0.08, min_token=0, max_token=5000, shuffle=True):
    """``wav_len`` in seconds, ``token_len`` in tokens (the json fields the reference reads).  Returns a list of index lists."""
    idx = list(range(len(wav_len)))
    if shuffle:
        random.shuffle(idx)                                          # dataset.py:263
    if batch_sort:
        idx.sort(key=lambda i: wav_len[i] * 16000 + token_len[i])    # dataset.py:265 (stable)
    idx = [i for i in idx if wav_len[i] <= max_duration and wav_len[i] >= min_duration and token_len[i] >= min_token and
           token_len[i] <= max_token and wav_len[i] / (token_len[i] + 0.1) > text_freq]          # dataset.py:267-278
    if batch_type == "size":
        return [idx[i:i + batch_size] for i in range(0, len(idx), batch_size)]                # dataset.py:289-290
    if batch_type != "duration":
        return [idx]
    out, bg, acc = [], 0, 0                                          # dataset.py:292-305
    for ed in range(len(idx)):
        acc += wav_len[idx[ed]]
        if acc >= batch_duration:
            out.append(idx[bg:ed + 1])
            bg, acc = ed + 1, 0
    if bg != len(idx):
        out.append(idx[bg:])
    return out


def packed_layout(n_samples, elem_bytes=4):
    """Offsets (elements) of a batch's utterances back to back with 16-byte aligned starts, and the total element count."""
    n = np.asarray(n_samples, dtype=np.int64)
    al = 16 // elem_bytes
    offs = np.zeros(len(n), dtype=np.int64)
    if len(n):
        np.cumsum((n[:-1] + al - 1) // al * al, out=offs[1:])
    total = int(offs[-1] + (n[-1] + al - 1) // al * al) if len(n) else 0
    return offs, total


def shard_batches(batches, durations, world_size):
    """Whole batches to ranks, greedy by descending audio duration: returns one list of batches per rank."""
    load = [0.0] * world_size
    out = [[] for _ in range(world_size)]
    order = sorted(range(len(batches)), key=lambda b: -sum(durations[i] for i in batches[b]))
    for b in order:
        r = min(range(world_size), key=lambda k: load[k])
        out[r].append(batches[b])
        load[r] += sum(durations[i] for i in batches[b])
    return out
