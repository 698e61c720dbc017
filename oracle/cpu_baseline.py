"""CPU timing of the reference path (TEST / BENCH INFRASTRUCTURE ONLY).

What the reference executes for the front end is, per utterance and inside single-threaded
DataLoader workers (R/bin/train_lighting.py:228, OMP_NUM_THREADS=1 in R/example/asr_en/run.sh:3):
``register_trans["fbank:80"]`` = torchaudio.compliance.kaldi.fbank on the 2^15-scaled waveform
(R/lasr/data/datatrans.py:73-102), then the padding collate ``batch_list`` (R/lasr/data/dataset.py:8-22).
/root/reference does not exist on the GPU box, so the chain is driven through the oracle's
restatement of those wrappers with ``use_torchaudio=True`` -- i.e. the arithmetic runs in the very
library the reference calls (torchaudio 2.11.0 from the image); CMVN (absent from the reference)
uses the oracle's fp64 definition.  Workers mirror the reference's worker model: N processes,
one torch thread each.
"""
import os
import time

import numpy as np


def _init_worker():
    os.environ["OMP_NUM_THREADS"] = "1"
    import torch
    torch.set_num_threads(1)


_WAVS = None     # set before the pool forks: workers inherit the waveforms (a DataLoader worker reads its
                 # own audio from disk; pickling the inputs to the workers would penalise the baseline)


_GLOBAL = None   # (mean, istd) of the global-CMVN chain, inherited by the forked workers like the waveforms
_OPTS = {}       # extra fbank options (e.g. sample_frequency=8000.0) and peak_norm


def _one(args):
    idx, cmvn, specaug = args
    wav = _WAVS[idx]
    from . import lasr_frontend
    opts = dict(_OPTS)
    if opts.pop("peak_norm", False):
        wav = lasr_frontend.voice_norm(wav)                 # vectorised restatement of the reference's pure-Python loop
    x = lasr_frontend.wav_to_kaldi_fbank(wav, use_torchaudio=True, **opts)
    if cmvn == "utt_meanvar":
        x = lasr_frontend.utterance_cmvn(x, True)
    elif cmvn == "utt_mean":
        x = lasr_frontend.utterance_cmvn(x, False)
    elif cmvn == "global":
        x = lasr_frontend.apply_cmvn(x, _GLOBAL[0], _GLOBAL[1])
    if specaug == "full":
        x = lasr_frontend.spec_augment_full(x)              # time warp (PIL BICUBIC) + masks: the registry transform `specaug`
        x = x[0] if isinstance(x, tuple) else x
    elif specaug:
        x, _ = lasr_frontend.spec_augment_masks(x)
    return x


def set_chain_options(global_cmvn=None, **opts):
    """Call BEFORE make_pool: global CMVN vectors (mean, istd) and extra fbank options for the workers."""
    global _GLOBAL, _OPTS
    _GLOBAL, _OPTS = global_cmvn, dict(opts)


def run_chain(wavs, cmvn="utt_meanvar", specaug=False, workers=1, pool=None, first=0):
    """fbank:80 (+ CMVN) (+ SpecAugment masks) for every utterance, then batch_list.  Returns the
    padded (B, Tmax, 80) float32 batch and the frame counts, exactly what collate_fn builds.
    With a pool, ``wavs`` must be the slice ``[first : first + len(wavs)]`` of the list the pool was forked with."""
    from . import lasr_frontend
    global _WAVS
    if pool is None:
        _WAVS = wavs
        jobs = [(i, cmvn, specaug) for i in range(len(wavs))]
    else:
        assert _WAVS is not None and first + len(wavs) <= len(_WAVS), "call make_pool(all_wavs, workers) first"
        jobs = [(first + i, cmvn, specaug) for i in range(len(wavs))]
    if pool is not None:
        feats = pool.map(_one, jobs, chunksize=max(1, len(jobs) // (4 * workers)))
    else:
        feats = [_one(j) for j in jobs]
    return lasr_frontend.batch_list(feats, pad_value=0), np.array([f.shape[0] for f in feats], dtype=np.int64)


def make_pool(wavs, workers):
    """Forks ``workers`` single-threaded processes that inherit ``wavs``."""
    import multiprocessing as mp
    global _WAVS
    _WAVS = wavs
    return mp.get_context("fork").Pool(workers, initializer=_init_worker)


def time_chain(wavs, sample_rate=16000.0, cmvn="utt_meanvar", specaug=False, workers=None, min_seconds=8.0, max_reps=50):
    """Times ``run_chain`` on ``wavs`` with ``workers`` single-threaded processes (default: all host
    cores); repeats until ``min_seconds`` of wall time.  Returns dict(value [audio-h/s], cores,
    seconds, reps, audio_hours_per_rep)."""
    import multiprocessing as mp
    workers = workers or os.cpu_count() or 1
    hours = sum(len(w) for w in wavs) / sample_rate / 3600.0
    pool = None
    if workers > 1:
        pool = make_pool(wavs, workers)
    else:
        _init_worker()
    try:
        run_chain(wavs[: max(2 * workers, 2)], cmvn, specaug, workers, pool)   # warm-up (imports, page-in)
        reps, t0 = 0, time.perf_counter()
        while True:
            run_chain(wavs, cmvn, specaug, workers, pool)
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds or reps >= max_reps:
                break
    finally:
        if pool is not None:
            pool.close()
            pool.join()
    return dict(value=hours * reps / dt, cores=workers, seconds=dt, reps=reps, audio_hours_per_rep=hours)
