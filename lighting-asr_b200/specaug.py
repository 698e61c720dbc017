"""Host side of SpecAugment: replays the reference's RNG consumption and emits rectangles.

The reference (lasr/utils/specaugment.py:47-106, driven by lasr/data/datatrans.py:106-151) draws
mask geometry from TWO global generators -- ``numpy.random`` (legacy MT19937 global state) and
CPython's ``random`` -- in a fixed interleaved order, per utterance, in batch order
(lasr/data/dataset.py:190-198).  To be bit-exact on mask positions for the same seeds the host
code here calls the very same global functions in the same order; only the *application* of the
rectangles (and the running mean fills) happens on the GPU.
"""
import array
import contextlib
import random

import numpy as np

_NO_LOCK = contextlib.nullcontext()

MAX_FREQ_MASKS = 4
MAX_TIME_MASKS = 4


def _draw_time_warp(num_frames, window):
    """Consumes exactly the draws of time_warp (specaugment.py:20-24) without warping."""
    if num_frames - window <= window:
        return None
    center = random.randrange(window, num_frames - window)
    warped = random.randrange(center - window, center + window) + 1
    return center, warped


def plan_utterance(num_frames, num_mel, max_freq_width=27, n_freq_mask=2, max_time_width=40,
                   n_time_mask=2, consume_time_warp_draws=False, max_time_warp=5):
    """Rectangles for ONE utterance, in application order (frequency masks, then time masks).

    Returns (freq, time): int32 arrays [n_freq_mask, 2] / [n_time_mask, 2] of (start, stop),
    already clipped the way numpy slicing clips; skipped or empty masks are (0, 0).
    """
    if consume_time_warp_draws:
        _draw_time_warp(num_frames, max_time_warp)
    freq = np.zeros((n_freq_mask, 2), dtype=np.int32)
    fs = np.random.randint(0, max_freq_width, size=(n_freq_mask, 2))  # specaugment.py:61
    for i, (f, w) in enumerate(fs):
        f0 = random.randrange(0, num_mel - int(f))  # :64 -- drawn before the skip test
        if int(f) == 0:  # :68-69
            continue
        lo, hi = min(f0, num_mel), min(f0 + int(w), num_mel)
        if hi > lo:
            freq[i] = (lo, hi)
    time = np.zeros((n_time_mask, 2), dtype=np.int32)
    ts = np.random.randint(0, max_time_width, size=(n_time_mask, 2))  # :90
    for i, (t, w) in enumerate(ts):
        if num_frames - int(t) <= 0:  # :93-94, no draw
            continue
        t0 = random.randrange(0, num_frames - int(t))  # :95
        if int(t) == 0:  # :98-99
            continue
        lo, hi = min(t0, num_frames), min(t0 + int(w), num_frames)
        if hi > lo:
            time[i] = (lo, hi)
    return freq, time


def plan_batch(frame_lens, num_mel, max_freq_width=27, n_freq_mask=2, max_time_width=40, n_time_mask=2,
               consume_time_warp_draws=False, max_time_warp=5, return_warp=False):
    """Rectangles for a batch, utterances visited in order (dataset.py:190): the same draws, in the same order, from the same
    two global generators as the reference's per-utterance calls -- replayed by ONE call into the C library
    (b200fe_specaug_plan: CPython's ``random.randrange`` and numpy's legacy ``randint`` on their MT19937 states, which are read
    with ``getstate`` / ``get_state`` and written back advanced).  Tested bit for bit, generator positions included, against
    ``plan_batch_reference_loop`` (the same plan drawn with the real Python calls) and against the reference's own functions.
    Returns (masks [B, n_f + n_t, 2] int32, row_bounds [B, 2 n_t] int32 sorted); with ``return_warp`` also
    warps [B, 2] int32 = (center, warped) of the time warp that precedes the masks (specaugment.py:20-24),
    (-1, -1) where the utterance is too short to be warped."""
    import ctypes as C
    from . import _lib
    n_f, n_t = n_freq_mask, n_time_mask
    if n_f > MAX_FREQ_MASKS or n_t > MAX_TIME_MASKS:
        raise ValueError("at most %d frequency and %d time masks are supported" % (MAX_FREQ_MASKS, MAX_TIME_MASKS))
    T = np.ascontiguousarray(np.asarray(frame_lens, dtype=np.int64).reshape(-1))
    B = len(T)
    masks = np.zeros((B, n_f + n_t, 2), dtype=np.int32)
    bounds = np.zeros((B, 2 * n_t), dtype=np.int32)
    warps = np.full((B, 2), -1, dtype=np.int32)
    if B == 0:
        return (masks, bounds, warps) if return_warp else (masks, bounds)
    # CPython's generator: state out as an array('I') (624 key words + position), back in as a tuple
    version, internal, gauss = random.getstate()
    py_state = array.array("I", internal)
    py_addr = py_state.buffer_info()[0]
    # numpy's legacy global generator: its MT19937 state struct {uint32 key[624]; int pos;} is advanced in place
    np_state = np_keep = None
    try:
        bitgen = np.random.mtrand._rand._bit_generator
        if type(bitgen).__name__ != "MT19937":
            raise AttributeError
        np_addr = int(bitgen.ctypes.state_address)
        lock = bitgen.lock
    except AttributeError:                                     # no raw access: copy the state out and back
        np_state = np.random.get_state()
        np_keep = np.concatenate([np.array(np_state[1], dtype=np.uint32), np.array([np_state[2]], dtype=np.uint32)])
        np_addr = np_keep.ctypes.data
        lock = _NO_LOCK
    lib = _lib.load()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    with lock:
        _lib.check(lib.b200fe_specaug_plan(C.c_void_p(py_addr), C.c_void_p(py_addr + 624 * 4), C.c_void_p(np_addr), C.c_void_p(np_addr + 624 * 4),
                                           vp(T), B, int(num_mel), int(max_freq_width), n_f, int(max_time_width), n_t,
                                           int(bool(consume_time_warp_draws)), int(max_time_warp), vp(masks), vp(bounds), vp(warps)),
                   "b200fe_specaug_plan")
    random.setstate((version, tuple(py_state), gauss))
    if np_state is not None:
        np.random.set_state((np_state[0], np_keep[:624], int(np_keep[624])) + tuple(np_state[3:]))
    return (masks, bounds, warps) if return_warp else (masks, bounds)


def plan_batch_reference_loop(frame_lens, num_mel, max_freq_width=27, n_freq_mask=2, max_time_width=40, n_time_mask=2,
                              consume_time_warp_draws=False, max_time_warp=5, return_warp=False):
    """The same plan with one Python call per draw on the real generators (what plan_batch replays in C): kept as the
    executable specification the fast planner is tested against."""
    n_f, n_t = n_freq_mask, n_time_mask
    if n_f > MAX_FREQ_MASKS or n_t > MAX_TIME_MASKS:
        raise ValueError("at most %d frequency and %d time masks are supported" % (MAX_FREQ_MASKS, MAX_TIME_MASKS))
    B = len(frame_lens)
    rr = random.randrange
    # The two generators are independent objects, so only the order WITHIN each stream matters.  The
    # numpy draws are unconditional (specaugment.py:61,90): per utterance 2*n_f values below
    # max_freq_width then 2*n_t values below max_time_width -> one vectorised call with per-element
    # bounds consumes the legacy MT19937 stream exactly like the per-utterance calls (tested).
    high = np.tile(np.array([max_freq_width] * (2 * n_f) + [max_time_width] * (2 * n_t), dtype=np.int64), B)
    draws = np.random.randint(0, high).reshape(B, n_f + n_t, 2).tolist() if B > 0 else []
    rows = []
    warps = []
    for T, dr in zip(frame_lens, draws):
        T = int(T)
        if consume_time_warp_draws and T - max_time_warp > max_time_warp:
            center = rr(max_time_warp, T - max_time_warp)
            warps.append((center, rr(center - max_time_warp, center + max_time_warp) + 1))
        else:
            warps.append((-1, -1))
        row = []
        for f, w in dr[:n_f]:
            f0 = rr(0, num_mel - f)
            if f == 0:
                row += (0, 0)
                continue
            lo = f0 if f0 < num_mel else num_mel
            hi = f0 + w if f0 + w < num_mel else num_mel
            row += (lo, hi) if hi > lo else (0, 0)
        for t, w in dr[n_f:]:
            if T - t <= 0:
                row += (0, 0)
                continue
            t0 = rr(0, T - t)
            if t == 0:
                row += (0, 0)
                continue
            lo = t0 if t0 < T else T
            hi = t0 + w if t0 + w < T else T
            row += (lo, hi) if hi > lo else (0, 0)
        rows.append(row)
    masks = np.asarray(rows, dtype=np.int32).reshape(B, n_f + n_t, 2)
    bounds = np.sort(masks[:, n_f:].reshape(B, 2 * n_t), axis=1)
    if return_warp:
        return masks, np.ascontiguousarray(bounds), np.asarray(warps, dtype=np.int32).reshape(B, 2)
    return masks, np.ascontiguousarray(bounds)
