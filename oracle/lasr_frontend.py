"""numpy restatement of the LASR-level front-end functions (TEST INFRASTRUCTURE ONLY).

R/ = /root/reference.  Each function names the reference symbol it follows.
"""
import random

import numpy as np

from . import kaldi_fbank


def voice_norm(wav):
    """R/lasr/data/datatrans.py:22-27 (registry key ``norm``): x / (max|x| + 1e-9) in float64.

    The reference walks the samples in a Python list comprehension; this is the same
    arithmetic vectorised (float64 divide, element-wise)."""
    wav = np.asarray(wav, dtype=np.float64)
    peak = float(np.max(np.abs(wav)))
    return wav / (peak + 1e-9)


def wav_to_kaldi_fbank(wav, dtype=np.float32, audio_bit=16, use_torchaudio=False, **opts):
    """R/lasr/data/datatrans.py:42-104 (registry key ``fbank:80``): cast to float32,
    scale by 2**(audio_bit-1) (:73-74), run kaldi fbank with LASR's defaults (:45-70).

    ``use_torchaudio=True`` sends the scaled waveform through the live
    torchaudio.compliance.kaldi.fbank instead of the numpy restatement (used by the CPU
    baseline: that call is exactly what the reference executes)."""
    x = np.asarray(wav).astype(np.float32) * np.float32(2 ** (audio_bit - 1))
    opts.setdefault("num_mel_bins", 80)
    if use_torchaudio:
        import torch
        import torchaudio  # noqa: F401
        from torchaudio.compliance import kaldi
        opts.pop("dither_noise", None)
        return kaldi.fbank(torch.from_numpy(x).unsqueeze(0), **opts).numpy()
    return kaldi_fbank.fbank(x, dtype=dtype, **opts)


def plan_freq_masks(num_mel, F=27, n_mask=2):
    """RNG replay of R/lasr/utils/specaugment.py:60-69: one numpy draw of shape (n,2), then per
    row a ``random.randrange`` for the start (drawn BEFORE the f == 0 skip test).
    Returns [(start, stop)] of the column slices that get overwritten (stop may be < start
    or > num_mel exactly as numpy slicing would receive it)."""
    fs = np.random.randint(0, F, size=(n_mask, 2))
    out = []
    for f, w in fs:
        f0 = random.randrange(0, num_mel - int(f))
        if int(f) == 0:
            continue
        out.append((f0, f0 + int(w)))
    return out


def plan_time_masks(num_frames, T=40, n_mask=2):
    """RNG replay of R/lasr/utils/specaugment.py:89-101: rows with ``len - t <= 0`` are skipped
    WITHOUT drawing; otherwise the start is drawn, then t == 0 skips."""
    ts = np.random.randint(0, T, size=(n_mask, 2))
    out = []
    for t, w in ts:
        if num_frames - int(t) <= 0:
            continue
        t0 = random.randrange(0, num_frames - int(t))
        if int(t) == 0:
            continue
        out.append((t0, t0 + int(w)))
    return out


def _bicubic(x):
    """Pillow's bicubic_filter (src/libImaging/Resample.c), a = -0.5."""
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_rows(src, out_rows):
    """Vertical BICUBIC resize of a float32 (H, W) array exactly as ``Image.fromarray(src).resize((W, out_rows),
    BICUBIC)`` computes it for mode 'F' (Pillow 12.2: precompute_coeffs + ImagingResampleVertical_32bpc:
    float64 coefficients normalised per output row, float64 accumulation in source order, one rounding to
    float32).  Verified bit for bit against Pillow in tests/test_oracle_golden.py."""
    src = np.asarray(src, dtype=np.float32)
    in_size = src.shape[0]
    if out_rows == in_size:
        return src.copy()
    scale = filterscale = in_size / out_rows
    if filterscale < 1.0:
        filterscale = 1.0
    support = 2.0 * filterscale
    out = np.empty((out_rows, src.shape[1]), dtype=np.float32)
    s64 = src.astype(np.float64)
    for xx in range(out_rows):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        k = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax - xmin)]
        ww = 0.0
        for w in k:
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        acc = np.zeros(src.shape[1], dtype=np.float64)
        for y, w in enumerate(k):
            acc = acc + s64[y + xmin] * w
        out[xx] = acc.astype(np.float32)
    return out


def time_warp(x, max_time_warp=5):
    """R/lasr/utils/specaugment.py:4-32 (mode "PIL", in place): no draw and no change when
    t - W <= W; else center ~ randrange(W, t - W), warped ~ randrange(center - W, center + W) + 1 and the
    two segments are resized to ``warped`` and ``t - warped`` rows.  Returns (x, (center, warped) | None)."""
    t = x.shape[0]
    w = max_time_warp
    if t - w <= w:
        return x, None
    center = random.randrange(w, t - w)
    warped = random.randrange(center - w, center + w) + 1
    left = pil_bicubic_rows(x[:center], warped)
    right = pil_bicubic_rows(x[center:], t - warped)
    x[:warped] = left
    x[warped:] = right
    return x, (center, warped)


def spec_augment_full(x, max_time_warp=5, **mask_kw):
    """The registry transform ``specaug`` as the reference runs it (R/lasr/data/datatrans.py:136-150):
    time warp, then frequency masks, then time masks."""
    x, wp = time_warp(x, max_time_warp)
    x, rects = spec_augment_masks(x, **mask_kw)
    return x, wp, rects


def spec_augment_masks(x, max_freq_width=27, n_freq_mask=2, max_time_width=40, n_time_mask=2,
                       replace_with_zero=False):
    """Masks-only part of R/lasr/data/datatrans.py:106-151 (freq_mask then time_mask, in place,
    mean fill taken from the CURRENT array before each overwrite: specaugment.py:71-74,102-105).
    The PIL time-warp (specaugment.py:4-45) is SURVEY.md §8(f) row F1 and is not part of this
    oracle.  Returns (x, rects) with rects = [("f"|"t", start, stop, fill)]."""
    assert x.ndim == 2
    rects = []
    for f0, f1 in plan_freq_masks(x.shape[1], max_freq_width, n_freq_mask):
        fill = np.float32(0) if replace_with_zero else x.mean()
        x[:, f0:f1] = fill
        rects.append(("f", f0, f1, float(fill)))
    for t0, t1 in plan_time_masks(x.shape[0], max_time_width, n_time_mask):
        fill = np.float32(0) if replace_with_zero else x.mean()
        x[t0:t1] = fill
        rects.append(("t", t0, t1, float(fill)))
    return x, rects


def batch_list(array_list, pad_value=0, dtype=np.float32):
    """R/lasr/data/dataset.py:8-22 with pad_dim=0: stack ragged (T_i, ...) arrays into
    (B, Tmax, ...) filled with ``pad_value``."""
    tmax = max(a.shape[0] for a in array_list)
    out = np.full((len(array_list), tmax) + tuple(array_list[0].shape[1:]), pad_value, dtype=dtype)
    for i, a in enumerate(array_list):
        out[i, : a.shape[0]] = a
    return out


def cmvn_stats(feat_list):
    """Kaldi compute-cmvn-stats layout (NOT in the reference -> parity unpinned):
    float64 [2, D+1]; row 0 = column sums + frame count, row 1 = column sums of squares + 0."""
    d = feat_list[0].shape[1]
    stats = np.zeros((2, d + 1), dtype=np.float64)
    for x in feat_list:
        x = x.astype(np.float64)
        stats[0, :d] += x.sum(axis=0)
        stats[1, :d] += (x * x).sum(axis=0)
        stats[0, d] += x.shape[0]
    return stats


def cmvn_from_stats(stats, norm_vars=True):
    """mean = sum/count; var = max(sumsq/count - mean^2, 1e-20); returns (mean, istd) float64."""
    d = stats.shape[1] - 1
    n = stats[0, d]
    mean = stats[0, :d] / n
    if not norm_vars:
        return mean, np.ones(d)
    var = np.maximum(stats[1, :d] / n - mean * mean, 1e-20)
    return mean, 1.0 / np.sqrt(var)


def apply_cmvn(x, mean, istd):
    """(x - mean) * istd evaluated in float32 like the device epilogue."""
    return ((x.astype(np.float32) - mean.astype(np.float32)) * istd.astype(np.float32)).astype(np.float32)


def utterance_cmvn(x, norm_vars=True):
    mean, istd = cmvn_from_stats(cmvn_stats([x]), norm_vars)
    return apply_cmvn(x, mean, istd)


def frontend_chain(wav, peak_norm=False, cmvn="none", global_stats=None, specaug=False, **fbank_opts):
    """norm -> fbank:80 -> CMVN -> specaug masks, the order adopted in SURVEY.md §8(c)."""
    if peak_norm:
        wav = voice_norm(wav)
    x = wav_to_kaldi_fbank(wav, **fbank_opts)
    if cmvn == "utt_mean":
        x = utterance_cmvn(x, norm_vars=False)
    elif cmvn == "utt_meanvar":
        x = utterance_cmvn(x, norm_vars=True)
    elif cmvn == "global":
        mean, istd = cmvn_from_stats(global_stats)
        x = apply_cmvn(x, mean, istd)
    rects = []
    if specaug:
        x, rects = spec_augment_masks(x)
    return x, rects


def make_pad_mask(lengths, max_length=-1):
    """lasr/utils/mask.py:5-45 for the (B, Tmax) case: True where t >= lengths[b], Tmax = max(max(lengths), max_length)."""
    lengths = [int(v) for v in lengths]
    maxlen = max(max(lengths), int(max_length))
    return np.arange(maxlen)[None, :] >= np.asarray(lengths)[:, None]


def src_mask(xlen, max_frames):
    """(~make_pad_mask(xlen.tolist(), max_length=T)).unsqueeze(-2)  (lasr/model/e2e_ctc_att/e2e_base.py:19-20)."""
    return ~make_pad_mask(xlen, max_frames)[:, None, :]


def subsampled_mask(xlen, max_frames):
    """Conv2dSubsampling's returned mask x_mask[:, :, :-2:2][:, :, :-2:2] (subsampling.py:60) and
    hs_len = sum(mask) (E2E_CTC_ATT.subfunction, e2e_base.py:47-49)."""
    m = src_mask(xlen, max_frames)[:, :, :-2:2][:, :, :-2:2]
    return m, m.sum(-1).squeeze(-1)
