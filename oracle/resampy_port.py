"""TEST INFRASTRUCTURE (CPU oracle): restatement of ``librosa.resample(y, orig_sr, target_sr, res_type="kaiser_fast")`` as the
reference calls it (R/lasr/data/datatrans.py:16-20).

librosa delegates to ``resampy.resample`` (band-limited sinc interpolation, J. O. Smith's algorithm) with resampy's pre-computed
``kaiser_fast`` filter and then fixes the length to ``ceil(n * ratio)``.  Neither library is in this image or vendored by the
reference, the reference pins no version: **parity unpinned** -- this file restates the published algorithm
(resampy/interpn.py ``_resample_loop``, resampy/core.py ``resample``, resampy/filters.py ``sinc_window``; librosa/core/audio.py
``resample`` with ``fix=True, scale=False``) with the filter parameters resampy documents for ``kaiser_fast``: 16 zero crossings,
Kaiser window beta = 8.555504641634386, roll-off 0.85, 2**9 table entries per zero crossing (``sinc_window``'s default precision;
the table is interpolated linearly, so a different table density would move the result by ~(1/512)^2).
Only tests import this module."""
import numpy as np

KAISER_FAST = dict(num_zeros=16, precision=9, rolloff=0.85, beta=8.555504641634386)


def sinc_window(num_zeros=16, precision=9, rolloff=0.85, beta=8.555504641634386):
    """resampy.filters.sinc_window with window = kaiser(beta): the right wing of the windowed sinc, 2**precision samples per zero
    crossing.  Returns (interp_win, num_table)."""
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = np.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


def resampy_resample(x, sr_orig, sr_new, **filt):
    """resampy.resample(x, sr_orig, sr_new, filter='kaiser_fast') for 1-D float64 input: the loop of resampy/interpn.py, one
    output sample at a time (vectorised over the taps only)."""
    p = dict(KAISER_FAST)
    p.update(filt)
    x = np.asarray(x, dtype=np.float64)
    sample_ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[0] * sample_ratio)
    interp_win, num_table = sinc_window(**p)
    interp_win = interp_win.copy()
    if sample_ratio < 1:
        interp_win *= sample_ratio
    interp_delta = np.zeros_like(interp_win)
    interp_delta[:-1] = np.diff(interp_win)
    scale = min(1.0, sample_ratio)
    time_increment = 1.0 / sample_ratio
    t_out = np.arange(n_out) * time_increment
    index_step = int(scale * num_table)
    nwin = interp_win.shape[0]
    n_orig = x.shape[0]
    y = np.zeros(n_out, dtype=np.float64)
    for t in range(n_out):
        time_register = t_out[t]
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i_max = min(n + 1, (nwin - offset) // index_step)
        if i_max > 0:
            idx = offset + np.arange(i_max) * index_step
            y[t] += np.dot(interp_win[idx] + eta * interp_delta[idx], x[n - np.arange(i_max)])
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_orig - n - 1, (nwin - offset) // index_step)
        if k_max > 0:
            idx = offset + np.arange(k_max) * index_step
            y[t] += np.dot(interp_win[idx] + eta * interp_delta[idx], x[n + 1 + np.arange(k_max)])
    return y


def librosa_resample(y, orig_sr, target_sr, **filt):
    """librosa.resample(y, orig_sr, target_sr, res_type='kaiser_fast') with its defaults fix=True, scale=False: resampy's output
    padded with zeros (or trimmed) to ceil(n * ratio) samples; the same object when the rates are equal."""
    if orig_sr == target_sr:
        return y
    ratio = float(target_sr) / orig_sr
    n_samples = int(np.ceil(np.asarray(y).shape[-1] * ratio))
    y_hat = resampy_resample(y, orig_sr, target_sr, **filt)
    out = np.zeros(n_samples, dtype=np.float64)
    m = min(n_samples, y_hat.shape[0])
    out[:m] = y_hat[:m]
    return out
