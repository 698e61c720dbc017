"""CPU: the host side of `resample:16k` (lighting-asr_b200/resample.py) -- resampy's `kaiser_fast` interpolation re-expressed as a
polyphase FIR for the device kernel -- against the oracle's direct restatement of resampy's tap loops and librosa's length fix
(oracle/resampy_port.py; librosa / resampy are absent from the image: the restatement follows their published algorithm)."""
import importlib

import numpy as np
import pytest

from oracle import resampy_port


def _kernel_formula(x, h, up, down, pre_remove, n_out):
    """out[m] = sum_j h[(m + pre_remove) * down - j * up] * x[j] in float64 (csrc/resample_kernels.cuh)."""
    y = np.zeros(n_out)
    for m in range(n_out):
        t = (m + pre_remove) * down
        j = np.arange(0, min(t // up, len(x) - 1) + 1)
        q = t - j * up
        ok = (q >= 0) & (q < len(h))
        y[m] = np.dot(h[q[ok]], x[j[ok]])
    return y


@pytest.mark.parametrize("src,dst,n", [(16000, 8000, 2001), (8000, 16000, 777), (44100, 16000, 4411), (48000, 16000, 3000), (22050, 16000, 2500),
                                       (16000, 8000, 37), (44100, 16000, 441), (11025, 16000, 1000)])
def test_kaiser_fast_polyphase_table_equals_the_interpolation_loop(src, dst, n):
    rs = importlib.import_module("lighting-asr_b200.resample")
    x = np.random.default_rng(n).normal(0, 0.3, n)
    want = resampy_port.librosa_resample(x, src, dst)
    r = rs.Resampler(src, dst, res_type="kaiser_fast")
    h, pre = rs.kaiser_fast_filter(r.up, r.down)
    n_out, n_valid = int(r.out_lengths(n)), int(r.valid_lengths(n))
    assert n_out == len(want) == int(np.ceil(n * (float(dst) / src)))
    got = _kernel_formula(x, h, r.up, r.down, pre, n_out)
    got[n_valid:] = 0.0
    assert np.abs(got - want).max() < 1e-11          # exact rational phases against resampy's float64 time register


def test_kaiser_fast_window_properties():
    """16 zero crossings, roll-off 0.85: unity gain at DC for interpolation, the first zero of the sinc at 1 / 0.85 samples."""
    win, num_table = resampy_port.sinc_window()
    assert num_table == 512 and len(win) == 16 * 512 + 1 and abs(win[0] - 0.85) < 1e-12 and abs(win[-1]) < 1e-4
    full = np.concatenate([win[:0:-1], win])[::512]              # the taps of phase 0 at ratio 1
    assert abs(full.sum() - 1.0) < 2e-3
    k = int(round(512 / 0.85))
    assert abs(win[k]) < 2e-3
