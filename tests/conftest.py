import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lasr_b200():
    import lasr_b200 as m
    return m


def tol_violations(got, ref, rtol=1e-4, atol=1e-5):
    """north_star tolerance: |got - ref| <= atol + rtol * |ref| element-wise (fp32)."""
    import numpy as np
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return int((np.abs(got - ref) > atol + rtol * np.abs(ref)).sum())


NOISE_FLOOR_RATIO = 1e-8


def fbank_parity(got, ref32, ref64, lin64, rtol=1e-4, atol=1e-5):
    """Parity of log-mel features against the fp32 reference with the north_star tolerance
    |got - ref32| <= atol + rtol |ref32|.

    fp32 noise floor (measured in SURVEY.md 7.2): any fp32 FFT carries an
    absolute error of ~1e-8 * sqrt(frame energy) per bin, so a mel bin whose energy is more than
    80 dB below its frame's total (ratio < 1e-8; a ~1e-6 fraction of cells on white noise, all
    in mel bins 0-8) differs between ANY two fp32 implementations by more than the tolerance --
    the reference (torchaudio fp32) itself violates the tolerance against its own fp64 evaluation
    there.  Those cells are checked against the fp64 oracle with a 10x wider band instead.

    Returns (hard, soft, below): hard = violations above the noise floor (must be 0), soft =
    cells below the floor outside even the wide band (must be 0), below = number of floor cells.
    """
    import numpy as np
    got = np.asarray(got, dtype=np.float64)
    ref32 = np.asarray(ref32, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    lin64 = np.asarray(lin64, dtype=np.float64)
    ratio = lin64 / np.maximum(lin64.sum(axis=1, keepdims=True), 1e-300)
    below = ratio < NOISE_FLOOR_RATIO
    bad = np.abs(got - ref32) > atol + rtol * np.abs(ref32)
    hard = int((bad & ~below).sum())
    wide = np.abs(got - ref64) > 10 * (atol + rtol * np.abs(ref64))
    soft = int((wide & below).sum())
    return hard, soft, int(below.sum())
