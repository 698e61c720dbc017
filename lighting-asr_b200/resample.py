"""Waveform ingest on the device: resampling and speed perturbation of ragged batches.

Reference: ``ReSample`` = ``librosa.resample(wav, ssr, 16000, res_type="kaiser_fast")`` (lasr/data/datatrans.py:16-20) and
``SoxSpeedPt`` = sox ``speed`` with a ratio drawn by ``numpy.random.choice([1, 1.1, 0.9])`` (datatrans.py:29-39; ``speed``
resamples by 1 / ratio and keeps the nominal rate, so pitch and tempo change together).  librosa / resampy / sox are not in
this image and not vendored by the reference, so their exact filters cannot be pinned.  Two filters are offered:
``res_type="kaiser_fast"`` restates librosa's path (resampy's sinc interpolation with its documented ``kaiser_fast`` window,
re-expressed as a polyphase FIR; checked against the oracle's direct restatement of resampy's loop) and ``res_type="poly"`` is
``scipy.signal.resample_poly``'s default -- a Kaiser (beta = 5) windowed sinc of half length ``10 * max(up, down)`` -- restated
below with numpy and checked against scipy itself in tests/test_gpu_resample.py.  The draws of ``SoxSpeedPt`` are replayed with
the same ``numpy.random.choice`` call, one per utterance, so the global generator ends where the reference leaves it.
"""
import ctypes as C
from fractions import Fraction

import numpy as np
import torch

from . import _lib


def poly_filter(up, down):
    """(h, pre_remove): the zero-padded, gain-corrected FIR of scipy.signal.resample_poly(x, up, down) (window=('kaiser', 5.0),
    padtype='constant') and the number of leading output samples it drops, so that
    ``out[m] = sum_j h[(m + pre_remove) * down - j * up] * x[j]``."""
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    numtaps = 2 * half_len + 1
    m = np.arange(numtaps, dtype=np.float64) - half_len
    h = f_c * np.sinc(f_c * m) * np.kaiser(numtaps, 5.0)       # scipy.signal.firwin(numtaps, f_c, window=('kaiser', 5.0))
    h /= h.sum()                                               # unity gain at DC (firwin scale=True)
    h *= up
    n_pre_pad = down - half_len % down
    pre_remove = (half_len + n_pre_pad) // down
    return np.concatenate([np.zeros(n_pre_pad), h]), pre_remove


def kaiser_fast_filter(up, down, num_zeros=16, precision=9, rolloff=0.85, beta=8.555504641634386):
    """(h, pre_remove, n_valid_fn) for the device polyphase kernel such that it reproduces
    ``librosa.resample(x, sr, sr * up / down, res_type="kaiser_fast")`` = resampy's band-limited sinc interpolation with its
    ``kaiser_fast`` window (16 zero crossings, Kaiser beta 8.5555, roll-off 0.85, 512 table entries per zero crossing, linear
    interpolation between entries; resampy/filters.py, resampy/interpn.py).

    For a rational ratio the interpolation positions repeat with period ``up``: output m sits at input time m * down / up =
    n + r / up, so resampy's two tap loops (left wing x[n - i], right wing x[n + 1 + k], weights read from the table at
    ``offset + i * index_step`` with ``offset = int(frac * num_table)``, ``index_step = int(scale * num_table)`` -- truncations
    included) are a fixed FIR per phase r.  They are laid out as one filter in the kernel's convention
    ``out[m] = sum_j h[(m + pre_remove) * down - j * up] * x[j]``: q = (m * down - j * up) = r + i * up on the left wing,
    r - (k + 1) * up on the right wing, shifted by pre_remove * down.  Phases are exact rationals here; resampy evaluates
    ``t * (1 / ratio)`` in float64, which moves a weight by ~1e-13."""
    num_table = 2 ** precision
    n = num_table * num_zeros
    win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True)) * np.kaiser(2 * n + 1, beta)[n:]
    ratio = up / down
    if ratio < 1:
        win = win * ratio
    delta = np.zeros_like(win)
    delta[:-1] = np.diff(win)
    scale = min(1.0, ratio)
    index_step = int(scale * num_table)
    nwin = win.shape[0]
    max_taps = nwin // index_step + 1
    pre_remove = -(-(max_taps * up) // down)                  # shift S = pre_remove * down >= max right-wing reach
    S = pre_remove * down
    h = np.zeros(S + (max_taps + 1) * up, dtype=np.float64)
    for r in range(up):
        frac = scale * (r / up)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i = np.arange((nwin - offset) // index_step)
        idx = offset + i * index_step
        h[S + r + i * up] = win[idx] + eta * delta[idx]       # left wing: x[n - i]
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k = np.arange((nwin - offset) // index_step)
        idx = offset + k * index_step
        h[S + r - (k + 1) * up] = win[idx] + eta * delta[idx]   # right wing: x[n + 1 + k]
    return h, pre_remove


def out_length(n, up, down):
    n = np.asarray(n, dtype=np.int64)
    return (n * up + down - 1) // down


class Resampler:
    """Polyphase resampling ``src_rate -> dst_rate`` (or an explicit ``up`` / ``down``) of a packed ragged batch on the GPU."""

    def __init__(self, src_rate=None, dst_rate=None, up=None, down=None, res_type="poly"):
        """``res_type``: ``"poly"`` = scipy.signal.resample_poly's filter (pinned against scipy); ``"kaiser_fast"`` = the filter
        and the index arithmetic of ``librosa.resample(..., res_type="kaiser_fast")``, what the reference's ``resample:16k``
        calls (datatrans.py:16-20; restated from resampy's documentation, the library is not in the image)."""
        if res_type not in ("poly", "kaiser_fast"):
            raise ValueError("res_type must be 'poly' or 'kaiser_fast'")
        self.res_type = res_type
        self.rates = (src_rate, dst_rate)
        if up is None:
            fr = Fraction(int(dst_rate), int(src_rate))
            up, down = fr.numerator, fr.denominator
        else:
            fr = Fraction(int(up), int(down))
            up, down = fr.numerator, fr.denominator
        self.up, self.down = up, down
        self.identity = up == 1 and down == 1
        self._tables = {}
        if not self.identity:
            h, self.pre_remove = kaiser_fast_filter(up, down) if res_type == "kaiser_fast" else poly_filter(up, down)
            self.h = h.astype(np.float32)

    def valid_lengths(self, n):
        """Samples the resampler itself produces.  ``kaiser_fast``: resampy yields int(n * ratio) samples (the ratio as the
        float64 quotient of the rates, as resampy computes it) and librosa pads with zeros to ceil(n * ratio) = out_lengths."""
        n = np.asarray(n, dtype=np.int64)
        if self.identity or self.res_type != "kaiser_fast":
            return self.out_lengths(n)
        src, dst = self.rates if self.rates[0] is not None else (self.down, self.up)
        ratio = float(dst) / src
        return np.minimum(np.array([int(int(v) * ratio) for v in n.reshape(-1)], dtype=np.int64).reshape(n.shape), self.out_lengths(n))

    def out_lengths(self, n):
        return np.asarray(n, dtype=np.int64).copy() if self.identity else out_length(n, self.up, self.down)

    def _filter(self, dev):
        key = dev.index or 0
        if key not in self._tables:
            self._tables[key] = torch.from_numpy(self.h).to(dev)
        return self._tables[key]

    @torch.no_grad()
    def __call__(self, wav, lens, offsets=None, out=None, out_offsets=None):
        """``wav``: float32 CUDA, ``(B, Nmax)`` zero padded or 1-D packed with ``offsets`` (elements).  Returns
        ``(packed 1-D CUDA float32, out_lens int64 ndarray, out_offsets int64 ndarray)`` with 16-byte aligned utterance starts --
        the layout ``GpuFbankFrontend.forward(wav, lens, wav_offsets=...)`` reads."""
        if not wav.is_cuda or wav.dtype != torch.float32:
            raise ValueError("Resampler needs a float32 CUDA tensor (there is no CPU path)")
        lens = np.ascontiguousarray(np.asarray(lens, dtype=np.int64).reshape(-1))
        B = len(lens)
        if offsets is None:
            if wav.dim() != 2 or wav.shape[0] != B or wav.stride(1) != 1:
                raise ValueError("wav must be (B, Nmax) with contiguous rows, or 1-D with offsets")
            offsets = np.arange(B, dtype=np.int64) * wav.stride(0)
        offsets = np.ascontiguousarray(np.asarray(offsets, dtype=np.int64).reshape(-1))
        n_out = self.out_lengths(lens)
        if out_offsets is None:
            out_offsets = np.zeros(B, dtype=np.int64)
            np.cumsum((n_out[:-1] + 3) // 4 * 4, out=out_offsets[1:])
        out_offsets = np.ascontiguousarray(np.asarray(out_offsets, dtype=np.int64).reshape(-1))
        total = int((out_offsets + (n_out + 3) // 4 * 4).max()) if B else 0
        dev = wav.device
        if out is None:
            out = torch.zeros((total + 64,), dtype=torch.float32, device=dev)
        elif out.numel() < total or out.dtype != torch.float32 or out.device != dev:
            raise ValueError("out is too small")
        tab = torch.from_numpy(np.stack([offsets, lens, out_offsets, n_out])).to(dev, non_blocking=True)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib = _lib.load()
        p = tab.data_ptr()
        if self.identity:
            tb = torch.from_numpy(np.stack([offsets * 4, out_offsets * 4, lens * 4])).to(dev, non_blocking=True)
            _lib.check(lib.b200fe_copy_ragged(C.c_void_p(wav.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(tb.data_ptr() + 8 * B),
                                              C.c_void_p(tb.data_ptr() + 16 * B), B, int(lens.max()) * 4, stream), "b200fe_copy_ragged")
        else:
            h = self._filter(dev)
            _lib.check(lib.b200fe_resample_poly(C.c_void_p(wav.data_ptr()), C.c_void_p(p), C.c_void_p(p + 8 * B), B, C.c_void_p(out.data_ptr()),
                                                C.c_void_p(p + 16 * B), C.c_void_p(p + 24 * B), int(n_out.max()), C.c_void_p(h.data_ptr()), int(h.numel()),
                                                self.up, self.down, int(self.pre_remove), 1.0, stream), "b200fe_resample_poly")
            if self.res_type == "kaiser_fast":
                # librosa's fix_length: where resampy stops one sample short of ceil(n * ratio), that sample is a zero
                short = np.nonzero(self.valid_lengths(lens) < n_out)[0]
                if len(short):
                    out[torch.from_numpy(out_offsets[short] + n_out[short] - 1).to(dev)] = 0.0
        return out, n_out, out_offsets


class SpeedPerturb:
    """``SoxSpeedPt`` for a batch (datatrans.py:29-39): one ``numpy.random.choice(sp)`` per utterance, in order; ratio r turns
    an utterance of n samples into ceil(n / r) samples (sox ``speed`` = resampling by 1 / r at an unchanged nominal rate)."""

    def __init__(self, sp=(1, 1.1, 0.9)):
        self.sp = [float(r) for r in sp]
        self._rs = {}
        for r in self.sp:
            fr = Fraction(str(r)).limit_denominator(1000)          # 1.1 -> 11/10: output rate / input rate = 10/11
            self._rs[r] = Resampler(up=fr.denominator, down=fr.numerator)

    def draw(self, B):
        return [float(np.random.choice(self.sp)) for _ in range(B)]              # the reference's call, once per utterance

    @torch.no_grad()
    def __call__(self, wav, lens, offsets=None, ratios=None):
        lens = np.asarray(lens, dtype=np.int64).reshape(-1)
        B = len(lens)
        if ratios is None:
            ratios = self.draw(B)
        if offsets is None:
            offsets = np.arange(B, dtype=np.int64) * wav.stride(0)
        offsets = np.asarray(offsets, dtype=np.int64)
        n_out = np.array([int(self._rs[r].out_lengths(n)) for r, n in zip(ratios, lens)], dtype=np.int64)
        out_off = np.zeros(B, dtype=np.int64)
        np.cumsum((n_out[:-1] + 3) // 4 * 4, out=out_off[1:])
        total = int(out_off[-1] + (n_out[-1] + 3) // 4 * 4)
        out = torch.zeros((total + 64,), dtype=torch.float32, device=wav.device)
        for r in self.sp:                                           # one launch per ratio class
            idx = np.array([i for i, q in enumerate(ratios) if q == r], dtype=np.int64)
            if len(idx):
                self._rs[r](wav, lens[idx], offsets[idx], out=out, out_offsets=out_off[idx])
        return out, n_out, out_off, ratios


def avg_channels(wav):
    """``AverageChanl`` (datatrans.py:10-14) on the device: (N, C) float32 CUDA -> (N,) float32, the float64 mean rounded once."""
    if not wav.is_cuda or wav.dtype != torch.float32 or wav.dim() != 2 or not wav.is_contiguous():
        raise ValueError("avg_channels needs a contiguous float32 CUDA tensor (N, C)")
    out = torch.empty((wav.shape[0],), dtype=torch.float32, device=wav.device)
    _lib.check(_lib.load().b200fe_avg_channels(C.c_void_p(wav.data_ptr()), C.c_void_p(out.data_ptr()), int(wav.shape[0]), int(wav.shape[1]),
                                               C.c_void_p(torch.cuda.current_stream(wav.device).cuda_stream)), "b200fe_avg_channels")
    return out
