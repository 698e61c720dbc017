"""numpy restatement of torchaudio.compliance.kaldi.fbank (TEST INFRASTRUCTURE ONLY).

TA:<line> below = torchaudio/compliance/kaldi.py of torchaudio 2.11.0+cu128.
The arithmetic is evaluated op by op in ``dtype`` (float32 to mirror the reference,
float64 as the "truth" used to arbitrate fp32 disagreements).
"""
import math

import numpy as np

EPSILON = float(np.finfo(np.float32).eps)  # TA:21-22


def next_power_of_2(x):  # TA:39-41
    return 1 if x == 0 else 2 ** (x - 1).bit_length()


def window_properties(num_samples, sample_frequency=16000.0, frame_shift=10.0, frame_length=25.0,
                      round_to_power_of_two=True):
    """TA:125-151 -- returns (window_shift, window_size, padded_window_size)."""
    window_shift = int(sample_frequency * frame_shift * 0.001)
    window_size = int(sample_frequency * frame_length * 0.001)
    padded = next_power_of_2(window_size) if round_to_power_of_two else window_size
    assert 2 <= window_size <= num_samples, \
        "choose a window size {} that is [2, {}]".format(window_size, num_samples)  # TA:142
    assert 0 < window_shift
    assert padded % 2 == 0
    return window_shift, window_size, padded


def num_frames(num_samples, window_size=400, window_shift=160):
    """snip_edges=True frame count, TA:63-67."""
    if num_samples < window_size:
        return 0
    return 1 + (num_samples - window_size) // window_shift


def feature_window(window_type, window_size, dtype, blackman_coeff=0.42):
    """TA:86-113.  torch.hann_window(periodic=False) = 0.5 - 0.5 cos(2 pi n/(N-1))."""
    n = np.arange(window_size, dtype=np.float64)
    a = 2.0 * math.pi / (window_size - 1)
    if window_type == "hanning":
        w = 0.5 - 0.5 * np.cos(a * n)
    elif window_type == "hamming":
        w = 0.54 - 0.46 * np.cos(a * n)
    elif window_type == "povey":
        w = (0.5 - 0.5 * np.cos(a * n)).astype(dtype) ** dtype(0.85)  # TA:98-100 (pow in dtype)
    elif window_type == "rectangular":
        w = np.ones(window_size)
    elif window_type == "blackman":
        w = blackman_coeff - 0.5 * np.cos(a * n) + (0.5 - blackman_coeff) * np.cos(2 * a * n)
    else:
        raise Exception("Invalid window type " + window_type)
    return np.asarray(w, dtype=dtype)


def mel_scale(freq):  # TA:326-331
    return 1127.0 * np.log(1.0 + freq / 700.0)


def get_mel_banks(num_bins, window_length_padded, sample_freq, low_freq, high_freq, dtype=np.float32):
    """TA:436-511 with vtln_warp_factor == 1.0 (the only value reachable from LASR).

    torchaudio evaluates this in float32 tensors built from python-float scalars; the same
    order of operations is kept here.  Returns (num_bins, window_length_padded // 2).
    """
    assert num_bins > 3
    assert window_length_padded % 2 == 0
    num_fft_bins = window_length_padded // 2
    nyquist = 0.5 * sample_freq
    if high_freq <= 0.0:
        high_freq += nyquist  # TA:457-458
    assert (0.0 <= low_freq < nyquist) and (0.0 < high_freq <= nyquist) and (low_freq < high_freq)
    fft_bin_width = sample_freq / window_length_padded
    mel_low = 1127.0 * math.log(1.0 + low_freq / 700.0)  # python floats, TA:466-467
    mel_high = 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)  # TA:471
    f = dtype
    b = np.arange(num_bins, dtype=f)[:, None]
    left = f(mel_low) + b * f(delta)  # TA:482-484
    center = f(mel_low) + (b + f(1.0)) * f(delta)
    right = f(mel_low) + (b + f(2.0)) * f(delta)
    mel = (f(1127.0) * np.log(f(1.0) + (f(fft_bin_width) * np.arange(num_fft_bins, dtype=f)) / f(700.0)))[None, :]
    up = (mel - left) / (center - left)  # TA:497-498
    down = (right - mel) / (right - center)
    return np.maximum(f(0.0), np.minimum(up, down)).astype(f)  # TA:502


def frame_signal(x, window_size, window_shift):
    """TA:44-83 with snip_edges=True: (m, window_size) copy of the strided view."""
    m = num_frames(len(x), window_size, window_shift)
    if m == 0:
        return np.empty((0, 0), dtype=x.dtype)
    idx = np.arange(m)[:, None] * window_shift + np.arange(window_size)[None, :]
    return x[idx]


def fbank(waveform, dtype=np.float32, dither=0.0, dither_noise=None, energy_floor=1.0, frame_length=25.0,
          frame_shift=10.0, high_freq=0.0, low_freq=20.0, num_mel_bins=80, preemphasis_coefficient=0.97,
          remove_dc_offset=True, round_to_power_of_two=True, sample_frequency=16000.0, subtract_mean=False,
          use_log_fbank=True, use_power=True, window_type="povey", blackman_coeff=0.42):
    """TA:514-645 for the option subset LASR can reach (snip_edges=True, use_energy=False,
    vtln_warp=1.0; lasr/data/datatrans.py:45-70).

    ``waveform``: 1-D array already scaled to int16 range.  ``dither_noise``: the
    (m, window_size) standard-normal matrix torchaudio would draw (TA:179-181); required
    when ``dither != 0`` because numpy cannot replay torch's generator.
    """
    f = dtype
    x = np.asarray(waveform).astype(f)
    shift, size, padded = window_properties(len(x), sample_frequency, frame_shift, frame_length,
                                            round_to_power_of_two)
    frames = frame_signal(x, size, shift)  # TA:174
    if dither != 0.0:  # TA:179-181
        assert dither_noise is not None and dither_noise.shape == frames.shape
        frames = frames + dither_noise.astype(f) * f(dither)
    if remove_dc_offset:  # TA:183-186
        frames = frames - frames.mean(axis=1, dtype=f, keepdims=True)
    if preemphasis_coefficient != 0.0:  # TA:193-198 (replicate pad on the left)
        prev = np.concatenate([frames[:, :1], frames[:, :-1]], axis=1)
        frames = frames - f(preemphasis_coefficient) * prev
    frames = frames * feature_window(window_type, size, f, blackman_coeff)[None, :]  # TA:201-204
    if padded != size:  # TA:207-211
        frames = np.concatenate([frames, np.zeros((frames.shape[0], padded - size), dtype=f)], axis=1)
    spectrum = np.abs(np.fft.rfft(frames, axis=1))  # TA:616 (complex64 in, float32 hypot out)
    spectrum = spectrum.astype(f)
    if use_power:
        spectrum = spectrum * spectrum  # TA:617-618
    mel = get_mel_banks(num_mel_bins, padded, sample_frequency, low_freq, high_freq, f)
    mel = np.concatenate([mel, np.zeros((num_mel_bins, 1), dtype=f)], axis=1)  # TA:627
    out = (spectrum @ mel.T).astype(f)  # TA:630
    if use_log_fbank:
        out = np.log(np.maximum(out, f(EPSILON)))  # TA:631-633
    if subtract_mean:  # TA:220-226, 644
        out = out - out.mean(axis=0, dtype=f, keepdims=True)
    return out.astype(f)
