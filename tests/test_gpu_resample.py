"""GPU: waveform ingest on the device -- polyphase resampling (`resample:16k`, datatrans.py:16-20), speed
perturbation (`soxspeed`, datatrans.py:29-39) and channel averaging (`avgchannel`, datatrans.py:10-14).

librosa / resampy / sox are absent from this image and from /root/reference, so the reference's exact resampling filters are
NOT pinned; the oracle is scipy.signal.resample_poly (whose default filter the device code restates), the RNG draws and the
`numpy.average` semantics are the reference's own."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ss = pytest.importorskip("scipy.signal")


def _pack(wavs):
    lens = np.array([len(w) for w in wavs], dtype=np.int64)
    offs = np.zeros(len(wavs), dtype=np.int64)
    np.cumsum((lens[:-1] + 3) // 4 * 4, out=offs[1:])
    buf = np.zeros(int(offs[-1] + lens[-1] + 8), dtype=np.float32)
    for w, o in zip(wavs, offs):
        buf[o:o + len(w)] = w
    return torch.from_numpy(buf).to(DEV), lens, offs


@pytest.mark.parametrize("src,dst", [(16000, 8000), (8000, 16000), (44100, 16000), (48000, 16000), (22050, 16000)])
def test_resampler_matches_scipy_resample_poly(lasr_b200, src, dst):
    rng = np.random.default_rng(src + dst)
    wavs = [rng.uniform(-0.5, 0.5, n).astype(np.float32) for n in (4001, 12345, 700, 30000)]
    rs = lasr_b200.resample.Resampler(src, dst)
    dw, lens, offs = _pack(wavs)
    out, n_out, o_out = rs(dw, lens, offs)
    got = out.cpu().numpy()
    for w, n, o in zip(wavs, n_out, o_out):
        want = ss.resample_poly(w.astype(np.float64), rs.up, rs.down)
        assert n == len(want) and o % 4 == 0
        assert np.abs(got[o:o + n] - want).max() < 2e-6              # float32 taps and accumulation against the float64 oracle
    assert rs.out_lengths(lens).tolist() == n_out.tolist()
    # padded (B, Nmax) input gives the same result
    nmax = int(lens.max())
    pad = np.zeros((len(wavs), nmax), dtype=np.float32)
    for i, w in enumerate(wavs):
        pad[i, : len(w)] = w
    out2, n2, o2 = rs(torch.from_numpy(pad).to(DEV), lens)
    assert torch.equal(out2[: int(o2[-1] + n2[-1])], out[: int(o_out[-1] + n_out[-1])])


@pytest.mark.parametrize("src,dst", [(16000, 8000), (8000, 16000), (44100, 16000), (48000, 16000), (22050, 16000)])
def test_resampler_kaiser_fast_matches_the_librosa_restatement(lasr_b200, src, dst):
    """`resample:16k` as the reference calls it: librosa.resample(wav, ssr, tsr, res_type="kaiser_fast") (datatrans.py:19).  The
    device path (resampy's interpolation re-expressed as a polyphase FIR) against the oracle's direct restatement of resampy's
    tap loops + librosa's length fix (oracle/resampy_port.py; the libraries themselves are absent: parity unpinned)."""
    from oracle import resampy_port
    rng = np.random.default_rng(src + 3 * dst)
    wavs = [rng.uniform(-0.5, 0.5, n).astype(np.float32) for n in (4001, 6173, 441, 37, 2500)]
    rs = lasr_b200.resample.Resampler(src, dst, res_type="kaiser_fast")
    dw, lens, offs = _pack(wavs)
    out, n_out, o_out = rs(dw, lens, offs)
    got = out.cpu().numpy()
    for w, n, o in zip(wavs, n_out, o_out):
        want = resampy_port.librosa_resample(w.astype(np.float64), src, dst)
        assert n == len(want) and o % 4 == 0                         # ceil(n * ratio) samples, librosa's fix_length
        assert np.abs(got[o:o + n] - want).max() < 3e-6              # float32 taps and accumulation against the float64 oracle
    # a tone well inside the pass band keeps its amplitude, one above the new Nyquist frequency is removed (decimation only)
    t = np.arange(20000) / float(src)
    tones = [np.sin(2 * np.pi * 0.2 * min(src, dst) * t).astype(np.float32), np.sin(2 * np.pi * 0.62 * min(src, dst) * t).astype(np.float32)]
    dw, lens, offs = _pack(tones)
    out, n_out, o_out = rs(dw, lens, offs)
    g = out.cpu().numpy()
    mid = lambda i: g[o_out[i] + 300:o_out[i] + n_out[i] - 300]
    # (resampy truncates its table step, int(ratio * 512): at 44.1 -> 16 kHz the pass-band gain is 185.76 / 185 = 1.004, reproduced here)
    assert abs(np.sqrt(2.0 * np.mean(mid(0).astype(np.float64) ** 2)) - 1.0) < 6e-3
    if dst < src:
        assert np.abs(mid(1)).max() < 1e-3


def test_switchboard_path_16k_to_8k_features(lasr_b200):
    """16 kHz audio -> 8 kHz on the device -> the 8 kHz front end (BASELINE config 5's SwitchBoard shape) against live
    torchaudio on the scipy-resampled waveform."""
    from torchaudio.compliance import kaldi
    rng = np.random.default_rng(8)
    wavs = [rng.uniform(-0.5, 0.5, n) for n in (32000, 48001)]
    from oracle import resampy_port
    for res_type in ("poly", "kaiser_fast"):
        col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True, input_rate=16000, sample_frequency=8000.0, res_type=res_type)
        batch = col(wavs)
        f, fl = batch["wav_array"].numpy(), batch["wav_len"].tolist()
        _check_switchboard(wavs, f, fl, (lambda w: ss.resample_poly(w, 1, 2)) if res_type == "poly" else (lambda w: resampy_port.librosa_resample(w, 16000, 8000)))
    # the drop-in default is the reference's own resampler
    assert lasr_b200.lasr_plugin.B200Collate(DEV, input_rate=16000, sample_frequency=8000.0).pipeline.resampler.res_type == "kaiser_fast"


def _check_switchboard(wavs, f, fl, resample):
    from torchaudio.compliance import kaldi
    for i, w in enumerate(wavs):
        w8 = resample(w)
        ref = kaldi.fbank(torch.from_numpy(w8.astype(np.float32) * 32768.0).unsqueeze(0), num_mel_bins=80, dither=0.0, energy_floor=1.0,
                          sample_frequency=8000.0).numpy()
        assert fl[i] == ref.shape[0]
        g = f[i, : fl[i]]
        assert np.mean(np.abs(g - ref) > 1e-4 + 1e-4 * np.abs(ref)) < 1e-3     # float32 resampling noise sits under the fbank tolerance
        assert np.all(f[i, fl[i]:] == 0)


def test_speed_perturbation_replays_the_reference_draws(lasr_b200):
    rng = np.random.default_rng(9)
    wavs = [rng.uniform(-0.5, 0.5, n).astype(np.float32) for n in (8000, 16001, 12000, 9000, 20000, 7777)]
    sp = lasr_b200.resample.SpeedPerturb((1, 1.1, 0.9))
    dw, lens, offs = _pack(wavs)
    np.random.seed(123)
    out, n_out, o_out, ratios = sp(dw, lens, offs)
    after = np.random.randint(0, 1 << 30)
    np.random.seed(123)
    want_ratios = [float(np.random.choice([1, 1.1, 0.9])) for _ in wavs]              # SoxSpeedPt's draw, one per utterance (datatrans.py:31)
    assert ratios == want_ratios and after == np.random.randint(0, 1 << 30)             # the global generator ends where the reference leaves it
    got = out.cpu().numpy()
    for w, r, n, o in zip(wavs, ratios, n_out, o_out):
        if r == 1.0:
            assert n == len(w) and np.array_equal(got[o:o + n], w)
        else:
            up, down = (10, 11) if r == 1.1 else (10, 9)
            want = ss.resample_poly(w.astype(np.float64), up, down)
            assert n == len(want) and np.abs(got[o:o + n] - want).max() < 2e-6
    col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True, speed_perturb=(1, 1.1, 0.9))
    np.random.seed(5)
    b = col([w.astype(np.float64) for w in wavs])
    np.random.seed(5)
    rr = [float(np.random.choice([1, 1.1, 0.9])) for _ in wavs]
    exp = [int(sp._rs[r].out_lengths(len(w))) for r, w in zip(rr, wavs)]
    assert b["wav_len"].tolist() == [1 + (n - 400) // 160 for n in exp]


def test_channel_average(lasr_b200):
    rng = np.random.default_rng(10)
    st = rng.uniform(-1, 1, (5001, 2))
    got = lasr_b200.resample.avg_channels(torch.from_numpy(st.astype(np.float32)).to(DEV)).cpu().numpy()
    assert np.array_equal(got, np.average(st.astype(np.float32).astype(np.float64), axis=1).astype(np.float32))
    # a stereo file in the collate's list goes through the reference's own numpy.average
    col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True)
    mono = np.average(st, axis=1)
    a = col([st, mono])["wav_array"]
    assert torch.equal(a[0], a[1])
