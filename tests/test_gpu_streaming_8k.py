"""GPU: the 8 kHz (256-point, two frames per complex FFT) path against the oracle, and BASELINE
config 5: 40 ms chunked streaming equals the offline front end frame for frame."""
import numpy as np
import pytest
import torch

from conftest import fbank_parity
from oracle import kaldi_fbank

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pad(wavs, align=4):
    n = np.array([len(w) for w in wavs], dtype=np.int64)
    nmax = int((n.max() + align - 1) // align * align)
    buf = np.zeros((len(wavs), nmax), dtype=np.float32)
    for i, w in enumerate(wavs):
        buf[i, : len(w)] = w
    return torch.from_numpy(buf).to(DEV), n


def test_8khz_offline_parity(lasr_b200):
    """SwitchBoard shape: sample_frequency=8000 -> window 200 / shift 80 / 256-point FFT / 80 mel bins over 20-4000 Hz."""
    from torchaudio.compliance import kaldi
    rng = np.random.default_rng(4)
    lens = [200, 279, 280, 281, 8000, 8000 * 11 + 17, 2759, 2760, 5 * 8000]
    wavs = [rng.uniform(-0.5, 0.5, n) for n in lens]
    fe = lasr_b200.GpuFbankFrontend(sample_frequency=8000.0)
    wav, n = _pad(wavs)
    feats, flen = fe(wav, n)
    g = feats.cpu().numpy()
    for i, w in enumerate(wavs):
        x = w.astype(np.float32) * np.float32(32768.0)
        ref = kaldi.fbank(torch.from_numpy(x).unsqueeze(0), num_mel_bins=80, dither=0.0, energy_floor=1.0, sample_frequency=8000.0).numpy()
        T = ref.shape[0]
        assert int(flen[i]) == T == kaldi_fbank.num_frames(len(w), 200, 80)
        r64 = kaldi_fbank.fbank(x, dtype=np.float64, sample_frequency=8000.0)
        lin = kaldi_fbank.fbank(x, dtype=np.float64, sample_frequency=8000.0, use_log_fbank=False)
        hard, soft, _ = fbank_parity(g[i, :T], ref, r64, lin)
        assert hard == 0 and soft == 0
        assert np.all(g[i, T:] == 0)
    # other 256-point option sets and the utterance-CMVN / statistics machinery on this path
    w = wavs[5]
    x = torch.from_numpy(w.astype(np.float32) * np.float32(32768.0)).unsqueeze(0)
    for kw in (dict(num_mel_bins=40, window_type="hamming"), dict(num_mel_bins=23, frame_length=20.0, preemphasis_coefficient=0.0)):
        fe2 = lasr_b200.GpuFbankFrontend(sample_frequency=8000.0, **kw)
        got = fe2(*_pad([w]))[0][0].cpu().numpy()
        ref = kaldi.fbank(x, dither=0.0, energy_floor=1.0, sample_frequency=8000.0, **kw).numpy()
        r64 = kaldi_fbank.fbank(x[0].numpy(), dtype=np.float64, sample_frequency=8000.0, **kw)
        kw2 = dict(kw, use_log_fbank=False)
        lin = kaldi_fbank.fbank(x[0].numpy(), dtype=np.float64, sample_frequency=8000.0, **kw2)
        hard, soft, _ = fbank_parity(got, ref, r64, lin)
        assert got.shape == ref.shape and hard == 0 and soft == 0, kw
    from oracle import lasr_frontend
    raw = fe(wav, n)[0].cpu().numpy()
    cm = lasr_b200.GpuFbankFrontend(sample_frequency=8000.0, cmvn="utt_meanvar")(wav, n)[0].cpu().numpy()
    for i in (4, 5, 8):
        T = kaldi_fbank.num_frames(lens[i], 200, 80)
        ref = lasr_frontend.utterance_cmvn(raw[i, :T])
        assert np.allclose(cm[i, :T], ref, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("sf,chunk", [(16000.0, 640), (8000.0, 320)])
def test_streaming_equals_offline(lasr_b200, sf, chunk):
    """40 ms chunks; first push yields 2 frames, steady state 4 frames per push."""
    S, nchunks = 5, 40
    rng = np.random.default_rng(5)
    audio = torch.from_numpy(rng.uniform(-0.5, 0.5, (S, chunk * nchunks)).astype(np.float32)).to(DEV)
    st = lasr_b200.StreamingFbank(S, device=DEV, sample_frequency=sf)
    outs, counts = [], []
    for c in range(nchunks):
        f = st.push(audio[:, c * chunk:(c + 1) * chunk])
        outs.append(f)
        counts.append(f.shape[1])
    assert counts[0] == 2 and all(c == 4 for c in counts[1:])
    online = torch.cat(outs, dim=1)
    off = lasr_b200.GpuFbankFrontend(sample_frequency=sf)(audio, np.full(S, chunk * nchunks, dtype=np.int64))[0]
    assert online.shape[1] == off.shape[1] == kaldi_fbank.num_frames(chunk * nchunks, int(sf * 0.025), int(sf * 0.01))
    assert torch.equal(online, off)
    # irregular chunk sizes
    st.reset()
    pos, outs = 0, []
    for c in (100, 7, 1500, 333, 640, 641, 2000):
        outs.append(st.push(audio[:, pos:pos + c]))
        pos += c
    online = torch.cat(outs, dim=1)
    ref = off[:, : online.shape[1]]
    if sf == 16000.0:
        assert torch.equal(online, ref)      # a frame's arithmetic does not depend on the launch / tile computing it
    else:
        # 8 kHz path: two frames share one complex FFT, so a frame's rounding depends on its partner; irregular
        # chunking changes the pairing -> equal within the fbank tolerance instead of bit for bit
        d = (online - ref).abs()
        tol = 1e-5 + 1e-4 * ref.abs()
        assert float((d > tol).float().mean()) <= 1e-4 and float(d.max()) < 5e-3 and float(d.mean()) < 1e-5
    assert online.shape[1] == kaldi_fbank.num_frames(pos, int(sf * 0.025), int(sf * 0.01))


@pytest.mark.parametrize("sf,chunk", [(16000.0, 640), (8000.0, 320)])
def test_streaming_many_streams_wraps_and_global_cmvn(lasr_b200, sf, chunk):
    """Multi-stream tiles (8 streams x 4 frames per 32-frame tile), a stream count that leaves a partial last tile, the sliding
    window wrapping several times (small max_chunk -> small buffer), and global CMVN applied in the fused launch."""
    S, nchunks = 27, 440
    rng = np.random.default_rng(9)
    audio = torch.from_numpy(rng.uniform(-0.5, 0.5, (S, chunk * nchunks)).astype(np.float32)).to(DEV)
    n = np.full(S, chunk * nchunks, dtype=np.int64)
    stats = lasr_b200.GpuFbankFrontend(sample_frequency=sf).accumulate_stats(audio, n).cpu().numpy()
    for kw in ({}, {"cmvn": "global", "cmvn_stats": stats}):
        st = lasr_b200.StreamingFbank(S, device=DEV, sample_frequency=sf, max_chunk=chunk, **kw)
        assert st.width < chunk * nchunks // 2                      # the buffer is exhausted more than once
        online = torch.cat([st.push(audio[:, c * chunk:(c + 1) * chunk]) for c in range(nchunks)], dim=1)
        off = lasr_b200.GpuFbankFrontend(sample_frequency=sf, **kw)(audio, n)[0]
        assert online.shape == off.shape
        assert torch.equal(online, off)
    # one and two streams (no packing possible / a single partial tile)
    for S1 in (1, 2):
        st = lasr_b200.StreamingFbank(S1, device=DEV, sample_frequency=sf)
        online = torch.cat([st.push(audio[:S1, c * chunk:(c + 1) * chunk]) for c in range(20)], dim=1)
        off = lasr_b200.GpuFbankFrontend(sample_frequency=sf)(audio[:S1, : 20 * chunk].contiguous(), n[:S1] * 0 + 20 * chunk)[0]
        assert torch.equal(online, off)
