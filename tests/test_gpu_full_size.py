"""GPU: BASELINE.json's full-size configurations through size-independent properties, plus oracle parity
on a sample of the utterances."""
import os
import random
import sys

import numpy as np
import pytest
import torch

from conftest import fbank_parity
from oracle import kaldi_fbank, lasr_frontend

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_batch  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_c2_full_batch_properties(lasr_b200):
    """configs[1]: 256 utterances of 1-35 s, fbank + utterance CMVN, padding / length tensors."""
    wav_np, n = make_batch(1)
    wav = torch.from_numpy(wav_np).to(DEV)
    T = np.array([kaldi_fbank.num_frames(int(x)) for x in n])
    raw, rlen = lasr_b200.GpuFbankFrontend()(wav, n)
    cm, clen = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")(wav, n)
    assert raw.shape == cm.shape == (256, int(T.max()), 80)
    assert rlen.cpu().tolist() == T.tolist() == clen.cpu().tolist()
    # padded rows are exactly zero, valid rows are not
    rows = torch.arange(raw.shape[1], device=DEV)[None, :] >= torch.from_numpy(T).to(DEV)[:, None]
    assert float(raw[rows].abs().max()) == 0.0 and float(cm[rows].abs().max()) == 0.0
    assert bool((raw[~rows].abs().sum(-1) > 0).all())
    # utterance CMVN: zero mean / unit variance per column over the valid rows (fp64 check on the device)
    valid = (~rows).unsqueeze(-1).double()
    cnt = valid.sum(1)
    mean = (cm.double() * valid).sum(1) / cnt
    var = ((cm.double() - mean[:, None, :]) ** 2 * valid).sum(1) / cnt
    assert float(mean.abs().max()) < 2e-5
    assert float((var - 1).abs().max()) < 2e-4
    # batch-composition invariance: an utterance processed alone gives the same bits
    for i in (0, 17, 101, 255):
        alone, _ = lasr_b200.GpuFbankFrontend()(wav[i:i + 1, : (int(n[i]) + 3) // 4 * 4].contiguous(), n[i:i + 1])
        assert torch.equal(alone[0], raw[i, : T[i]])
    # oracle parity on a sample of utterances (torchaudio fp32 + fp64 oracle)
    g = raw.cpu().numpy()
    direct = 0
    for i in range(0, 256, 16):
        w = wav_np[i, : n[i]].astype(np.float64)
        ta = lasr_frontend.wav_to_kaldi_fbank(w, use_torchaudio=True)
        r64 = lasr_frontend.wav_to_kaldi_fbank(w, dtype=np.float64)
        lin = lasr_frontend.wav_to_kaldi_fbank(w, dtype=np.float64, use_log_fbank=False)
        hard, soft, _ = fbank_parity(g[i, : T[i]], ta, r64, lin)
        assert hard == 0 and soft == 0
        direct += int((np.abs(g[i, : T[i]] - ta) > 1e-5 + 1e-4 * np.abs(ta)).sum())
    assert direct <= 1e-5 * 16 * int(T.max()) * 80


def test_c3_full_batch_properties(lasr_b200):
    """configs[2]: batch 512 x 10 s, fbank + global CMVN + SpecAugment (2 freq / 2 time masks)."""
    rng = np.random.default_rng(2)
    B, N = 512, 160000
    wav = torch.from_numpy(np.clip(rng.normal(0, 0.1, (B, N)), -1, 1).astype(np.float32)).to(DEV)
    n = np.full(B, N, dtype=np.int64)
    plain = lasr_b200.GpuFbankFrontend()
    raw, _ = plain(wav, n)
    st = plain.accumulate_stats(wav, n)
    # statistics: checksum of checksums against a device fp64 reduction of the features
    ref_sum = raw.double().sum((0, 1))
    ref_sq = (raw.double() ** 2).sum((0, 1))
    assert float(st[0, 80]) == B * 998
    assert torch.allclose(st[0, :80], ref_sum, rtol=1e-7) and torch.allclose(st[1, :80], ref_sq, rtol=1e-7)
    mean, istd = lasr_frontend.cmvn_from_stats(st.cpu().numpy())
    fe = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=st.cpu().numpy(), specaug=True)
    random.seed(2)
    np.random.seed(2)
    out, _ = fe(wav, n)
    masks = fe.last["masks"].cpu().numpy()
    fills = fe.last["fills"].cpu().numpy()
    norm = (raw - torch.from_numpy(mean).to(DEV)) * torch.from_numpy(istd).to(DEV)
    # masks replayed on the host: the same planner draws give the same rectangles; cells outside are untouched
    random.seed(2)
    np.random.seed(2)
    m2, _ = lasr_b200.specaug.plan_batch([998] * B, 80)
    assert np.array_equal(masks, m2)
    o, x = out.cpu().numpy(), norm.cpu().numpy()
    for i in range(0, B, 37):
        masked = np.zeros((998, 80), dtype=bool)
        last = np.full((998, 80), -1)
        for k, (lo, hi) in enumerate(masks[i]):
            if k < 2:
                masked[:, lo:hi] = True
                last[:, lo:hi] = k
            else:
                masked[lo:hi] = True
                last[lo:hi] = k
        assert np.allclose(o[i][~masked], x[i][~masked], rtol=1e-5, atol=1e-5)
        for k in range(4):
            sel = last == k
            if sel.any():
                assert np.all(o[i][sel] == fills[i, k])
        # the oracle (reference semantics: each fill = mean of the current array) on the same CMVN'd features
        y = x[i].copy()
        for k, (lo, hi) in enumerate(masks[i]):
            f = y.mean()
            if k < 2:
                y[:, lo:hi] = f
            else:
                y[lo:hi] = f
            if hi > lo:
                assert abs(float(f) - float(fills[i, k])) <= 2e-5 + 1e-4 * abs(float(f))
    # linearity: halving the waveform shifts every log-mel value by ln(1/4)
    half, _ = plain(wav[:32] * 0.5, n[:32])
    assert float((half - raw[:32] - np.log(0.25)).abs().max()) < 5e-6
