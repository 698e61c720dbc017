"""GPU parity: peak normalisation, CMVN (utterance / global) and SpecAugment masks vs the oracle."""
import random

import numpy as np
import pytest
import torch

from conftest import fbank_parity
from oracle import kaldi_fbank, lasr_frontend

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pad(wavs, align=4):
    n = np.array([len(w) for w in wavs], dtype=np.int64)
    nmax = int((n.max() + align - 1) // align * align)
    buf = np.zeros((len(wavs), nmax), dtype=np.float32)
    for i, w in enumerate(wavs):
        buf[i, : len(w)] = w
    return torch.from_numpy(buf).to(DEV), n


def _close(got, ref, rtol=1e-4, atol=1e-5):
    return int((np.abs(got.astype(np.float64) - ref) > atol + rtol * np.abs(ref)).sum())


def test_peak_norm(lasr_b200):
    """VoiceNorm (datatrans.py:22-27) folded into the fused launch: abs-max is exact, the features
    match the oracle chain norm -> fbank:80."""
    rng = np.random.default_rng(5)
    wavs = [rng.uniform(-a, a, n).astype(np.float32).astype(np.float64) for a, n in ((0.5, 16000), (0.05, 23457), (0.9, 8000))]
    fe = lasr_b200.GpuFbankFrontend(peak_norm=True)
    wav, n = _pad(wavs)
    feats, flen = fe(wav, n)
    g = feats.cpu().numpy()
    peak = fe.last["peak"].cpu().numpy()
    for i, w in enumerate(wavs):
        assert peak[i] == np.float32(np.abs(w).max())
        xn = lasr_frontend.voice_norm(w)
        ref = lasr_frontend.wav_to_kaldi_fbank(xn, use_torchaudio=True)
        r64 = lasr_frontend.wav_to_kaldi_fbank(xn, dtype=np.float64)
        lin = lasr_frontend.wav_to_kaldi_fbank(xn, dtype=np.float64, use_log_fbank=False)
        T = ref.shape[0]
        hard, soft, _ = fbank_parity(g[i, :T], ref, r64, lin)
        assert hard == 0 and soft == 0
        assert np.all(g[i, T:] == 0)


def test_utterance_cmvn(lasr_b200):
    """BASELINE config 2 semantics on a small ragged batch: fbank + utterance CMVN (mean, var)."""
    rng = np.random.default_rng(1)
    lens = [400, 4000, 16000, 16000 * 7 + 13, 16000 * 3, 959, 35 * 16000]
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
    wav, n = _pad(wavs)
    raw = lasr_b200.GpuFbankFrontend()(wav, n)[0].cpu().numpy()
    for mode, nv in (("utt_meanvar", True), ("utt_mean", False)):
        fe = lasr_b200.GpuFbankFrontend(cmvn=mode, l2_chunk_bytes=3 << 20)     # forces several utterance groups
        feats, flen = fe(wav, n)
        g = feats.cpu().numpy()
        for i, w in enumerate(wavs):
            T = kaldi_fbank.num_frames(len(w))
            assert int(flen[i]) == T
            # (a) the CMVN stage itself, given the device's own fbank output
            ref = lasr_frontend.utterance_cmvn(raw[i, :T], norm_vars=nv)
            assert _close(g[i, :T], ref.astype(np.float64)) == 0
            assert np.all(g[i, T:] == 0)
            # (b) end to end against the reference fbank: fbank tolerance propagated through the affine map
            ta = lasr_frontend.wav_to_kaldi_fbank(w, use_torchaudio=True)
            mean, istd = lasr_frontend.cmvn_from_stats(lasr_frontend.cmvn_stats([ta]), nv)
            e2e = lasr_frontend.apply_cmvn(ta, mean, istd).astype(np.float64)
            r64 = lasr_frontend.wav_to_kaldi_fbank(w, dtype=np.float64)
            lin = lasr_frontend.wav_to_kaldi_fbank(w, dtype=np.float64, use_log_fbank=False)
            ok = (lin / lin.sum(1, keepdims=True)) >= 1e-8
            tol = (1e-5 + 1e-4 * np.abs(ta)) * istd[None, :] + 2e-5
            if T > 1 or not nv:
                assert int(((np.abs(g[i, :T] - e2e) > tol) & ok).sum()) == 0


def test_cmvn_against_torchaudio_sliding_window_cmn(lasr_b200):
    """Row A11 pinned: the product's utterance CMVN (mean / mean + variance), its statistics accumulation and global CMVN against
    torchaudio.functional.sliding_window_cmn (torchaudio's port of Kaldi's apply-cmvn-sliding) with a window that covers the whole
    matrix: per utterance that is utterance CMVN, on the concatenated corpus it is global CMVN from the corpus' statistics.
    (1) the committed goldens (tests/golden/cmvn_reference.npz: the reference's own fbank:80 outputs through torchaudio),
    (2) live torchaudio on a ragged batch.  Tolerance: the fbank tolerance carried through the affine map (times istd)."""
    import os
    import torchaudio.functional as F
    from torchaudio.compliance import kaldi
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    fb, cm = np.load(os.path.join(gold, "fbank_reference.npz")), np.load(os.path.join(gold, "cmvn_reference.npz"))

    def run(wavs, refs_of):
        wav, n = _pad(wavs)
        st = lasr_b200.GpuFbankFrontend().accumulate_stats(wav, n).cpu().numpy()
        fbs = [refs_of("fbank", i) for i in range(len(wavs))]
        assert st[0, 80] == sum(x.shape[0] for x in fbs)
        gmean, gistd = {}, {}
        for nv in (False, True):
            gmean[nv], gistd[nv] = lasr_frontend.cmvn_from_stats(lasr_frontend.cmvn_stats(fbs), norm_vars=nv)
        checked = floor_cells = cells = 0
        for mode, nv, key in (("utt_mean", False, "utt_mean"), ("utt_meanvar", True, "utt_meanvar"), ("global", False, "global_mean"),
                              ("global", True, "global_meanvar")):
            fe = lasr_b200.GpuFbankFrontend(cmvn=mode, cmvn_stats=st if mode == "global" else None)
            if mode == "global":
                fe.set_global_cmvn(st, norm_vars=nv)            # the product's own accumulated statistics
            g = fe(wav, n)[0].cpu().numpy()
            for i, x in enumerate(fbs):
                ref = refs_of(key, i)
                if ref is None:
                    continue
                T = x.shape[0]
                if mode == "global":
                    istd = gistd[nv]
                else:
                    istd = lasr_frontend.cmvn_from_stats(lasr_frontend.cmvn_stats([x]), norm_vars=nv)[1]
                # cells in the fp32 noise floor of the FFT (conftest.fbank_parity; ~1e-6 of the cells of a white signal) are exempt
                lin = lasr_frontend.wav_to_kaldi_fbank(wavs[i], dtype=np.float64, use_log_fbank=False)
                ok = (lin / lin.sum(1, keepdims=True)) >= 1e-8
                tol = (1e-5 + 1e-4 * np.abs(x)) * istd[None, :] * 2.0 + 2e-5
                assert int(((np.abs(g[i, :T] - ref) > tol) & ok).sum()) == 0, (mode, nv, i)
                floor_cells += int((~ok).sum())
                cells += ok.size
                assert np.all(g[i, T:] == 0)
                checked += 1
        assert floor_cells <= 4 + cells // 10000          # the exemption stays a handful of cells
        return checked

    # (1) committed goldens
    wavs = [fb["wav_%d" % i] for i in range(6)]
    n1 = run(wavs, lambda key, i: (fb["fbank_%d" % i] if key == "fbank" else (cm["%s_%d" % (key, i)] if "%s_%d" % (key, i) in cm.files else None)))
    assert n1 >= 17
    # (2) live torchaudio on a ragged batch of broadband signals
    rng = np.random.default_rng(5)
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in (16000, 16000 * 4 + 7, 8000, 16000 * 11, 1200)]
    fbs = [kaldi.fbank(torch.from_numpy(w).float()[None] * 32768.0, num_mel_bins=80, dither=0.0, energy_floor=1.0).numpy() for w in wavs]
    bounds = np.cumsum([0] + [x.shape[0] for x in fbs])

    def cmn(x, nv):
        t = torch.from_numpy(np.asarray(x, dtype=np.float64))[None]
        return F.sliding_window_cmn(t, cmn_window=2 * t.shape[1] + 1000, min_cmn_window=1, center=True, norm_vars=nv)[0].numpy()

    cat = {nv: cmn(np.concatenate(fbs), nv) for nv in (False, True)}

    def live(key, i):
        if key == "fbank":
            return fbs[i]
        if key.startswith("utt"):
            return cmn(fbs[i], key == "utt_meanvar")
        return cat[key == "global_meanvar"][bounds[i]:bounds[i + 1]]

    assert run(wavs, live) >= 15


def _run_specaug(lasr_b200, wavs, seed, replace_with_zero, cmvn_stats):
    wav, n = _pad(wavs)
    fe = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=cmvn_stats, specaug=True, replace_with_zero=replace_with_zero)
    random.seed(seed)
    np.random.seed(seed)
    feats, flen = fe(wav, n)
    return fe, feats.cpu().numpy()


def test_global_cmvn_and_specaug_masks(lasr_b200):
    """BASELINE config 3 semantics (small batch): fbank + global CMVN + 2 freq / 2 time masks.
    Mask POSITIONS are bit-exact for the same seeds (specaugment.py:47-106); fills within tolerance."""
    rng = np.random.default_rng(2)
    lens = [160000, 160000, 48000, 7 * 16000 + 77, 6000, 1200]
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
    wav, n = _pad(wavs)
    plain = lasr_b200.GpuFbankFrontend()
    raw = plain(wav, n)[0].cpu().numpy()
    T = [kaldi_fbank.num_frames(x) for x in lens]
    # statistics pass (no feature output) vs the fp64 definition on the device's own features
    st = plain.accumulate_stats(wav, n).cpu().numpy()
    ref_st = lasr_frontend.cmvn_stats([raw[i, : T[i]] for i in range(len(lens))])
    assert st.shape == (2, 81) and st[0, 80] == sum(T)
    assert np.allclose(st, ref_st, rtol=2e-7, atol=1e-4)      # fp32 pivoted partial sums per tile part, fp64 across
    mean, istd = lasr_frontend.cmvn_from_stats(ref_st)
    for zero in (False, True):
        fe, g = _run_specaug(lasr_b200, wavs, 11, zero, st)
        random.seed(11)
        np.random.seed(11)
        for i in range(len(lens)):
            x = lasr_frontend.apply_cmvn(raw[i, : T[i]], mean, istd)
            base = x.copy()
            y, rects = lasr_frontend.spec_augment_masks(x, replace_with_zero=zero)   # oracle, same RNG streams, same order
            gi = g[i, : T[i]]
            masked = np.zeros_like(base, dtype=bool)
            for kind, lo, hi, _ in rects:
                if kind == "f":
                    masked[:, max(lo, 0):max(hi, 0)] = True
                else:
                    masked[max(lo, 0):max(hi, 0)] = True
            # un-masked cells are the CMVN'd features, masked cells carry the oracle's value
            assert _close(gi[~masked], base[~masked].astype(np.float64)) == 0
            if zero:
                assert np.all(gi[masked] == 0)
            else:
                assert _close(gi[masked], y[masked].astype(np.float64), rtol=1e-4, atol=2e-5) == 0
            # positions: exactly the planned rectangles changed
            assert np.array_equal(masked, lasr_b200_mask(fe, i, T[i]))
            assert np.all(g[i, T[i]:] == 0)


def lasr_b200_mask(fe, i, T):
    m = fe.last["masks"].cpu().numpy()[i]
    out = np.zeros((T, 80), dtype=bool)
    for lo, hi in m[:2]:
        out[:, lo:hi] = True
    for lo, hi in m[2:]:
        out[lo:hi] = True
    return out


def test_specaug_without_cmvn_mean_fill(lasr_b200):
    """specaug directly on log-mel (LASR's own chain norm -> fbank:80 -> specaug has no CMVN)."""
    rng = np.random.default_rng(9)
    wavs = [rng.uniform(-0.5, 0.5, n) for n in (32000, 16000 * 4 + 5, 3000)]
    wav, n = _pad(wavs)
    raw = lasr_b200.GpuFbankFrontend()(wav, n)[0].cpu().numpy()
    fe = lasr_b200.GpuFbankFrontend(specaug=True)
    random.seed(4)
    np.random.seed(4)
    g = fe(wav, n)[0].cpu().numpy()
    random.seed(4)
    np.random.seed(4)
    for i, w in enumerate(wavs):
        T = kaldi_fbank.num_frames(len(w))
        y, _ = lasr_frontend.spec_augment_masks(raw[i, :T].copy())
        assert _close(g[i, :T], y.astype(np.float64), rtol=1e-4, atol=2e-5) == 0


def test_full_specaug_with_time_warp(lasr_b200):
    """Row F1: warp + masks as the reference's registry transform `specaug` (datatrans.py:136-150).  The warp
    repeats Pillow's float64 BICUBIC arithmetic, so the warped features are bit-identical to the oracle's
    (itself pinned bit for bit to the reference); masks follow with the usual fill tolerance."""
    rng = np.random.default_rng(13)
    lens = [160000, 48000, 7 * 16000 + 77, 2000, 1840, 400 + 160 * 10]      # incl. T = 11 (no warp) and T = 10
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
    wav, n = _pad(wavs)
    T = [kaldi_fbank.num_frames(x) for x in lens]
    for kw in ({}, {"cmvn": "utt_meanvar"}, {"replace_with_zero": True}):
        base_kw = {k: v for k, v in kw.items() if k != "replace_with_zero"}
        pre = lasr_b200.GpuFbankFrontend(**base_kw)(wav, n)[0].cpu().numpy()
        fe = lasr_b200.GpuFbankFrontend(specaug=True, time_warp=True, **kw)
        random.seed(21)
        np.random.seed(21)
        g = fe(wav, n)[0].cpu().numpy()
        warps = fe.last["warp"].cpu().numpy()
        random.seed(21)
        np.random.seed(21)
        for i in range(len(lens)):
            x = pre[i, : T[i]].copy()
            xw, wp = lasr_frontend.time_warp(x)
            warped_only = xw.copy()
            y, rects = lasr_frontend.spec_augment_masks(xw, replace_with_zero=kw.get("replace_with_zero", False))
            assert (wp is None and warps[i, 0] < 0) or tuple(warps[i]) == wp
            masked = np.zeros_like(y, dtype=bool)
            for kind, lo, hi, _ in rects:
                if kind == "f":
                    masked[:, max(lo, 0):max(hi, 0)] = True
                else:
                    masked[max(lo, 0):max(hi, 0)] = True
            gi = g[i, : T[i]]
            assert np.array_equal(gi[~masked], warped_only[~masked])          # bit-identical warp
            assert _close(gi[masked], y[masked].astype(np.float64), rtol=1e-4, atol=2e-5) == 0
            assert np.all(g[i, T[i]:] == 0)
        # both generators end where the reference's full transform leaves them
        after = (random.random(), np.random.rand())
        random.seed(21)
        np.random.seed(21)
        for i in range(len(lens)):
            lasr_frontend.spec_augment_full(pre[i, : T[i]].copy())
        assert after == (random.random(), np.random.rand())


@pytest.mark.gpu
def test_time_warp_masks_in_the_warp_launch_match_the_separate_launches(lasr_b200):
    """b200fe_time_warp with d_masks (finalize + fill by the CTA that completes an utterance) against the three-launch
    chain warp -> finalize -> post pass: same cells, same fills; twice in a row (the completion counters must end at zero)."""
    rng = np.random.default_rng(5)
    lens = [int(x) for x in rng.integers(400 + 160 * 3, 16000 * 12, size=37)] + [16000 * 30, 2000]
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
    wav, n = _pad(wavs)
    for kw in ({}, {"replace_with_zero": True}):
        outs = []
        for fused in (True, False, True):
            fe = lasr_b200.GpuFbankFrontend(specaug=True, time_warp=True, **kw)
            fe.fuse_warp_masks = fused
            for rep in range(2):
                random.seed(3)
                np.random.seed(3)
                g = fe(wav, n)[0]
            outs.append((g.cpu().numpy(), fe.last["fills"].cpu().numpy(), fe.launch_count))
        assert np.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-7)
        same_fill = np.all(outs[0][1] == outs[1][1], axis=1)
        assert same_fill.mean() > 0.9                       # fp64 atomics in another order may move a fill by one float32 ulp
        assert np.array_equal(outs[0][0][same_fill], outs[1][0][same_fill])
        assert np.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=1e-6)
        assert np.allclose(outs[0][0], outs[2][0], rtol=1e-6, atol=1e-6)
        assert outs[0][2] < outs[1][2]                      # two launches fewer per call


def test_mean_fills_inside_the_fused_launch_match_the_post_pass(lasr_b200):
    """apply_cmvn_mode 3 (completion tiles: fills + masked cells written by the fused launch itself) against finalize + post pass:
    the same cells, the same fills, with and without global CMVN; no completion tile ever gave up waiting."""
    rng = np.random.default_rng(8)
    lens = [int(x) for x in rng.integers(400, 16000 * 14, size=70)] + [16000 * 33, 400, 560, 2000]
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
    wav, n = _pad(wavs)
    st = lasr_b200.GpuFbankFrontend().accumulate_stats(wav, n).cpu().numpy()
    for kw in ({}, {"cmvn": "global", "cmvn_stats": st}):
        outs = []
        for inlaunch in (True, False):
            fe = lasr_b200.GpuFbankFrontend(specaug=True, **kw)
            fe.inlaunch_fills = inlaunch
            fe.fill_lag = 3 if inlaunch else 64               # a short lag makes some completion tiles wait for their frame tiles
            for rep in range(2):
                random.seed(17)
                np.random.seed(17)
                g = fe(wav, n)[0]
            flags = fe.last["apply_flags"]
            if inlaunch:
                assert flags is not None and int(flags[0][flags[1]].item()) == 0
            outs.append((g.cpu().numpy(), fe.last["fills"].cpu().numpy(), fe.launch_count))
        assert np.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-7)
        same = np.all(outs[0][1] == outs[1][1], axis=1)
        assert same.mean() > 0.9
        assert np.array_equal(outs[0][0][same], outs[1][0][same])
        assert np.allclose(outs[0][0], outs[1][0], rtol=1e-6, atol=1e-6)
        assert outs[0][2] < outs[1][2]


@pytest.mark.parametrize("nmel", [40, 30])
def test_time_warp_other_bin_counts(lasr_b200, nmel):
    """The warp launch with 40 bins (ten float4 per row, 25 row slots) and with 30 bins (not a multiple of four: the scalar
    path, no staged window, generic statistics and fill loops), masks inside the launch and as separate launches."""
    rng = np.random.default_rng(nmel)
    lens = [16000 * 9 + 123, 16000 * 2, 5000, 16000 * 5 + 7, 400 + 160 * 11, 400 + 160 * 64]
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
    wav, n = _pad(wavs)
    T = [kaldi_fbank.num_frames(x) for x in lens]
    pre = lasr_b200.GpuFbankFrontend(num_mel_bins=nmel)(wav, n)[0].cpu().numpy()
    for fused in (True, False):
        fe = lasr_b200.GpuFbankFrontend(num_mel_bins=nmel, specaug=True, time_warp=True)
        fe.fuse_warp_masks = fused
        random.seed(31)
        np.random.seed(31)
        g = fe(wav, n)[0].cpu().numpy()
        random.seed(31)
        np.random.seed(31)
        for i in range(len(lens)):
            xw, _ = lasr_frontend.time_warp(pre[i, : T[i]].copy())
            warped_only = xw.copy()
            y, rects = lasr_frontend.spec_augment_masks(xw)
            masked = np.zeros_like(y, dtype=bool)
            for kind, lo, hi, _ in rects:
                if kind == "f":
                    masked[:, max(lo, 0):max(hi, 0)] = True
                else:
                    masked[max(lo, 0):max(hi, 0)] = True
            gi = g[i, : T[i]]
            assert np.array_equal(gi[~masked], warped_only[~masked])
            assert _close(gi[masked], y[masked].astype(np.float64), rtol=1e-4, atol=2e-5) == 0
            assert np.all(g[i, T[i]:] == 0)
