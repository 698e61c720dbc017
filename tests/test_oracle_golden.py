"""CPU: the oracle restatement against fixtures produced by the unmodified reference
(oracle/gen_golden.py: lasr.data.datatrans / lasr.utils.specaugment / batch_list + torchaudio 2.11.0)."""
import os
import random

import numpy as np
import pytest

from conftest import fbank_parity
from oracle import kaldi_fbank, lasr_frontend

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def fb():
    return np.load(os.path.join(GOLD, "fbank_reference.npz"))


@pytest.fixture(scope="module")
def sg():
    return np.load(os.path.join(GOLD, "specaug_reference.npz"))


def test_frame_counts_known_answers():
    # N=160000 -> 998, 1 s -> 98, 35 s -> 3498 ; TA:63-67
    assert kaldi_fbank.num_frames(160000) == 998
    assert kaldi_fbank.num_frames(16000) == 98
    assert kaldi_fbank.num_frames(35 * 16000) == 3498
    assert kaldi_fbank.num_frames(399) == 0 and kaldi_fbank.num_frames(400) == 1
    assert kaldi_fbank.num_frames(559) == 1 and kaldi_fbank.num_frames(560) == 2
    assert kaldi_fbank.window_properties(16000) == (160, 400, 512)
    assert kaldi_fbank.window_properties(8000, sample_frequency=8000.0) == (80, 200, 256)
    with pytest.raises(AssertionError):
        kaldi_fbank.window_properties(399)


def test_fbank_matches_reference_fixture(fb):
    for i in range(6):
        w = fb["wav_%d" % i]
        ref = fb["fbank_%d" % i]
        got = lasr_frontend.wav_to_kaldi_fbank(w)
        r64 = lasr_frontend.wav_to_kaldi_fbank(w, dtype=np.float64)
        lin = lasr_frontend.wav_to_kaldi_fbank(w, dtype=np.float64, use_log_fbank=False)
        assert got.shape == ref.shape and got.dtype == np.float32
        hard, soft, _ = fbank_parity(got, ref, r64, lin)
        assert hard == 0 and soft == 0
        hard, soft, _ = fbank_parity(r64, ref, r64, lin)      # the fp64 oracle agrees with the reference too
        assert hard == 0 and soft == 0


def test_silence_is_log_eps():
    out = lasr_frontend.wav_to_kaldi_fbank(np.zeros(4000))
    assert np.all(out == np.float32(np.log(np.float32(kaldi_fbank.EPSILON))))
    assert abs(float(out[0, 0]) + 15.942385) < 1e-6


def test_voice_norm_bit_exact(fb):
    for i in range(6):
        assert np.array_equal(lasr_frontend.voice_norm(fb["wav_%d" % i]), fb["norm_%d" % i])


def test_norm_then_fbank(fb):
    for i in range(6):
        w = fb["wav_%d" % i]
        got = lasr_frontend.wav_to_kaldi_fbank(lasr_frontend.voice_norm(w))
        r64 = lasr_frontend.wav_to_kaldi_fbank(lasr_frontend.voice_norm(w), dtype=np.float64)
        lin = lasr_frontend.wav_to_kaldi_fbank(lasr_frontend.voice_norm(w), dtype=np.float64, use_log_fbank=False)
        hard, soft, _ = fbank_parity(got, fb["norm_fbank_%d" % i], r64, lin)
        assert hard == 0 and soft == 0


def test_subtract_mean_pins_utterance_mean_normalisation(fb):
    """fbank(subtract_mean=True) (TA:220-226) is the only CMVN-like op on the reference's path; the
    oracle's utterance mean normalisation must reproduce it."""
    for i in range(6):
        raw = fb["fbank_%d" % i]
        got = lasr_frontend.utterance_cmvn(raw, norm_vars=False)
        assert np.allclose(got, fb["fbank_cms_%d" % i], rtol=0, atol=2e-5)


def test_batch_list(fb):
    got = lasr_frontend.batch_list([fb["fbank_%d" % i] for i in range(6)], pad_value=0)
    assert np.array_equal(got, fb["batch_list"])


def _pattern(T):
    from oracle.gen_golden import pattern
    return pattern(T)


def test_specaug_masks_match_reference(sg):
    for key in sg["cases"]:
        seed, T = int(key.split("_")[0][1:]), int(key.split("_T")[1])
        random.seed(seed)
        np.random.seed(seed)
        x = _pattern(T)
        y, rects = lasr_frontend.spec_augment_masks(x.copy())
        assert np.array_equal(y, sg[key + "_out"]), key          # positions AND fills, bit for bit
        assert np.array_equal(np.array([random.random(), np.random.rand()]), sg[key + "_rng_after"]), key
        random.seed(seed)
        np.random.seed(seed)
        z, _ = lasr_frontend.spec_augment_masks(x.copy(), replace_with_zero=True)
        assert np.array_equal(np.packbits(z != x), sg[key + "_zero_changed"]), key


def test_pil_bicubic_restatement_is_bit_exact(sg):
    for key in sg["pil_cases"]:
        h, o = (int(v) for v in key.split("_"))
        got = lasr_frontend.pil_bicubic_rows(sg["pil_in_" + key], o)
        assert np.array_equal(got, sg["pil_out_" + key]), key


def test_full_specaug_with_time_warp_matches_reference(sg):
    """Row F1: the registry transform ``specaug`` (warp + masks) as the reference runs it."""
    for key in sg["full_cases"]:
        seed, T = int(key.split("_")[0][1:]), int(key.split("_T")[1])
        random.seed(seed)
        np.random.seed(seed)
        y, wp, _ = lasr_frontend.spec_augment_full(_pattern(T))
        assert np.array_equal(y, sg["full_" + key]), key
        assert (wp is None) == (T - 5 <= 5)


def test_cmvn_definition():
    rng = np.random.default_rng(0)
    feats = [rng.normal(3, 2, (T, 80)).astype(np.float32) for T in (10, 57, 300)]
    st = lasr_frontend.cmvn_stats(feats)
    assert st.shape == (2, 81) and st[0, 80] == 367 and st[1, 80] == 0
    mean, istd = lasr_frontend.cmvn_from_stats(st)
    cat = np.concatenate(feats).astype(np.float64)
    assert np.allclose(mean, cat.mean(0)) and np.allclose(istd, 1 / cat.std(0))
    y = lasr_frontend.utterance_cmvn(feats[2])
    assert np.allclose(y.mean(0), 0, atol=1e-5) and np.allclose(y.std(0), 1, atol=1e-4)


def test_fbank_option_sets_match_reference_fixture():
    """The other option sets of WavToKaldiFbank (datatrans.py:43-71; the 8 kHz family of BASELINE config 5 among them) against the
    UNMODIFIED reference called with keyword arguments (oracle/gen_golden.py --options)."""
    from oracle.gen_golden import OPTION_SETS
    g = np.load(os.path.join(GOLD, "fbank_options_reference.npz"))
    assert sorted(OPTION_SETS) == [str(n) for n in g["names"]]
    for name, kw in OPTION_SETS.items():
        for i in range(2):
            ref = g["%s_%d" % (name, i)]
            got = lasr_frontend.wav_to_kaldi_fbank(g["wav_%d" % i], **kw)
            assert got.shape == ref.shape and got.dtype == np.float32
            band = 1e-5 + 1e-4 * np.abs(ref)
            if not kw.get("use_log_fbank", True):
                band = band + 1e-7 * np.abs(ref).sum(axis=1, keepdims=True)      # linear energies: the fp32 FFT floor of the frame
            assert np.all(np.abs(got - ref) <= band), (name, i)


def test_cmvn_pinned_to_torchaudio_sliding_window_cmn(fb):
    """Row A11: utterance CMVN (mean / mean + variance), the Kaldi statistics layout and global CMVN from accumulated statistics
    against torchaudio.functional.sliding_window_cmn (torchaudio's port of Kaldi's apply-cmvn-sliding) with a window that covers
    the whole matrix -- on one utterance that is utterance CMVN, on the concatenated corpus it is global CMVN
    (oracle/gen_golden.py --cmvn; inputs are the reference's own fbank:80 outputs)."""
    cm = np.load(os.path.join(GOLD, "cmvn_reference.npz"))
    feats = [fb["fbank_%d" % i] for i in range(6)]
    stats = lasr_frontend.cmvn_stats(feats)
    assert stats[0, 80] == sum(x.shape[0] for x in feats)
    for nv, ukey, gkey in ((False, "utt_mean", "global_mean"), (True, "utt_meanvar", "global_meanvar")):
        mean, istd = lasr_frontend.cmvn_from_stats(stats, norm_vars=nv)
        for i, x in enumerate(feats):
            for key, (m, s) in ((gkey, (mean, istd)), (ukey, lasr_frontend.cmvn_from_stats(lasr_frontend.cmvn_stats([x]), norm_vars=nv))):
                if "%s_%d" % (key, i) not in cm.files:
                    continue
                ref = cm["%s_%d" % (key, i)]
                # (a) the DEFINITION (statistics layout, mean, population variance, no epsilon) in float64: equal to torchaudio's
                #     result up to the cancellation in sumsq / n - mean^2 (relative 1e-16 * mean^2 / var)
                x64 = x.astype(np.float64)
                d64 = np.abs((x64 - m) * s - ref)
                assert np.all(d64 <= 1e-9 + 1e-9 * (np.abs(x64) + np.abs(m)) ** 2 * s * s), (key, i, d64.max())
                # (b) the float32 evaluation the device epilogue mirrors: the tolerance plus the rounding of (x - mean) in float32
                #     (half an ulp of |x| and of |mean|), amplified by istd where a column hardly varies
                y = lasr_frontend.apply_cmvn(x, m, s)
                band = 1e-5 + 1e-4 * np.abs(ref) + 2.4e-7 * (np.abs(x64) + np.abs(m)) * s
                assert np.all(np.abs(y - ref) <= band), (key, i)
    # one frame: sliding_window_cmn divides by a zero variance; the definition floors it (Kaldi's apply-cmvn floor 1e-20)
    assert feats[0].shape[0] == 1 and "utt_meanvar_0" not in cm.files
    assert np.all(lasr_frontend.utterance_cmvn(feats[0]) == 0)


def test_masks_match_reference_fixture():
    """make_pad_mask / the encoder source mask / Conv2dSubsampling's mask slicing / hs_len against the reference's own
    functions (oracle/gen_golden.py: lasr.utils.mask.make_pad_mask + the slicing of subsampling.py:60)."""
    mk = np.load(os.path.join(GOLD, "mask_reference.npz"))
    for ci in mk["cases"]:
        lens, T = mk["c%d_len" % ci], int(mk["c%d_T" % ci])
        assert np.array_equal(lasr_frontend.make_pad_mask(lens, T), mk["c%d_pad" % ci])
        assert np.array_equal(lasr_frontend.src_mask(lens, T), mk["c%d_src" % ci])
        sub, hs = lasr_frontend.subsampled_mask(lens, T)
        assert np.array_equal(sub, mk["c%d_sub" % ci]) and np.array_equal(hs, mk["c%d_hslen" % ci])
