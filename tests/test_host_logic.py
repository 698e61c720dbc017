"""CPU: host-side logic of the product (no CUDA needed): SpecAugment RNG replay against the
reference fixtures, CMVN statistics helpers, C-ABI library symbols, plugin boundary."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sg():
    return np.load(os.path.join(GOLD, "specaug_reference.npz"))


def _pattern(T):
    from oracle.gen_golden import pattern
    return pattern(T)


def _apply(masks_f, masks_t, T, D=80):
    m = np.zeros((T, D), dtype=bool)
    for lo, hi in masks_f:
        m[:, lo:hi] = True
    for lo, hi in masks_t:
        m[lo:hi] = True
    return m


def test_planner_positions_bit_exact(lasr_b200, sg):
    """Mask rectangles planned by the product equal the cells the reference overwrote
    (specaugment.py:47-106) for the same seeds; both RNG streams end in the same state."""
    for key in sg["cases"]:
        seed, T = int(key.split("_")[0][1:]), int(key.split("_T")[1])
        random.seed(seed)
        np.random.seed(seed)
        f, t = lasr_b200.specaug.plan_utterance(T, 80)
        assert np.array_equal(np.array([random.random(), np.random.rand()]), sg[key + "_rng_after"]), key
        planned = _apply(f, t, T)
        assert np.array_equal(np.packbits(planned), sg[key + "_zero_changed"]), key
        x = _pattern(T)
        changed = sg[key + "_out"] != x
        assert not np.any(changed & ~planned), key     # mean fill may coincide with a value, never exceed the plan


def test_planner_can_consume_time_warp_draws(lasr_b200, sg):
    """With consume_time_warp_draws the planner leaves both generators where the reference's full
    ``specaug`` transform (warp + masks, datatrans.py:136-150) leaves them."""
    for key in sg["cases"]:
        seed, T = int(key.split("_")[0][1:]), int(key.split("_T")[1])
        random.seed(seed)
        np.random.seed(seed)
        lasr_b200.specaug.plan_utterance(T, 80, consume_time_warp_draws=True)
        assert np.array_equal(np.array([random.random(), np.random.rand()]), sg[key + "_rng_after_full"]), key


def test_plan_batch_equals_sequential_planning(lasr_b200):
    """plan_batch vectorises the numpy draws; masks and both generator states must equal the
    per-utterance replay (which test_planner_positions_bit_exact pins to the reference)."""
    lens = [998] * 40 + [5, 12, 39, 98, 300, 7, 11, 1, 2, 40, 41, 3498]
    for kw in ({}, {"consume_time_warp_draws": True}, {"n_freq_mask": 3, "n_time_mask": 1, "max_freq_width": 15}):
        random.seed(21)
        np.random.seed(21)
        m, b = lasr_b200.specaug.plan_batch(lens, 80, **kw)
        after = (random.random(), np.random.rand(), int(np.random.randint(0, 1000)))
        random.seed(21)
        np.random.seed(21)
        ref = [lasr_b200.specaug.plan_utterance(t, 80, **kw) for t in lens]
        assert after == (random.random(), np.random.rand(), int(np.random.randint(0, 1000)))
        nf = kw.get("n_freq_mask", 2)
        for i, (f, t) in enumerate(ref):
            assert np.array_equal(m[i, :nf], f) and np.array_equal(m[i, nf:], t)
            assert np.array_equal(b[i], np.sort(t.reshape(-1)))


def test_plan_batch_layout(lasr_b200):
    random.seed(3)
    np.random.seed(3)
    masks, bounds = lasr_b200.specaug.plan_batch([98, 300, 12], 80)
    assert masks.shape == (3, 4, 2) and masks.dtype == np.int32
    assert bounds.shape == (3, 4) and np.all(np.diff(bounds, axis=1) >= 0)
    assert np.all(masks[:, :2] <= 80) and np.all(masks[0, 2:] <= 98) and np.all(masks[2, 2:] <= 12)
    with pytest.raises(ValueError):
        lasr_b200.specaug.plan_batch([98], 80, n_time_mask=9)


def test_header_symbols_exported(lasr_b200):
    """Every function declared in include/b200fe.h is exported by the built library."""
    hdr = open(os.path.join(ROOT, "include", "b200fe.h")).read()
    names = set(re.findall(r"\b(b200fe_[a-z0-9_]+)\s*\(", hdr))
    names -= {"b200fe_opts", "b200fe_plan", "b200fe_fbank_args", "b200fe_post_args"}
    assert {"b200fe_plan_create", "b200fe_fbank_fused", "b200fe_postpass", "b200fe_peak_absmax"} <= names
    lib = lasr_b200._lib.load()
    for n in sorted(names):
        assert hasattr(lib, n), n
    assert names == set(lasr_b200._lib.EXPORTS)


def test_cmvn_from_stats_host_function(lasr_b200):
    from oracle import lasr_frontend
    rng = np.random.default_rng(1)
    feats = [rng.normal(2, 3, (T, 80)).astype(np.float32) for T in (20, 33)]
    st = lasr_frontend.cmvn_stats(feats)
    lib = lasr_b200._lib.load()
    mean = np.zeros(80, np.float32)
    istd = np.zeros(80, np.float32)
    rc = lib.b200fe_cmvn_from_stats(np.ascontiguousarray(st).ctypes.data_as(C.POINTER(C.c_double)), 80, 1,
                                    mean.ctypes.data_as(lasr_b200._lib.c_fp), istd.ctypes.data_as(lasr_b200._lib.c_fp))
    assert rc == 0
    m, s = lasr_frontend.cmvn_from_stats(st)
    assert np.allclose(mean, m, rtol=1e-6) and np.allclose(istd, s, rtol=1e-6)
    m2, s2 = lasr_b200.cmvn.mean_istd(st)
    assert np.array_equal(m2, mean) and np.array_equal(s2, istd)
    bad = np.zeros((2, 81))
    assert lib.b200fe_cmvn_from_stats(bad.ctypes.data_as(C.POINTER(C.c_double)), 80, 1,
                                      mean.ctypes.data_as(lasr_b200._lib.c_fp), istd.ctypes.data_as(lasr_b200._lib.c_fp)) < 0
    assert b"zero frame count" in lib.b200fe_last_error()


def test_stats_file_roundtrip(lasr_b200, tmp_path):
    st = np.arange(2 * 81, dtype=np.float64).reshape(2, 81) * 1.5 + 0.1
    p = str(tmp_path / "cmvn.stats")
    lasr_b200.cmvn.save_stats(p, st)
    assert np.array_equal(lasr_b200.cmvn.load_stats(p), st)


def test_shard_utterances_balanced(lasr_b200):
    rng = np.random.default_rng(0)
    n = (rng.uniform(1, 35, 256) * 16000).astype(np.int64)
    for ws in (1, 2, 4, 8):
        parts = lasr_b200.cmvn.shard_utterances(n, ws)
        assert sorted(np.concatenate(parts).tolist()) == list(range(256))
        loads = np.array([n[p].sum() for p in parts])
        assert loads.max() - loads.min() <= n.max()


def test_no_cpu_path(lasr_b200):
    import torch
    fe = lasr_b200.GpuFbankFrontend()
    with pytest.raises(RuntimeError):
        fe(torch.zeros(1, 16000), np.array([16000]))
    with pytest.raises(ValueError):
        lasr_b200.GpuFbankFrontend(snip_edges=False)
    with pytest.raises(ValueError):
        lasr_b200.GpuFbankFrontend(cmvn="global")


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "lasr")), reason="the reference checkout is only present in the build container")
def test_plugin_points_with_the_real_reference(lasr_b200, capsys):
    """The two plug-in points of SURVEY 8(b) against the UNMODIFIED reference code (no GPU needed: plans are created lazily):
    the transform registry accepts the overrides with its own warning, and BaseConfig.check_kwargs accepts the YAML keys of
    the dataset class (every key is a named __init__ parameter; unknown keys raise as they do for the reference's classes)."""
    import sys
    import types
    for name in ("librosa", "soundfile"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from lasr.data.datatrans import register_trans
    from lasr.utils.generater import BaseConfig
    before = register_trans["fbank:80"]
    try:
        table = lasr_b200.lasr_plugin.install(register_trans, device="cuda:0", return_tensor=True)
        assert "has been registered before" in capsys.readouterr().out          # lasr/utils/register.py:10-12
        assert register_trans["fbank:80"] is table["fbank:80"] and "b200:norm+fbank:80" in register_trans
        assert callable(register_trans["b200:norm+fbank:80+specaug"])
        with pytest.raises(KeyError):
            register_trans["no-such-transform"]
    finally:
        register_trans.register("fbank:80")(before)
    cls = lasr_b200.lasr_plugin.make_dataset_class()
    assert cls.__mro__[1].__name__ == "BatchAudioDataSet"
    kwargs = dict(wav_list="wav.scp", text_list="text", batch_type="duration", batch_duration=500, peak_norm=True,
                  cmvn="utt_meanvar", specaug=True, device="cuda:0")
    cfg = BaseConfig("lasr_b200.lasr_plugin:B200BatchAudioDataSet", kwargs)                 # runs check_kwargs (generater.py:91-99)
    assert cfg.conf_class is cls
    with pytest.raises(ValueError):
        BaseConfig("lasr_b200.lasr_plugin:B200BatchAudioDataSet", dict(kwargs, not_a_parameter=1))
    # the front-end class itself is config-loadable too (model-side wrappers construct it from YAML)
    BaseConfig("lasr_b200:GpuFbankFrontend", dict(num_mel_bins=80, cmvn="global", cmvn_stats=[[0.0] * 80, [1.0] * 80], specaug=True))


def test_c_planner_replays_both_generators_exactly(lasr_b200):
    """b200fe_specaug_plan (one C call per batch) against the same plan drawn with the real ``random.randrange`` /
    ``numpy.random.randint`` calls: identical rectangles, row bounds and warp points, and both global generators end at the
    same position (next uniform and next Gaussian draws agree), from mid-block positions and for non-default mask counts."""
    sa = lasr_b200.specaug
    for seed in range(25):
        for consume in (False, True):
            for kw in ({}, {"max_freq_width": 15, "n_freq_mask": 3, "max_time_width": 70, "n_time_mask": 4}, {"n_freq_mask": 0, "n_time_mask": 1}):
                rng = np.random.default_rng(seed)
                T = rng.integers(1, 400, 48)
                T[:6] = [1, 5, 10, 11, 12, 39]
                random.seed(seed)
                np.random.seed(seed)
                for _ in range(seed % 3):
                    random.random()
                    np.random.rand()
                st0 = (random.getstate(), np.random.get_state())
                a = sa.plan_batch(T, 80, consume_time_warp_draws=consume, return_warp=True, **kw)
                after_a = (random.random(), np.random.rand(), random.gauss(0, 1), np.random.randn())
                random.setstate(st0[0])
                np.random.set_state(st0[1])
                b = sa.plan_batch_reference_loop(T, 80, consume_time_warp_draws=consume, return_warp=True, **kw)
                after_b = (random.random(), np.random.rand(), random.gauss(0, 1), np.random.randn())
                assert all(np.array_equal(x, y) for x, y in zip(a, b)), (seed, consume, kw)
                assert after_a == after_b, (seed, consume, kw)
    # 624-word block boundaries: enough draws to regenerate both generators' states several times
    random.seed(7)
    np.random.seed(7)
    T = np.full(700, 998)
    st0 = (random.getstate(), np.random.get_state())
    a = sa.plan_batch(T, 80, consume_time_warp_draws=True, return_warp=True)
    after_a = (random.random(), np.random.rand())
    random.setstate(st0[0])
    np.random.set_state(st0[1])
    b = sa.plan_batch_reference_loop(T, 80, consume_time_warp_draws=True, return_warp=True)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and after_a == (random.random(), np.random.rand())
