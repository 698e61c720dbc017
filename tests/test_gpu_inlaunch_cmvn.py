"""GPU: utterance CMVN applied INSIDE the fused launch (opt-in CMVN-apply tiles of b200fe_build_work_list_device) against the
two-launch finalize + post-pass path and against the oracle's definition (`_subtract_column_mean`, TA:220-226, and its
mean/variance extension)."""
import numpy as np
import pytest
import torch

from oracle import kaldi_fbank, lasr_frontend

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(lens, seed):
    rng = np.random.default_rng(seed)
    n = np.asarray(lens, dtype=np.int64)
    nmax = int((n.max() + 3) // 4 * 4)
    buf = np.zeros((len(lens), nmax), dtype=np.float32)
    for i, k in enumerate(lens):
        buf[i, :k] = np.clip(rng.normal(0, 0.1, k), -1, 1)
    return torch.from_numpy(buf).to(DEV), n


def _error_flag(fe):
    work, idx = fe.last["apply_flags"]
    return int(work[idx].item())


@pytest.mark.parametrize("mode", ["utt_meanvar", "utt_mean"])
def test_apply_tiles_equal_post_pass(lasr_b200, mode):
    """Ragged batch with utterances shorter than a tile, exactly one apply tile (256 rows), one row more, and several tiles:
    the in-launch path must reproduce the post-pass path (same fp64 -> fp32 vectors, same fp32 arithmetic; the statistics
    differ only by the order of the fp64 atomics)."""
    lens = [400, 400 + 160 * 30, 400 + 160 * 31, 400 + 160 * 32, 400 + 160 * 255, 400 + 160 * 256, 400 + 160 * 257,
            16000 * 11 + 7, 959, 16000 * 35, 16000 * 2, 4000] + [16000 + 1237 * i for i in range(40)]
    wav, n = _batch(lens, 21)
    ref_fe = lasr_b200.GpuFbankFrontend(cmvn=mode)
    ref_fe.inlaunch_cmvn = False
    ref, rlen = ref_fe(wav, n)
    assert ref_fe.last["apply_flags"] is None
    for lag in (16, 1, 3, 1000):
        fe = lasr_b200.GpuFbankFrontend(cmvn=mode)
        fe.inlaunch_cmvn = True
        fe.apply_lag = lag
        got, glen = fe(wav, n)
        torch.cuda.synchronize()
        assert fe.last["apply_flags"] is not None, "the in-launch path did not run"
        assert _error_flag(fe) == 0
        assert torch.equal(glen, rlen)
        assert torch.allclose(fe.last["utt_mean"], ref_fe.last["utt_mean"], rtol=1e-6, atol=1e-6)
        assert torch.allclose(fe.last["utt_istd"], ref_fe.last["utt_istd"], rtol=1e-6, atol=0)
        d = (got - ref).abs()
        assert float(d.max()) <= 2e-5, (lag, float(d.max()))
        g = got.cpu().numpy()
        for i, k in enumerate(lens):
            T = kaldi_fbank.num_frames(k)
            assert np.all(g[i, T:] == 0)


def test_apply_tiles_against_oracle_definition(lasr_b200):
    """The apply tiles against the fp64 definition evaluated on the device's own log-mel features."""
    lens = [16000 * 9 + 5, 400, 16000 * 4, 700, 16000 * 20 + 3]
    wav, n = _batch(lens, 22)
    raw = lasr_b200.GpuFbankFrontend()(wav, n)[0].cpu().numpy()
    for mode, nv in (("utt_meanvar", True), ("utt_mean", False)):
        fe = lasr_b200.GpuFbankFrontend(cmvn=mode)
        fe.inlaunch_cmvn = True
        g = fe(wav, n)[0].cpu().numpy()
        assert fe.last["apply_flags"] is not None and _error_flag(fe) == 0
        for i, k in enumerate(lens):
            T = kaldi_fbank.num_frames(k)
            ref = lasr_frontend.utterance_cmvn(raw[i, :T], norm_vars=nv).astype(np.float64)
            assert int((np.abs(g[i, :T] - ref) > 1e-5 + 1e-4 * np.abs(ref)).sum()) == 0
            assert np.all(g[i, T:] == 0)


def test_apply_tiles_repeated_calls_and_single_utterance(lasr_b200):
    """One frontend object reused over batches of different shapes (fresh counters per call), including a single utterance
    (every apply tile closes the list) and a caller-provided output buffer that holds garbage."""
    fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
    fe.inlaunch_cmvn = True
    ref_fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
    ref_fe.inlaunch_cmvn = False
    for seed, lens in ((1, [16000 * 30]), (2, [16000 * 3] * 7), (3, [401, 16000 * 17, 560]), (4, [16000 * 30])):
        wav, n = _batch(lens, seed)
        T = [kaldi_fbank.num_frames(k) for k in lens]
        out = torch.full((len(lens), max(T) + 5, 80), float("nan"), device=DEV)
        got, _ = fe(wav, n, max_frames=max(T) + 5, out=out)
        ref, _ = ref_fe(wav, n, max_frames=max(T) + 5)
        torch.cuda.synchronize()
        assert _error_flag(fe) == 0
        assert not torch.isnan(got).any()
        assert float((got - ref).abs().max()) <= 2e-5


def test_apply_tiles_c_abi_argument_errors(lasr_b200):
    """b200fe_fbank_fused rejects apply_cmvn_mode without the pieces it needs (no silent fall-through)."""
    import ctypes as C
    _lib = lasr_b200._lib
    fe = lasr_b200.GpuFbankFrontend()
    plan = fe.plan(torch.device(DEV))
    lib = plan.lib
    wav, n = _batch([16000], 5)
    nd = torch.from_numpy(n).to(DEV)
    out = torch.empty((1, 98, 80), device=DEV)
    a = _lib.FbankArgs()
    a.d_wav, a.wav_stride, a.d_nsamp, a.batch = wav.data_ptr(), wav.stride(0), nd.data_ptr(), 1
    a.d_out, a.max_frames = out.data_ptr(), 98
    a.apply_cmvn_mode = 2                      # no d_utt_done / work list / statistics
    assert lib.b200fe_fbank_fused(plan.handle, C.byref(a), None) != 0
    assert b"in-launch utterance CMVN" in lib.b200fe_last_error()
    a.apply_cmvn_mode = 3
    assert lib.b200fe_fbank_fused(plan.handle, C.byref(a), None) != 0
    # the list builder refuses a lag of 0 and option sets without apply tiles
    work = torch.empty((64,), dtype=torch.int32, device=DEV)
    # the list builder refuses apply tiles without counters, a misaligned zero-fill buffer, and option sets without apply tiles
    assert lib.b200fe_build_work_list_device(plan.handle, nd.data_ptr(), 1, 98, 1, 4, work.data_ptr(), 8, work.data_ptr() + 128,
                                             work.data_ptr() + 132, None, None, 0, None) != 0
    assert lib.b200fe_build_work_list_device(plan.handle, nd.data_ptr(), 1, 98, 1, 0, work.data_ptr(), 8, work.data_ptr() + 128,
                                             work.data_ptr() + 132, None, work.data_ptr() + 4, 16, None) != 0
    fe40 = lasr_b200.GpuFbankFrontend(num_mel_bins=40, cmvn="utt_meanvar")
    fe40.inlaunch_cmvn = True
    got, _ = fe40(wav, n)                      # falls back to the post pass: the option set has no lean kernel
    assert fe40.last["apply_flags"] is None and got.shape == (1, 98, 40)
