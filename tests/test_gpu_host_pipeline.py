"""GPU: the host pipeline behind B200Collate (lists of host ndarrays in, the reference's batch tensors out) and the
round-1 advisor findings: host synchronisation before host tensors are handed out, batches of CHANGING geometry,
zero-fill masks together with utterance CMVN, int16 statistics with peak normalisation, sharded statistics."""
import random

import numpy as np
import pytest
import torch

from oracle import kaldi_fbank, lasr_frontend

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(got, ref, rtol=1e-4, atol=1e-5):
    return int((np.abs(got.astype(np.float64) - ref) > atol + rtol * np.abs(ref)).sum())


def _batches(rng, shapes):
    return [[np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens] for lens in shapes]


def test_collate_to_host_needs_no_external_sync_and_handles_changing_geometry(lasr_b200):
    """B200Collate(to_host=True) returns finished host tensors (no torch.cuda.synchronize by the caller) for batches whose B,
    Nmax and Tmax all change from call to call; rows past an utterance's frames are zero (batch_list, dataset.py:8-22)."""
    rng = np.random.default_rng(21)
    shapes = [(16000, 4800, 32001), (48000, 400, 9999, 16000, 25000), (8000,), (160000, 16000), (4800, 32001, 16000)]
    batches = _batches(rng, shapes)
    col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True, cmvn="utt_meanvar", ring=2)
    direct = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
    plain = lasr_b200.GpuFbankFrontend()
    for rep in range(2):
        for wavs, lens in zip(batches, shapes):
            out = col(wavs)
            feats, flen = out["wav_array"], out["wav_len"]
            assert not feats.is_cuda and feats.is_pinned() and feats.dtype == torch.float32 and feats.is_contiguous()
            T = [kaldi_fbank.num_frames(n) for n in lens]
            assert tuple(feats.shape) == (len(lens), max(T), 80) and flen.tolist() == T
            g = feats.numpy().copy()                              # read immediately: no synchronize in between
            nmax = (max(lens) + 3) // 4 * 4
            buf = np.zeros((len(lens), nmax), dtype=np.float32)
            for i, w in enumerate(wavs):
                buf[i, : len(w)] = w
            dw = torch.from_numpy(buf).to(DEV)
            want = direct(dw, np.array(lens, dtype=np.int64))[0].cpu().numpy()
            assert np.allclose(g, want, rtol=1e-5, atol=1e-5)
            raw = plain(dw, np.array(lens, dtype=np.int64))[0].cpu().numpy()      # parity of the raw features: test_gpu_fbank_parity.py
            for i, t in enumerate(T):
                assert np.all(g[i, t:] == 0)
                ref = lasr_frontend.utterance_cmvn(raw[i, :t])                      # fp64 definition on the device's own features
                assert _close(g[i, :t], ref.astype(np.float64), rtol=1e-4, atol=1e-5) == 0


def test_collate_input_types_and_prefetch(lasr_b200):
    rng = np.random.default_rng(22)
    shapes = [(16000, 24000, 4321), (32000, 8000), (12345, 54321, 16000, 999)]
    b64 = _batches(rng, shapes)
    col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True)
    want = [col(w)["wav_array"].clone() for w in b64]
    got32 = [col([x.astype(np.float32) for x in w])["wav_array"].clone() for w in b64]
    for a, b in zip(want, got32):
        assert torch.equal(a, b)                                  # float64 -> float32 happens once, on the host, round to nearest
    pcm = [[np.round(x * 32767).astype(np.int16) for x in w] for w in b64]
    for w16, w in zip(pcm, b64):
        a = col(w16)["wav_array"].clone()
        b = col([x.astype(np.float32) / 32768.0 for x in w16])["wav_array"]
        assert torch.equal(a, b)                                  # (float)s16 == float sample * 2^15 exactly
    mixed = [pcm[0][0], b64[0][1].astype(np.float32), (pcm[0][2].astype(np.float64) / 32768.0)]      # int16 + float32 + float64 in one list
    want_mixed = col([pcm[0][0].astype(np.float32) / 32768.0, b64[0][1].astype(np.float32), pcm[0][2].astype(np.float32) / 32768.0])["wav_array"].clone()
    assert torch.equal(col(mixed)["wav_array"], want_mixed)
    outs = [d["wav_array"].clone() for d in col.prefetch(b64)]
    assert len(outs) == len(want) and all(torch.equal(a, b) for a, b in zip(outs, want))
    dev_col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=False)
    d = dev_col(b64[0])
    assert d["wav_array"].is_cuda and torch.equal(d["wav_array"].cpu(), want[0]) and d["wav_len"].cpu().tolist() == [kaldi_fbank.num_frames(n) for n in shapes[0]]
    with pytest.raises(AssertionError):
        col([np.zeros(16000), np.zeros(399)])                     # torchaudio asserts on short input (TA:142)
    with pytest.raises(ValueError):
        col([np.zeros((100, 2, 2))])


def test_float64_lists_that_hold_pcm16_values_go_up_as_int16(lasr_b200):
    """The reference's reader hands over float64 = int16 / 32768 (soundfile.read of PCM_16 files, reader.py:24).  Such lists are
    staged and uploaded as int16 (half the bytes) with BIT-IDENTICAL features; a batch with one sample off the PCM grid -- where the
    probe does not look -- is caught by the check inside the packing and repeated as float32; float lists never take the path."""
    rng = np.random.default_rng(21)
    lens = (16000, 48123, 9000, 70001, 33333)
    pcm = [np.round(np.clip(rng.normal(0, 0.1, n), -1, 1) * 32767).astype(np.int16) for n in lens]
    f64 = [k.astype(np.float64) / 32768.0 for k in pcm]
    for kw in (dict(cmvn="utt_meanvar"), dict(peak_norm=True)):
        auto = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True, **kw)
        plain = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True, **kw)
        plain.pipeline.pcm16_auto = False
        a, b = auto(f64), plain(f64)
        assert auto.pipeline.pcm16_batches == 1 and plain.pipeline.pcm16_batches == 0
        assert auto.pipeline.h2d_bytes * 2 <= plain.pipeline.h2d_bytes + 64
        assert torch.equal(a["wav_len"], b["wav_len"]) and torch.equal(a["wav_array"], b["wav_array"])
        # the same as int16 lists
        c = plain(pcm)
        assert torch.equal(a["wav_array"], c["wav_array"])
        # one sample off the grid in the middle of an utterance: caught while packing, repeated as float32, then a pause
        bad = [w.copy() for w in f64]
        bad[3][35000] += 1e-7                # (the probe's windows of this utterance start at 0, 9991, ..., 29973, 39964, ...)
        turn = auto.pipeline._turn
        a2, b2 = auto(bad), plain(bad)
        assert auto.pipeline._turn == turn + 1                  # the repeat reuses the abandoned attempt's ring slots
        assert auto.pipeline.pcm16_batches == 1 and auto.pipeline._pcm16_skip == auto.pipeline._pcm16_backoff
        assert torch.equal(a2["wav_array"], b2["wav_array"]) and torch.equal(a2["wav_len"], b2["wav_len"])
        a3 = auto(f64)                       # inside the pause: float32 path, same features
        assert auto.pipeline.pcm16_batches == 1 and torch.equal(a3["wav_array"], b["wav_array"])
        auto.pipeline._pcm16_skip = 0
        # float data: the probe says no, nothing is attempted
        flt = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
        a4, b4 = auto(flt), plain(flt)
        assert auto.pipeline.pcm16_batches == 1 and auto.pipeline._pcm16_skip == 0
        assert torch.equal(a4["wav_array"], b4["wav_array"])
        # features stay on the device / prefetch
        dev_col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=False, **kw)
        d = dev_col(f64)
        torch.cuda.synchronize()
        ref = plain(f64)                     # (b's ring slot has been handed out again by now)
        assert dev_col.pipeline.pcm16_batches == 1 and torch.equal(d["wav_array"].cpu(), ref["wav_array"])


def test_zero_masks_with_utterance_cmvn_are_applied_after_normalisation(lasr_b200):
    """specaug + replace_with_zero + utterance CMVN: statistics over the UNMASKED features, zeros written after the
    normalisation (CMVN first, masks second -- the order of the mean-fill and time-warp paths)."""
    rng = np.random.default_rng(23)
    lens = [48000, 160000, 7 * 16000 + 77, 6000]
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1) for n in lens]
    nmax = (max(lens) + 3) // 4 * 4
    buf = np.zeros((len(lens), nmax), dtype=np.float32)
    for i, w in enumerate(wavs):
        buf[i, : len(w)] = w
    wav, n = torch.from_numpy(buf).to(DEV), np.array(lens, dtype=np.int64)
    raw = lasr_b200.GpuFbankFrontend()(wav, n)[0].cpu().numpy()
    for mode, nv in (("utt_meanvar", True), ("utt_mean", False)):
        fe = lasr_b200.GpuFbankFrontend(cmvn=mode, specaug=True, replace_with_zero=True)
        random.seed(5)
        np.random.seed(5)
        g = fe(wav, n)[0].cpu().numpy()
        random.seed(5)
        np.random.seed(5)
        for i, k in enumerate(lens):
            T = kaldi_fbank.num_frames(k)
            x = lasr_frontend.utterance_cmvn(raw[i, :T], nv)
            y, rects = lasr_frontend.spec_augment_masks(x.copy(), replace_with_zero=True)
            masked = y != x
            assert np.all(g[i, :T][y == 0] == 0)
            assert _close(g[i, :T], y.astype(np.float64), rtol=1e-4, atol=2e-5) == 0
            assert masked.any() or not rects
            assert np.all(g[i, T:] == 0)


def test_accumulate_stats_int16_peak_norm_and_validation(lasr_b200):
    rng = np.random.default_rng(24)
    lens = [16000, 23456, 8000]
    pcm = [np.round(np.clip(rng.normal(0, 0.2, n), -1, 1) * 32767).astype(np.int16) for n in lens]
    nmax = (max(lens) + 7) // 8 * 8
    b16 = np.zeros((3, nmax), dtype=np.int16)
    for i, w in enumerate(pcm):
        b16[i, : len(w)] = w
    n = np.array(lens, dtype=np.int64)
    fe = lasr_b200.GpuFbankFrontend(peak_norm=True)
    st16 = fe.accumulate_stats(torch.from_numpy(b16).to(DEV), n).cpu().numpy()
    st32 = fe.accumulate_stats(torch.from_numpy(b16.astype(np.float32) / 32768.0).to(DEV), n).cpu().numpy()
    assert np.allclose(st16, st32, rtol=1e-9, atol=1e-6) and st16[0, 80] == sum(kaldi_fbank.num_frames(k) for k in lens)
    feats = fe(torch.from_numpy(b16).to(DEV), n)[0].cpu().numpy()
    ref = lasr_frontend.cmvn_stats([feats[i, : kaldi_fbank.num_frames(k)] for i, k in enumerate(lens)])
    assert np.allclose(st16, ref, rtol=2e-7, atol=1e-4)
    with pytest.raises(AssertionError):
        fe.accumulate_stats(torch.zeros((1, 400), device=DEV), np.array([399]))
    with pytest.raises(ValueError):
        fe.accumulate_stats(torch.zeros((1, 400), device=DEV), np.array([401]))
    with pytest.raises(ValueError):
        fe.accumulate_stats(torch.zeros((1, 400)), np.array([400]))


def test_sharded_statistics_sum_to_the_fp64_definition(lasr_b200):
    """the product's accumulate_stats on two utterance shards (cmvn.shard_utterances), summed like the
    all-reduce does, equals the fp64 definition (oracle cmvn_stats) on the whole batch's features."""
    rng = np.random.default_rng(25)
    lens = np.round(rng.uniform(0.5, 6.0, 24) * 16000).astype(np.int64)
    wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1).astype(np.float32) for n in lens]
    fe = lasr_b200.GpuFbankFrontend()
    total = torch.zeros((2, 81), dtype=torch.float64, device=DEV)
    shards = lasr_b200.cmvn.shard_utterances(lens, 2)
    assert sorted(np.concatenate(shards).tolist()) == list(range(24))
    feats_all = []
    for idx in shards:
        nmax = int((lens[idx].max() + 3) // 4 * 4)
        buf = np.zeros((len(idx), nmax), dtype=np.float32)
        for j, u in enumerate(idx):
            buf[j, : lens[u]] = wavs[u]
        w = torch.from_numpy(buf).to(DEV)
        part = fe.accumulate_stats(w, lens[idx])
        total += part                                             # what all_reduce(sum) computes over the ranks
        f = fe(w, lens[idx])[0].cpu().numpy()
        feats_all += [f[j, : kaldi_fbank.num_frames(int(lens[u]))] for j, u in enumerate(idx)]
    ref = lasr_frontend.cmvn_stats(feats_all)
    assert total[0, 80].item() == ref[0, 80]
    assert np.allclose(total.cpu().numpy(), ref, rtol=2e-7, atol=1e-4)
    mean, istd = lasr_b200.cmvn.mean_istd(total)
    rm, ri = lasr_frontend.cmvn_from_stats(ref)
    assert np.allclose(mean, rm, rtol=1e-6, atol=1e-6) and np.allclose(istd, ri, rtol=1e-6, atol=1e-6)


def test_bfloat16_feature_emission(lasr_b200):
    """Row F2: features in the consumer's precision (Conv2dSubsampling under autocast, subsampling.py:53-57): the float32
    result rounded once to bfloat16 (nearest even, = tensor.to(torch.bfloat16)), on the device and through the D2H copy."""
    rng = np.random.default_rng(26)
    shapes = [(16000, 24000, 4321), (40000, 8000, 30011, 12000)]
    batches = _batches(rng, shapes)
    ref_col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True, cmvn="utt_meanvar")
    for to_host in (True, False):
        col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=to_host, cmvn="utt_meanvar", out_dtype=torch.bfloat16)
        for wavs in batches + batches[::-1]:
            want = ref_col(wavs)
            got = col(wavs)
            f = got["wav_array"]
            assert f.dtype == torch.bfloat16 and f.is_cuda != to_host and tuple(f.shape) == tuple(want["wav_array"].shape)
            assert torch.equal(f.cpu(), want["wav_array"].to(torch.bfloat16))
            assert torch.equal(got["wav_len"].cpu(), want["wav_len"])
    fe = lasr_b200.GpuFbankFrontend()
    buf = np.zeros((2, 24000), dtype=np.float32)
    buf[0, :16000] = batches[0][0]
    buf[1] = batches[0][1]
    w = torch.from_numpy(buf).to(DEV)
    n = np.array([16000, 24000])
    f32, _ = fe(w, n)
    f16, fl = fe(w, n, out_dtype=torch.bfloat16)
    assert f16.dtype == torch.bfloat16 and torch.equal(f16, f32.to(torch.bfloat16)) and fl.tolist() == [98, 148]
    with pytest.raises(ValueError):
        fe(w, n, out_dtype=torch.float16)


def test_cmvn_stats_cli(lasr_b200, tmp_path):
    """Row F4: tools/compute_cmvn_stats.py writes the Kaldi text statistics of a corpus (npy + PCM-16 wav files); the file
    loads into cmvn='global' and equals the fp64 definition on the device's features."""
    import sys
    import wave
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1] / "tools"))
    import compute_cmvn_stats as cli
    rng = np.random.default_rng(27)
    paths, wavs = [], []
    for i, n in enumerate((16000, 24000, 4321, 300, 9999)):
        w = np.clip(rng.normal(0, 0.2, n), -1, 1)
        if i % 2 == 0:
            p = tmp_path / ("u%d.npy" % i)
            np.save(p, w.astype(np.float32))
            wavs.append(w.astype(np.float32))
        else:
            p = tmp_path / ("u%d.wav" % i)
            pcm = np.round(w * 32767).astype(np.int16)
            with wave.open(str(p), "wb") as f:
                f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(pcm.tobytes())
            wavs.append(pcm.astype(np.float32) / 32768.0)
        paths.append(str(p))
    lst = tmp_path / "wavs.txt"
    lst.write_text("\n".join("utt%d %s" % (i, p) for i, p in enumerate(paths)) + "\n")
    out = tmp_path / "cmvn.stats"
    cli.main(["--list", str(lst), "--out", str(out), "--batch-seconds", "2"])
    st = lasr_b200.cmvn.load_stats(str(out))
    fe = lasr_b200.GpuFbankFrontend()
    feats = []
    for w in wavs:
        if len(w) < 400:
            continue
        pad = (-len(w)) % 4
        f, fl = fe(torch.from_numpy(np.pad(w, (0, pad))).to(DEV).unsqueeze(0), np.array([len(w)]))
        feats.append(f[0, : int(fl[0])].cpu().numpy())
    ref = lasr_frontend.cmvn_stats(feats)
    assert st.shape == (2, 81) and st[0, 80] == ref[0, 80]
    assert np.allclose(st, ref, rtol=2e-7, atol=1e-4)
    g = lasr_b200.GpuFbankFrontend(cmvn="global", cmvn_stats=st)
    assert np.allclose(g.cmvn_mean.numpy(), lasr_frontend.cmvn_from_stats(ref)[0], rtol=1e-6, atol=1e-6)
