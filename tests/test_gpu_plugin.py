"""GPU: the reference-facing plug-in points: registry transforms and the batched collate."""
import numpy as np
import pytest
import torch

from conftest import fbank_parity
from oracle import kaldi_fbank, lasr_frontend

pytestmark = pytest.mark.gpu


def _parity(got, wav64, ref32):
    """north_star tolerance with the documented noise-floor rule (tests/conftest.py::fbank_parity): no violation above the
    floor, floor cells within the wide band of the fp64 oracle; prints the direct violation count against live torchaudio."""
    ref64 = lasr_frontend.wav_to_kaldi_fbank(wav64, dtype=np.float64)
    lin64 = lasr_frontend.wav_to_kaldi_fbank(wav64, dtype=np.float64, use_log_fbank=False)
    hard, soft, below = fbank_parity(got, ref32, ref64, lin64)
    direct = int((np.abs(got - ref32) > 1e-5 + 1e-4 * np.abs(ref32)).sum())
    print("direct violations vs live torchaudio: %d of %d cells (%d below the fp32 noise floor)" % (direct, ref32.size, below))
    assert hard == 0 and soft == 0 and below <= max(3, ref32.size // 10000)


class Register(dict):
    """Behavioural stand-in for lasr/utils/register.py:1-41 (register(name) decorator, override with a
    warning, lookup by key) -- /root/reference is not present on the GPU box."""

    def register(self, target):
        def add(key, value):
            if not callable(value):
                raise Exception("register object must be callable")
            if key in self:
                print("warning: %s has been registered before, so we will overriden it" % value.__name__)   # register.py:10-11
            self[key] = value
            return value
        return (lambda x: add(target, x)) if not callable(target) else add(target.__name__, target)


def test_registry_override_and_fused_chain(lasr_b200, capsys):
    reg = Register()
    reg.register("fbank:80")(lambda w: None)                       # the reference's own CPU transform
    table = lasr_b200.lasr_plugin.install(reg, device="cuda:0")
    assert "has been registered before" in capsys.readouterr().out   # override semantics of Register.register
    assert set(table) <= set(reg.keys())
    rng = np.random.default_rng(3)
    wav = rng.uniform(-0.5, 0.5, 24001)                              # float64, as soundfile.read returns
    out = reg["fbank:80"](wav)                                      # exactly one positional argument (dataset.py:196-197)
    ref = lasr_frontend.wav_to_kaldi_fbank(wav, use_torchaudio=True)
    assert isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == ref.shape
    assert out.shape[0] == kaldi_fbank.num_frames(24001)             # wav_len = wav_array.shape[0] (dataset.py:198)
    _parity(out, wav, ref)
    fused = reg["b200:norm+fbank:80"](wav)                          # replaces audio_trans: [norm, "fbank:80"]
    ref2 = lasr_frontend.wav_to_kaldi_fbank(lasr_frontend.voice_norm(wav), use_torchaudio=True)
    _parity(fused, lasr_frontend.voice_norm(wav), ref2)
    t = lasr_b200.lasr_plugin.GpuTransform("cuda:0", return_tensor=True)(wav)   # ASRProcess path: torch.as_tensor(feats)
    assert t.is_cuda and torch.equal(torch.as_tensor(t).cpu(), torch.from_numpy(out))
    with pytest.raises(ValueError):
        reg["fbank:80"](np.zeros((100, 2)))
    with pytest.raises(AssertionError):
        reg["fbank:80"](np.zeros(399))                               # torchaudio asserts on short input (TA:142)


def test_batched_collate_matches_reference_batch_dict(lasr_b200):
    rng = np.random.default_rng(4)
    wavs = [rng.uniform(-0.5, 0.5, n) for n in (16000, 4800, 32001)]
    for to_host in (False, True):
        col = lasr_b200.lasr_plugin.B200Collate("cuda:0", to_host=to_host)
        batch = col(wavs)
        feats, flen = batch["wav_array"], batch["wav_len"]
        assert feats.is_cuda != to_host and feats.dtype == torch.float32 and flen.dtype == torch.int64
        ref = lasr_frontend.batch_list([lasr_frontend.wav_to_kaldi_fbank(w, use_torchaudio=True) for w in wavs], pad_value=0)
        assert tuple(feats.shape) == ref.shape
        assert flen.cpu().tolist() == [kaldi_fbank.num_frames(len(w)) for w in wavs]
        g = feats.cpu().numpy()
        for i, w in enumerate(wavs):
            t = kaldi_fbank.num_frames(len(w))
            _parity(g[i, :t], w, ref[i, :t])
        assert np.array_equal(g == 0, ref == 0)                      # identical zero padding


def test_device_masks_match_reference(lasr_b200):
    """Row F2: the encoder masks from the device-resident frame counts equal the reference's host-built ones bit for bit
    (goldens from lasr.utils.mask.make_pad_mask, the slicing of subsampling.py:60 and subfunction, e2e_base.py:47-49)."""
    import os
    import numpy as np
    from oracle import lasr_frontend
    mk = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mask_reference.npz"))
    M = lasr_b200.mask
    for ci in mk["cases"]:
        lens, T = mk["c%d_len" % ci], int(mk["c%d_T" % ci])
        d = torch.from_numpy(lens).cuda()
        assert np.array_equal(M.make_pad_mask(d, max_length=T).cpu().numpy(), mk["c%d_pad" % ci])
        src = M.src_mask(d, T)
        assert src.dtype == torch.bool and np.array_equal(src.cpu().numpy(), mk["c%d_src" % ci])
        sub, hs = M.subsampled_mask(d, T)
        assert np.array_equal(sub.cpu().numpy(), mk["c%d_sub" % ci]) and np.array_equal(hs.cpu().numpy(), mk["c%d_hslen" % ci])
    # straight from sample counts, chained after the front end without touching the host
    rng = np.random.default_rng(2)
    n = np.round(rng.uniform(0.1, 4.0, 9) * 16000).astype(np.int64)
    fe = lasr_b200.GpuFbankFrontend()
    wav = torch.zeros((9, int(n.max())), device="cuda")
    feats, flen = fe(wav, n)
    T = feats.shape[1]
    m1, l1 = M.src_mask_from_samples(fe.plan(wav.device), torch.from_numpy(n).cuda(), T)
    assert torch.equal(m1, M.src_mask(flen, T)) and torch.equal(l1, flen)
    m4, l4 = M.src_mask_from_samples(fe.plan(wav.device), torch.from_numpy(n).cuda(), T, subsample=4)
    want, whs = lasr_frontend.subsampled_mask(flen.cpu().numpy(), T)
    assert np.array_equal(m4.cpu().numpy(), want) and np.array_equal(l4.cpu().numpy(), whs)
    # (B, T, D)-shaped request, as make_pad_mask(lengths, xs, length_dim=1)
    pad3 = M.make_pad_mask(flen, feats, 1)
    assert pad3.shape == feats.shape and torch.equal(pad3[:, :, 0], ~M.src_mask(flen, T).squeeze(1))
    with pytest.raises(RuntimeError):
        M.src_mask(flen.cpu(), T)
