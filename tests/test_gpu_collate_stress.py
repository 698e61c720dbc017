"""GPU: one B200Collate object fed random batches of changing geometry and element type (float64, float32, int16 PCM, float64 that
holds PCM values) -- grow-only rings, stale padding ranges, head / tail taper, the automatic int16 upload -- against the device
path on the same waveforms, bit for bit."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_collate_random_batches_match_the_device_path(lasr_b200):
    rng = np.random.default_rng(99)
    col = lasr_b200.lasr_plugin.B200Collate(DEV, to_host=True, cmvn="utt_meanvar")
    col.pipeline.group_bytes = 1 << 20                        # several groups per batch at test sizes (taper at both ends)
    fe = lasr_b200.GpuFbankFrontend(cmvn="utt_meanvar")
    taken = 0
    for case in range(36):
        B = int(rng.integers(1, 48))
        lens = rng.integers(400, int(rng.choice([2000, 60000, 250000])), B)
        kind = ("f64", "f32", "i16", "f64pcm")[case % 4]
        pcm = [np.round(np.clip(rng.normal(0, 0.1, n), -1, 1) * 32767).astype(np.int16) for n in lens]
        if kind == "i16":
            wavs, ref = pcm, [k.astype(np.float32) / np.float32(32768.0) for k in pcm]
        elif kind == "f64pcm":
            wavs = [k.astype(np.float64) / 32768.0 for k in pcm]
            ref = [w.astype(np.float32) for w in wavs]
        else:
            wavs = [np.clip(rng.normal(0, 0.1, n), -1, 1).astype(np.float64 if kind == "f64" else np.float32) for n in lens]
            ref = [w.astype(np.float32) for w in wavs]
        before = col.pipeline.pcm16_batches
        batch = col(wavs)
        taken += col.pipeline.pcm16_batches - before
        assert (col.pipeline.pcm16_batches - before == 1) == (kind == "f64pcm"), (case, kind)
        nmax = int((max(lens) + 3) // 4 * 4)
        buf = np.zeros((B, nmax), dtype=np.float32)
        for i, w in enumerate(ref):
            buf[i, : len(w)] = w
        want, wlen = fe(torch.from_numpy(buf).to(DEV), np.asarray(lens, dtype=np.int64))
        torch.cuda.synchronize()
        assert torch.equal(batch["wav_len"], wlen.cpu()), (case, kind)
        assert torch.equal(batch["wav_array"], want.cpu()), (case, kind)
    assert taken == 9
