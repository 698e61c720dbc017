"""GPU: randomised shapes through the fused launch -- work list, padding tiles, the three-stage descriptor pipeline, the static grid
walk of small uniform batches.  Every utterance of a random batch against the same utterance run alone: a tile never mixes
utterances, so plain fbank rows are bit-identical; utterance CMVN may differ by the order of its fp64 atomics only."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_random_batches_equal_single_utterance_runs(lasr_b200):
    rng = np.random.default_rng(2026)
    torch.manual_seed(7)
    for case in range(48):
        B = int(rng.integers(1, 40))
        kind = case % 4
        if kind == 0:
            lens = rng.integers(400, 3000, B)                  # one or two tiles per utterance: far fewer tiles than resident CTAs
        elif kind == 1:
            lens = rng.integers(400, 120000, B)
        elif kind == 2:
            lens = np.full(B, int(rng.integers(400, 40000)))   # uniform: the static grid walk
        else:
            lens = np.where(rng.random(B) < 0.5, 400, rng.integers(30000, 200000, B))      # extreme raggedness: many padding tiles
        lens = lens.astype(np.int64)
        nmax = int((lens.max() + 3) // 4 * 4)
        wav = (torch.randn((B, nmax), device=DEV) * 0.1).clamp_(-1, 1)
        for i, n in enumerate(lens):
            wav[i, n:] = 0
        mode = ("none", "utt_meanvar", "utt_mean")[case % 3]
        fe = lasr_b200.GpuFbankFrontend(cmvn=mode)
        T = 1 + (lens - 400) // 160
        feats, flen = fe(wav, lens, max_frames=int(T.max()) + int(rng.integers(0, 40)))
        assert flen.cpu().tolist() == T.tolist()
        for i in range(B):
            one, _ = fe(wav[i:i + 1, : int((lens[i] + 3) // 4 * 4)].contiguous(), lens[i:i + 1])
            a, b = feats[i, : T[i]], one[0, : T[i]]
            if mode == "none":
                assert torch.equal(a, b), (case, i)
            else:
                assert float((a - b).abs().max()) <= 1e-5, (case, i)
            assert not bool((feats[i, T[i]:] != 0).any()), (case, i)
