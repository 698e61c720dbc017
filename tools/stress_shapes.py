"""Randomised shape stress of the fused launch (descriptor pipeline, work list, padding tiles): many small random batches, every
utterance's rows against the same utterance run alone (bit for bit: a tile never mixes utterances), padding rows zero, frame counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lasr_b200

dev = "cuda:0"
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
plain = lasr_b200.GpuFbankFrontend()
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 150
bad = 0
for case in range(n_cases):
    B = int(rng.integers(1, 48))
    kind = case % 4
    if kind == 0:
        lens = rng.integers(400, 3000, B)              # one or two tiles per utterance, far fewer tiles than CTAs
    elif kind == 1:
        lens = rng.integers(400, 120000, B)
    elif kind == 2:
        lens = np.full(B, int(rng.integers(400, 40000)))   # uniform: the static grid walk
    else:
        lens = np.where(rng.random(B) < 0.5, 400, rng.integers(30000, 200000, B))      # extreme raggedness: many padding tiles
    lens = lens.astype(np.int64)
    nmax = int((lens.max() + 3) // 4 * 4)
    wav = (torch.randn((B, nmax), device=dev) * 0.1).clamp_(-1, 1)
    for i, n in enumerate(lens):
        wav[i, n:] = 0
    mode = ("none", "utt_meanvar", "utt_mean")[case % 3]
    fe = lasr_b200.GpuFbankFrontend(cmvn=mode)
    extra = int(rng.integers(0, 40))
    T = 1 + (lens - 400) // 160
    feats, flen = fe(wav, lens, max_frames=int(T.max()) + extra)
    torch.cuda.synchronize()
    assert flen.cpu().tolist() == T.tolist(), (case, "frame counts")
    for i in range(B):
        one, _ = fe(wav[i:i + 1, : int((lens[i] + 3) // 4 * 4)].contiguous(), lens[i:i + 1])
        if not torch.equal(feats[i, : T[i]], one[0, : T[i]]) or bool((feats[i, T[i]:] != 0).any()):
            bad += 1
            print("MISMATCH case", case, "utt", i, "B", B, "len", int(lens[i]), "mode", mode, flush=True)
print("cases", n_cases, "mismatches", bad)
sys.exit(1 if bad else 0)
