"""CPU: batch formation against goldens from the UNMODIFIED reference
(BatchAudioDataSet.check_dataset -> make_batch_size / make_batch_duration, lasr/data/dataset.py:260-305; generator:
oracle/gen_golden.py --batching) and the packed layout / sharding helpers."""
import os
import random

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "batch_reference.npz")


def test_batches_equal_the_reference(lasr_b200):
    import importlib
    bt = importlib.import_module("lighting-asr_b200.batching")
    g = np.load(GOLD)
    for ci in g["cases"]:
        kw = eval(str(g["c%d_kw" % ci]))                                   # plain dict literal written by the generator
        wav_len, token_len = g["c%d_wav_len" % ci].tolist(), g["c%d_token_len" % ci].tolist()
        random.seed(100 + int(ci))
        batches = bt.plan_batches(wav_len, token_len, **kw)
        assert [len(b) for b in batches] == g["c%d_sizes" % ci].tolist()
        assert [i for b in batches for i in b] == g["c%d_flat" % ci].tolist()      # same utterances, same order, ties included
        assert random.random() == float(g["c%d_after" % ci])                       # the global generator ends where the reference leaves it
        if kw["batch_type"] == "duration":
            for b in batches[:-1]:
                assert sum(wav_len[i] for i in b) >= kw["batch_duration"] > sum(wav_len[i] for i in b[:-1])


def test_packed_layout_and_sharding(lasr_b200):
    import importlib
    bt = importlib.import_module("lighting-asr_b200.batching")
    n = np.array([16000, 401, 7, 12344], dtype=np.int64)
    offs, total = bt.packed_layout(n)
    assert offs.tolist() == [0, 16000, 16404, 16412] and total == 16412 + 12344 and all(o % 4 == 0 for o in offs)
    offs16, total16 = bt.packed_layout(n, elem_bytes=2)
    assert all(o % 8 == 0 for o in offs16) and total16 >= int(n.sum())
    dur = np.random.default_rng(3).uniform(1, 30, 40).tolist()
    batches = bt.plan_batches(dur, [10] * 40, batch_type="duration", batch_duration=60, shuffle=False)
    shards = bt.shard_batches(batches, dur, 4)
    assert sorted(i for s in shards for b in s for i in b) == sorted(i for b in batches for i in b)
    loads = [sum(dur[i] for b in s for i in b) for s in shards]
    assert max(loads) - min(loads) <= max(sum(dur[i] for i in b) for b in batches)
