"""CPU, world_size 2, gloo: the path's only collective (all-reduce of the global CMVN statistics)
and the per-utterance sharding give the same statistics as a single process."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, lens, seed, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lasr_b200
    from oracle import lasr_frontend
    rng = np.random.default_rng(seed)
    feats = [rng.normal(1.0, 2.0, (int(t), 80)).astype(np.float32) for t in lens]       # stand-in features
    mine = lasr_b200.cmvn.shard_utterances(lens, world)[rank]
    st = torch.from_numpy(lasr_frontend.cmvn_stats([feats[i] for i in mine]))
    lasr_b200.cmvn.allreduce_stats(st)
    q.put((rank, st.numpy().copy(), mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_equals_single_process():
    from oracle import lasr_frontend
    lens = [98, 300, 12, 998, 57, 640, 33, 1200, 5]
    seed, world = 17, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(seed)
    feats = [rng.normal(1.0, 2.0, (int(t), 80)).astype(np.float32) for t in lens]
    ref = lasr_frontend.cmvn_stats(feats)
    seen = sorted(i for _, _, idx in got for i in idx)
    assert seen == list(range(len(lens)))                      # every utterance on exactly one rank
    for _, st, _ in got:
        assert st.shape == (2, 81)
        assert np.allclose(st, ref, rtol=1e-12, atol=1e-9)     # fp64 partial sums: order-independent to ~1e-16
        assert st[0, 80] == sum(lens)
