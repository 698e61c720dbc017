"""Global CMVN statistics: accumulation across ranks and a Kaldi-compatible text file.

The reference has no CMVN code; the semantics adopted are Kaldi's
``compute-cmvn-stats`` / ``apply-cmvn --norm-vars``: a float64 ``[2, D+1]`` matrix whose row 0 is
the per-dimension sum followed by the frame count and whose row 1 is the per-dimension sum of
squares followed by 0.  The only collective on the whole hot path is one all-reduce (sum) of this
2 x (D+1) matrix: NCCL over NVLink on GPUs, gloo on CPU tensors in tests.
"""
import numpy as np
import torch


def allreduce_stats(stats, group=None):
    """In-place sum of the [2, D+1] float64 statistics over all ranks (no-op without a process group).

    Deterministic: every rank contributes one fp64 partial; the sum of <= 8 doubles per entry is
    order-independent to ~1e-16 relative, far below the fp32 apply step."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def shard_utterances(lengths, world_size):
    """Length-balanced assignment of whole utterances to ranks: greedy by descending length
    (SURVEY.md 8(e)).  Returns a list of index arrays, one per rank; every index appears once."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    buckets = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        buckets[r].append(int(i))
        load[r] += int(lengths[i])
    return [np.array(sorted(b), dtype=np.int64) for b in buckets]


def mean_istd(stats, norm_vars=True, var_floor=1e-20):
    """[2, D+1] statistics -> (mean, istd) float32 arrays (same arithmetic as b200fe_cmvn_from_stats)."""
    st = np.asarray(stats.cpu() if torch.is_tensor(stats) else stats, dtype=np.float64)
    d = st.shape[1] - 1
    n = st[0, d]
    if not n > 0:
        raise ValueError("CMVN statistics hold no frames")
    mean = st[0, :d] / n
    var = np.maximum(st[1, :d] / n - mean * mean, var_floor)
    istd = 1.0 / np.sqrt(var) if norm_vars else np.ones(d)
    return mean.astype(np.float32), istd.astype(np.float32)


def save_stats(path, stats):
    """Kaldi text matrix: `` [\\n  row0\\n  row1 ]``."""
    st = np.asarray(stats.cpu() if torch.is_tensor(stats) else stats, dtype=np.float64)
    with open(path, "w") as f:
        f.write(" [\n")
        for r, row in enumerate(st):
            f.write("  " + " ".join(repr(float(v)) for v in row) + (" ]\n" if r == len(st) - 1 else "\n"))


def load_stats(path):
    with open(path) as f:
        txt = f.read().replace("[", " ").replace("]", " ")
    rows = [ln.split() for ln in txt.strip().splitlines() if ln.strip()]
    return np.array([[float(v) for v in r] for r in rows], dtype=np.float64)
