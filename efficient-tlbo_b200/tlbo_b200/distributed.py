"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink; gloo on
CPU for tests).

The algorithm has exactly one exchange step: with the candidate pool
sharded by rows, every rank emits the ``(ei, key, global row, #negative)`` quadruple of its
best unevaluated row and all ranks all-gather the G quadruples (32 B each) and apply the
same lexicographic rule, arriving at the same row.  Independent BO trials are replicas and
need no collective."""
import numpy as np


def lexi_better(a, b):
    """Order of ``np.lexsort((key, ei))[::-1]``: larger EI, then larger key, then larger row."""
    if a[0] != b[0]:
        return a[0] > b[0]
    if a[1] != b[1]:
        return a[1] > b[1]
    return a[2] > b[2]


def merge_best(quads):
    """Combine per-shard ``(ei, key, row, n_negative)`` rows (row < 0 = shard had no
    unevaluated candidate)."""
    best = None
    nneg = 0
    for q in quads:
        nneg += int(q[3])
        if q[2] < 0:
            continue
        cand = (float(q[0]), float(q[1]), int(q[2]))
        if best is None or lexi_better(cand, best):
            best = cand
    if best is None:
        return (-np.inf, -np.inf, -1, nneg)
    return (best[0], best[1], best[2], nneg)


class PoolShard(object):
    """Row-range sharding of a candidate table over the ranks of a process group."""

    def __init__(self, rank=None, world_size=None, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size

    def row_range(self, N):
        per = (N + self.world_size - 1) // self.world_size
        lo = min(self.rank * per, N)
        return lo, min(lo + per, N)

    def reduce_best(self, best):
        """All-gather the 4-double result of the fused kernel and merge.  ``best`` is a
        device tensor under NCCL or a CPU tensor under gloo."""
        import torch
        out = [torch.empty_like(best) for _ in range(self.world_size)]
        self.dist.all_gather(out, best, group=self.group)
        quads = torch.stack(out).cpu().numpy()
        return merge_best(quads)
