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
        self._gather = None

    def row_range(self, N):
        per = (N + self.world_size - 1) // self.world_size
        lo = min(self.rank * per, N)
        return lo, min(lo + per, N)

    def reduce_best(self, best):
        """All-gather the 4-double result of the fused kernel and merge.  ``best`` is a device tensor under
        NCCL (merged ON the device by ``tlbo_merge_best``, the fused kernel's own final merge stage -- no
        host round trip between the collective and the merge; the merged
        quadruple is then read back once, which the caller needs anyway to name the selected row) or a
        CPU tensor under gloo (merged on the host)."""
        import torch
        if best.is_cuda:
            if self._gather is None or self._gather.device != best.device:
                self._gather = torch.empty(self.world_size, 4, dtype=best.dtype, device=best.device)
            self.dist.all_gather_into_tensor(self._gather, best.reshape(1, 4), group=self.group)
            return tuple(merge_best_device(self._gather))
        out = [torch.empty_like(best) for _ in range(self.world_size)]
        self.dist.all_gather(out, best, group=self.group)
        quads = torch.stack(out).cpu().numpy()
        return merge_best(quads)


def merge_best_device(q):
    """``merge_best`` over a device tensor ``q [G, 4]``: ONE launch of the fused kernel's own final merge stage
    (``tlbo_merge_best``, stream-ordered behind the collective), then the 32-byte read-back."""
    from tlbo_b200 import ops
    h = ops.d2h(ops.merge_best(q))
    if h[2] < 0:
        return (-np.inf, -np.inf, -1, int(h[3]))
    return (float(h[0]), float(h[1]), int(h[2]), int(h[3]))
