"""world_size-2 CPU (gloo) test of the only exchange step of the path: candidate-pool row
sharding + all-gather of the per-rank (ei, key, row, n_negative) quadruple + replicated
lexicographic arg-max."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, N, seed, out):
    sys.path.insert(0, os.path.join(ROOT, 'efficient-tlbo_b200'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tlbo_b200.distributed import PoolShard
    rng = np.random.RandomState(seed)
    ei = np.round(rng.rand(N), 2)
    key = rng.rand(N)
    excluded = set(rng.choice(N, size=N // 10, replace=False).tolist())
    shard = PoolShard()
    lo, hi = shard.row_range(N)
    rows = [r for r in range(lo, hi) if r not in excluded]
    if rows:
        b = max(rows, key=lambda r: (ei[r], key[r], r))
        quad = torch.tensor([ei[b], key[b], float(b), 0.0], dtype=torch.float64)
    else:
        quad = torch.tensor([-np.inf, -np.inf, -1.0, 0.0], dtype=torch.float64)
    res = shard.reduce_best(quad)
    order = np.lexsort((key, ei))[::-1]
    want = next(int(r) for r in order if int(r) not in excluded)
    out[rank] = (res[2], want, lo, hi)
    dist.destroy_process_group()


@pytest.mark.parametrize('N,seed', [(1000, 0), (7, 1), (64, 2)])
def test_pool_shard_argmax_world2(N, seed):
    world = 2
    port = 29500 + (os.getpid() + seed * 7) % 2000
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, N, seed, out), nprocs=world, join=True)
    assert len(out) == 2
    ranges = sorted((out[r][2], out[r][3]) for r in range(world))
    assert ranges[0][0] == 0 and ranges[-1][1] == N and ranges[0][1] == ranges[1][0]
    for r in range(world):
        assert out[r][0] == out[r][1]
