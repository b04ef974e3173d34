"""Tie-break key generation on the device: one CTA against the jump-ahead sliced generator."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'efficient-tlbo_b200'))
import numpy as np, torch
from tlbo_b200 import ops
for N in (10 ** 5, 10 ** 6, 10 ** 7):
    for hint in (0, N):
        dev = ops.DeviceRandomState(np.random.RandomState(1), n_hint=hint)
        out = torch.empty(N, dtype=torch.float64, device='cuda')
        for _ in range(2):
            dev.rand_into(out)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dev.rand_into(out)
        e1.record(); torch.cuda.synchronize()
        print('N=%8d slices=%2d: %.3f ms per rng.rand(N) (incl. the next draw\'s jump)' % (N, dev._slices, e0.elapsed_time(e1) / 5))
