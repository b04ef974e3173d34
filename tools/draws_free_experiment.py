"""Timing experiment (NOT a parity path): is the host's 1 ms of bootstrap draws on the critical path of the single-trial
step?  The same step with np.random.standard_normal replaced by a slice of one pre-drawn array."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    sys.path.insert(0, p)
import numpy as np, torch
import bench


class A:
    pass


args = A()
args.method = 'rgpe'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for mode in ('real draws', 'pre-drawn', 'real draws', 'pre-drawn'):
    wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
    fac, smbo = bench.make_product(args, wl)
    orig = np.random.standard_normal
    pool = np.random.RandomState(1).standard_normal(1 << 20)
    if mode == 'pre-drawn':
        np.random.standard_normal = lambda size=None: pool[:int(np.prod(size))].reshape(size).copy()
    try:
        for _ in range(5):
            smbo.iterate()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            flush.zero_()
            smbo.iterate()
        e1.record(); torch.cuda.synchronize()
        print('%-12s %.3f ms/step' % (mode, e0.elapsed_time(e1) / 20))
    finally:
        np.random.standard_normal = orig
    del fac, smbo
