"""Inclusive host time per BO step of the calls that make up an RGPE trial's host segment (solo trial)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from tlbo_b200 import ops  # noqa: E402
from tlbo_b200.utils import normalization  # noqa: E402


class A:
    pass


args = A()
args.method = 'rgpe'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
fac, smbo = bench.make_product(args, wl)
for _ in range(5):
    smbo.iterate()
acc = {}
T = time.perf_counter


def wrap(obj, name, label=None):
    f = getattr(obj, name)
    label = label or name

    def g(*a, **k):
        t = T()
        try:
            return f(*a, **k)
        finally:
            acc[label] = acc.get(label, 0.0) + (T() - t)
    setattr(obj, name, g)


acq = smbo.acq_optimizer
wrap(smbo, 'choose_next')
wrap(fac, 'train')
wrap(fac, '_ranking_enqueue')
wrap(fac, '_prepare_cv_jobs')
wrap(fac._stage, 'upload', 'stage.upload')
wrap(fac, '_predict_slots')
wrap(ops, 'rank_loss_counts_normal')
wrap(ops, 'rgpe_weights')
wrap(ops, 'gp_fit_batched')
wrap(fac, '_ranking_finish')
wrap(fac, '_apply_weights')
wrap(acq, 'prepare', 'acq.prepare')
wrap(acq, 'best_unevaluated', 'acq.best_unevaluated')
wrap(acq, '_consume_keys', 'acq._consume_keys')
wrap(smbo.acquisition_function, 'update', 'ei.update')
wrap(smbo.acquisition_function, 'best_unevaluated', 'ei.best_unevaluated')
wrap(ops, 'd2h')
wrap(np.random, 'standard_normal')
wrap(smbo, 'evaluate')
N = 40
torch.cuda.synchronize()
t0 = T()
for _ in range(N):
    smbo.iterate()
torch.cuda.synchronize()
print('ms/step %.3f' % ((T() - t0) / N * 1e3))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print('  %-28s %8.1f us / step' % (k, v / N * 1e6))
