"""Host time between the arg-max result of one RGPE iteration and the fit launch of the next (the device is idle in
between), itemised."""
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
import tlbo_b200._cabi as C  # noqa: E402


class A:
    pass


args = A()
args.method = 'rgpe'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
fac, smbo = bench.make_product(args, wl)
for _ in range(5):
    smbo.iterate()
marks = {}
T = time.perf_counter


def stamp(name):
    marks.setdefault(name, []).append(T())


def wrap(obj, name, before=None, after=None):
    f = getattr(obj, name)

    def g(*a, **k):
        if before:
            stamp(before)
        r = f(*a, **k)
        if after:
            stamp(after)
        return r
    setattr(obj, name, g)


wrap(smbo.acq_optimizer, 'best_unevaluated', after='best_done')
wrap(smbo, 'choose_next', before='choose_next')
wrap(fac, 'train', before='train')
wrap(fac, '_prepare_cv_jobs', before='prep_cv', after='prep_cv_done')
wrap(ops, 'prepare_jobs', before='native_prep', after='native_prep_done')
wrap(ops, 'gp_fit_batched', before='fit_batched', after='fit_launched')
orig = C.lib.tlbo_gp_fit_batched


def fit_call(*a):
    stamp('fit_call')
    return orig(*a)


ops.lib.tlbo_gp_fit_batched = fit_call
N = 40
torch.cuda.synchronize()
t0 = T()
for _ in range(N):
    stamp('iterate')
    smbo.iterate()
    stamp('iterate_done')
torch.cuda.synchronize()
print('ms/step %.3f' % ((T() - t0) / N * 1e3))
m = {k: np.array(v) for k, v in marks.items()}
n = N - 1
seq = ['best_done', 'iterate_done', 'iterate', 'choose_next', 'train', 'prep_cv', 'native_prep', 'native_prep_done',
       'prep_cv_done', 'fit_batched', 'fit_call', 'fit_launched']
base = m['best_done'][:n]
prev = base
for name in seq[1:]:
    v = m[name][1:n + 1] if name != 'iterate_done' else m[name][:n]
    print('  -> %-18s +%6.1f us  (cumulative %6.1f us)' % (name, 1e6 * np.mean(v - prev), 1e6 * np.mean(v - base)))
    prev = v
