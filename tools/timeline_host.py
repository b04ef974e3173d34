"""Wall-clock timeline of one RGPE iteration on the host (where the non-overlapped time goes)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficient-tlbo_b200'))
import numpy as np, torch
import bench
from tlbo_b200 import ops
import tlbo_b200.facade.rgpe as R
import tlbo_b200.model.gp as G
class A: pass
args = A(); args.method = 'rgpe'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
fac, smbo = bench.make_product(args, wl)
for _ in range(5): smbo.iterate()
T = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); T[key] = T.get(key, 0.0) + time.perf_counter() - t0; return r
    setattr(obj, name, g)
wrap(G, 'fit_models_async', 'fit_models_async (prep+launch)')
wrap(ops, 'gp_fit_batched', '  gp_fit_batched (uploads+launch)')
wrap(np.random, 'standard_normal', 'standard_normal draws')
wrap(smbo.acq_optimizer, 'prepare', 'prepare (keys, excluded)')
wrap(ops, 'gp_predict', 'gp_predict enqueue')
wrap(ops, 'rank_loss_counts_normal', 'rank kernel enqueue')
wrap(ops, 'd2h', 'd2h (incl. GPU wait)')
wrap(ops, 'h2d', 'h2d pageable')
wrap(ops, 'predict_ei_fused', 'fused call (launch+sync)')
wrap(fac, '_ranking_losses', '_ranking_losses total')
wrap(fac, 'train', 'train total')
wrap(smbo, 'choose_next', 'choose_next total')
import tlbo_b200.facade.base_facade as BF
BF.fit_models_async = G.fit_models_async
torch.cuda.synchronize(); t0 = time.perf_counter()
N = 20
for _ in range(N): smbo.iterate()
torch.cuda.synchronize(); tot = time.perf_counter() - t0
print('ms/step %.3f' % (1e3 * tot / N))
for k, v in T.items(): print('  %-36s %.3f ms/step' % (k, 1e3 * v / N))
