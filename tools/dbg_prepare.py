import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficient-tlbo_b200'))
import numpy as np, torch
import bench
from tlbo_b200.optimizer import ei_offline_optimizer as eo
class A: pass
args = A(); args.method = 'topo'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
fac, smbo = bench.make_product(args, wl)
orig = eo.OfflineSearch.prepare
def timed_prepare(self, evaluated_rows):
    from tlbo_b200 import ops
    t = [time.perf_counter()]
    keys = self.rng.rand(self.N); t.append(time.perf_counter())
    self._keys_host.copy_(torch.from_numpy(keys[self.lo:self.hi])); t.append(time.perf_counter())
    self._keys_dev.copy_(self._keys_host, non_blocking=True); t.append(time.perf_counter())
    ex = None
    if len(evaluated_rows):
        ex = self._stage.upload('excluded', np.sort(np.asarray(evaluated_rows, dtype=np.int64)))
    t.append(time.perf_counter())
    self._prepared = (ex,)
    print('prepare: rand %.2f  host copy %.2f  h2d async %.2f  excluded %.2f ms' % tuple(1e3 * (b - a) for a, b in zip(t[:-1], t[1:])))
eo.OfflineSearch.prepare = timed_prepare
for _ in range(6): smbo.iterate()
