#!/usr/bin/env python
"""cProfile of the online iteration (SMBO_ONLINE.iterate) after warm-up: where the host time of one iteration goes.
    python tools/prof_online.py [rgpe|topo] [iters]"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from tlbo_b200 import ops, synthetic  # noqa: E402
from tlbo_b200.config_space import Configuration, TableSpace  # noqa: E402
from tlbo_b200.facade.rgpe import RGPE  # noqa: E402
from tlbo_b200.facade.topo_variant1 import OBTLV  # noqa: E402
from tlbo_b200.framework.smbo_online import SMBO_ONLINE  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else 'rgpe'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
seed, K = 4465, 29
types, const = synthetic.space_types('rf')
fam = synthetic.TaskFamily(types, const, n_tasks=30, seed=4466)
sources = []
for k in range(1, 30):
    rng = np.random.RandomState(1000 + k)
    X = synthetic.sample_table(rng, 50, types, const)
    sources.append((X, fam.evaluate(k, X, rng)))


def objective(config):
    return float(fam.evaluate(0, config.get_array()[np.newaxis, :])[0])


cs = TableSpace(types, const_mask=const)
fac = dict(rgpe=RGPE, topo=OBTLV)[method](cs, sources, None, seed, surrogate_type='gp', num_src_hpo_trial=50)
smbo = SMBO_ONLINE(objective, cs, fac, random_seed=seed, max_runs=10 ** 9, num_points=5000)
X0 = cs.sample_array(20)
for i in range(20):
    c = Configuration(cs, X0[i], -1)
    smbo.configurations.append(c)
    smbo.perfs.append(objective(c))
    smbo.history_container.add(c, smbo.perfs[-1])
for _ in range(3):
    smbo.iterate()
torch.cuda.synchronize()
ops.STATS.reset()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(iters):
    smbo.iterate()
torch.cuda.synchronize()
pr.disable()
print('ms/iter', (time.perf_counter() - t0) / iters * 1e3, 'launches/iter', ops.STATS.total_launches() / iters,
      dict(ops.STATS.launches))
print('local search', smbo.acq_optimizer.local_search.last_stats)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45)
print(s.getvalue()[:9000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(25)
print(s.getvalue()[:5000])
