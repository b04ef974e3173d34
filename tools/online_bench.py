#!/usr/bin/env python
"""Online-scenario benchmark (BASELINE configs[2] shape; SURVEY 8 f rank 1): BO iterations / second
of ``SMBO_ONLINE`` -- facade fit + weights, then ``InterleavedLocalAndRandomSearch`` (10 hill-climbs
from the best evaluated configurations + ~4990 random configurations sorted by EI) -- with K source
surrogates, on one GPU; beside it the oracle's literal sequential restatement on one host core for a
bounded number of iterations.  The objective is the synthetic task family (no real model training).

    python tools/online_bench.py [--method rgpe|topo] [--sources 29] [--steps 20] [--warmup 3] [--cpu-steps 1]
"""
import argparse
import json
import os
import sys
import time

for _v in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS'):      # as tools/online_benchmark.py:3-8 does
    os.environ[_v] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--method', default='rgpe', choices=['rgpe', 'topo'])
    ap.add_argument('--sources', type=int, default=29)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--n0', type=int, default=20)
    ap.add_argument('--num-points', type=int, default=5000)
    ap.add_argument('--cpu-steps', type=int, default=1)
    args = ap.parse_args()
    import torch
    from tlbo_b200 import ops, synthetic
    from tlbo_b200.config_space import Configuration, TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.facade.topo_variant1 import OBTLV
    from tlbo_b200.framework.smbo_online import SMBO_ONLINE
    seed, K = 4465, args.sources
    types, const = synthetic.space_types('rf')
    fam = synthetic.TaskFamily(types, const, n_tasks=30, seed=4466)
    ids = [i for i in range(1, 30)][:K]
    sources = []
    for k in ids:
        rng = np.random.RandomState(1000 + k)
        X = synthetic.sample_table(rng, 50, types, const)
        sources.append((X, fam.evaluate(k, X, rng)))

    def objective(config):
        x = config.get_array() if hasattr(config, 'get_array') else np.asarray(config)
        return float(fam.evaluate(0, x[np.newaxis, :])[0])

    cs = TableSpace(types, const_mask=const)
    fac = dict(rgpe=RGPE, topo=OBTLV)[args.method](cs, sources, None, seed, surrogate_type='gp', num_src_hpo_trial=50)
    smbo = SMBO_ONLINE(objective, cs, fac, random_seed=seed, max_runs=10 ** 9, num_points=args.num_points)
    X0 = cs.sample_array(args.n0)
    for i in range(args.n0):
        c = Configuration(cs, X0[i], -1)
        smbo.configurations.append(c)
        smbo.perfs.append(objective(c))
        smbo.history_container.add(c, smbo.perfs[-1])
    for _ in range(args.warmup):
        smbo.iterate()
    ops.STATS.reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps_ls, nb_ls = 0, 0
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        smbo.iterate()
        st = smbo.acq_optimizer.local_search.last_stats
        steps_ls += st['steps']
        nb_ls += st['neighbours']
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    line = {
        'metric': 'BO iters/sec, online scenario (fit+weights+interleaved local/random EI search) at %d sources' % K,
        'value': args.steps / (ms * 1e-3), 'unit': 'iters/s', 'ms_per_step': ms / args.steps,
        'host_wall_ms_per_step': 1e3 * wall / args.steps, 'steps': args.steps, 'warmup': args.warmup,
        'config': {'workload': '%s online BO, %d source GPs x 50 obs, %d challengers per iteration (10 local searches '
                               '+ sorted random search), d=9 random_forest space, target n=%d..%d'
                               % (args.method, K, args.num_points, args.n0 + args.warmup,
                                  args.n0 + args.warmup + args.steps - 1)},
        'gpu_launches': ops.STATS.total_launches(), 'gpu_launch_detail': dict(ops.STATS.launches),
        'local_search': {'hill_climb_steps_per_iter': steps_ls / float(args.steps),
                         'neighbours_the_reference_would_score_one_by_one_per_iter': nb_ls / float(args.steps)},
    }
    if args.cpu_steps > 0:
        import oracle.facade as of
        from oracle import ei_optimization as oeo
        ofac = dict(rgpe=of.OracleRGPE, topo=of.OracleOBTLV)[args.method](types, sources, seed, literal_loops=True)
        osmbo = oeo.OracleSMBOOnline(objective, oeo.OracleSpace(types, const), ofac, seed, num_points=args.num_points)
        osmbo.configurations = [c.get_array() for c in smbo.configurations[:args.n0 + args.warmup]]
        osmbo.perfs = list(smbo.perfs[:args.n0 + args.warmup])
        t0 = time.perf_counter()
        for _ in range(args.cpu_steps):
            osmbo.iterate()
        cpu = (time.perf_counter() - t0) / args.cpu_steps
        line['cpu_baseline'] = {'value': 1.0 / cpu, 'unit': 'iters/s', 'cores': 1, 'kind': 'port',
                                'sample': '%d iteration(s) of the literal sequential oracle (one acquisition call per '
                                          'neighbour, Python pair loops) at target n=%d' % (args.cpu_steps,
                                                                                           args.n0 + args.warmup)}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
