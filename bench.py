#!/usr/bin/env python
"""Headline benchmark of the surrogate-and-acquisition hot path.

Metric (BASELINE.json): BO iterations / second, one iteration = fit (target + 5 CV GPs, 11
L-BFGS-B restarts each) + ensemble weights (RGPE bootstrap ranking loss or TOPO SLSQP) +
fused predict / EI / arg-max over the whole candidate pool, at 29 source surrogates.
Workload = BASELINE.json configs[1]: 29 source GPs (50 observations each), 10^5-candidate
pool, random_forest-shaped space (d = 9), synthetic histories.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

Prints ONE JSON line (rank 0).  N > 1: one process per GPU (torchrun); ranks run independent
BO trials (different target task / seed) -- the path shards by trial with no data-path
collective ("weak" scaling); a pool-sharded run of ONE trial (tiny NCCL all-gather of the
per-rank arg-max) is reported beside it under "pool_sharded".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# The reference's driver pins every BLAS/OpenMP pool to one thread (tools/offline_benchmark.py:3-8); must
# happen before numpy is imported.  Both arms run that way: the CPU arm because the reference mandates it,
# the CUDA arm because its only host numerics are scipy's SLSQP iterations on 30-variable problems, where a
# multi-threaded BLAS pool costs far more in wake-ups than it saves.
for _v in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'VECLIB_MAXIMUM_THREADS',
           'NUMEXPR_NUM_THREADS'):
    os.environ[_v] = '1'

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'efficient-tlbo_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = 'BO iters/sec (fit+weights+EI) at 29 sources'
UNIT = 'iters/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--method', default='rgpe', choices=['rgpe', 'topo', 'topo_v3', 'obtl', 'trans_rgpe', 'tst', 'pogpe'])
    ap.add_argument('--surrogate', default='gp', choices=['gp', 'gp_mcmc'],
                    help="'gp_mcmc': MCMC-marginalised GP hyper-parameters (BASELINE configs[4])")
    ap.add_argument('--mcmc-steps', type=int, nargs=2, default=[250, 250], metavar=('BURNIN', 'CHAIN'),
                    help='emcee burn-in / chain steps per fit (reference: 250 250, model_builder.py:74-89)')
    ap.add_argument('--cpu-mcmc-steps', type=int, default=5,
                    help='reference arm with gp_mcmc: burn-in = chain = this many steps, fit time scaled to the '
                         'full chain length (the sampler cost is linear in the number of steps)')
    ap.add_argument('--pool', type=int, default=100000)
    ap.add_argument('--sources', type=int, default=29)
    ap.add_argument('--n0', type=int, default=40, help='target observations before the first step')
    ap.add_argument('--cpu-sample-rows', type=int, default=10000)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-pool-sharded', action='store_true')
    ap.add_argument('--distinct-trials', action='store_true',
                    help='N > 1: give every rank its own target task / seed instead of a replica of trial 0')
    ap.add_argument('--ref-procs', type=int, default=0, help='reference arm: worker processes (0 = all cores)')
    ap.add_argument('--pair-loops', default='literal', choices=['literal', 'vectorised'],
                    help='reference arm: literal Python pair loops (as the reference) or numpy broadcasting')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# workload
def build_workload(task, seed, pool, K, n0, n_tasks=30):
    """Synthetic histories shaped like the reference's pickles: 30 tasks,
    RF-shaped space; the K sources contribute their first 50 rows, the target a pool table."""
    from tlbo_b200 import synthetic
    types, const = synthetic.space_types('rf')
    fam = synthetic.TaskFamily(types, const, n_tasks=n_tasks, seed=4466)
    ids = [i for i in range(n_tasks) if i != task]
    np.random.RandomState(seed).shuffle(ids)
    ids = ids[:K]
    sources = []
    for k in ids:
        rng = np.random.RandomState(1000 + k)
        X = synthetic.sample_table(rng, 50, types, const)
        sources.append((X, fam.evaluate(k, X, rng)))
    rng = np.random.RandomState(2000 + task)
    Xt = synthetic.sample_table(rng, pool, types, const)
    yt = fam.evaluate(task, Xt, rng)
    rows0 = [0] + list(np.random.RandomState(seed + 1).choice(np.arange(1, pool), size=n0 - 1, replace=False))
    return dict(types=types, sources=sources, Xt=Xt, yt=yt, rows0=[int(r) for r in rows0], seed=seed)


def fit_flops(models, d_cont, d_cat):
    """SURVEY.md 8(d) algorithmic flops of the NLL + gradient evaluations one batched fit
    actually performed (sum over surrogates and restarts of nfev)."""
    H = d_cont + d_cat + 2
    f, evals = 0.0, 0
    for m in models:
        info = getattr(m, 'fit_info_', None)
        if info is None:
            continue
        n = m.n_train_
        nfev = int(info['restart_info'][:, 0].sum())
        evals += nfev
        f += nfev * (n * n * (3 * d_cont + 2 * d_cat + 12) + n ** 3 / 3.0 + 2.0 * n ** 3 + 2.0 * n * n * H)
    return f, evals


def fused_flops_per_candidate(n_list, d_cont, d_cat):
    """SURVEY.md 8(d) algorithmic flops of one candidate through the surrogates actually
    evaluated (triangular L^-1 form: n^2 + 2n for the quadratic form)."""
    f = 30.0
    for n in n_list:
        f += n * (3 * d_cont + 2 * d_cat + 12) + 2 * n + n * n + 2 * n
    return f


# ---------------------------------------------------------------------------------------------
# clocks
class ClockSampler(object):
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                    samples=len(sm))


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle (numpy/scipy restatement of the reference path), reference-faithful
# variant (literal Python pair loops as in tlbo/facade/rgpe.py:74-109).
def _oracle_facade(method, types, sources, seed, literal, surrogate='gp', mcmc_steps=None):
    import oracle.facade as of
    base, kw = {'rgpe': (of.OracleRGPE, {}), 'trans_rgpe': (of.OracleRGPE, {'smooth': 0.7}),
                'topo': (of.OracleOBTLV, {}), 'topo_v3': (of.OracleTOPOV3, {}), 'obtl': (of.OracleOBTL, {}), 'tst': (of.OracleTST, {}), 'pogpe': (of.OraclePOGPE, {})}[method]
    if surrogate == 'gp_mcmc':
        base = type('Mcmc' + base.__name__, (base,), dict(surrogate_type='gp_mcmc', mcmc_steps=tuple(mcmc_steps)))
    return base(types, sources, seed, literal_loops=literal, **kw)


def oracle_trial(method, task, seed, pool, K, n0, steps, warmup, sample_rows, literal=True, surrogate='gp',
                 mcmc_steps=None, mcmc_full=None):
    """Runs warmup + steps BO iterations of the oracle.  Per step: the full fit + weights
    (timed), and EI + ordering over the first ``sample_rows`` pool rows (timed, then scaled
    by pool / sample_rows: the work is exactly linear in the number of candidates).
    Returns (train_s[], maximize_s_scaled[], setup_s)."""
    from oracle.acquisition import OracleEI, first_unevaluated, sort_by_acq_value
    from oracle.facade import zero_mean_unit_var_normalization, zero_one_normalization
    wl = build_workload(task, seed, pool, K, n0)
    t0 = time.perf_counter()
    fac = _oracle_facade(method, wl['types'], wl['sources'], seed, literal, surrogate, mcmc_steps)
    setup = time.perf_counter() - t0
    # gp_mcmc: the chains run `mcmc_steps` steps and the fit time is scaled to the full length
    # (log-probability evaluations: W * (steps + 2) per fit, the rest of train() is negligible)
    fit_scale = 1.0
    if surrogate == 'gp_mcmc':
        fit_scale = (sum(mcmc_full) + 2.0) / (sum(mcmc_steps) + 2.0)
    np.random.seed(seed)
    acq_rng = np.random.RandomState(seed)
    ei = OracleEI(fac)
    Xs = wl['Xt'][:sample_rows]
    evaluated = list(wl['rows0'])
    scale = float(pool) / float(sample_rows)
    tr, mx = [], []
    for it in range(warmup + steps):
        X = wl['Xt'][evaluated]
        Y = wl['yt'][evaluated].copy()
        a = time.perf_counter()
        fac.train(X, Y)
        b = time.perf_counter()
        y_ = (zero_one_normalization(Y) if fac.method_id in ['tst', 'tstm'] else zero_mean_unit_var_normalization(Y))[0]
        ei.update(model=fac, eta=np.min(y_), num_data=len(evaluated))
        acq = ei(Xs)
        order, _ = sort_by_acq_value(acq, acq_rng)
        idx = first_unevaluated(order, evaluated)
        c = time.perf_counter()
        evaluated.append(idx)
        if it >= warmup:
            tr.append((b - a) * fit_scale)
            mx.append((c - b) * scale)
    return tr, mx, setup


def _oracle_worker(args):
    return oracle_trial(*args)


def cpu_baseline_sample(args, n_start):
    """One reference-faithful oracle iteration on this box's host (1 core, BLAS pinned to one
    thread as the reference mandates), bounded sample.  Runs in a child process so the thread
    pins take effect before numpy loads."""
    def child(loops):
        cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--ref-procs', '1', '--steps', '1',
               '--warmup', '0', '--n0', str(n_start), '--method', args.method, '--pool', str(args.pool),
               '--sources', str(args.sources), '--cpu-sample-rows', str(args.cpu_sample_rows),
               '--pair-loops', loops, '--surrogate', args.surrogate, '--mcmc-steps', str(args.mcmc_steps[0]),
               str(args.mcmc_steps[1]), '--cpu-mcmc-steps', str(args.cpu_mcmc_steps)]
        env = dict(os.environ)
        for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
            env.pop(k, None)
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=1200)
        if out.returncode != 0:
            raise RuntimeError('cpu baseline child failed: ' + out.stderr[-2000:])
        return json.loads(out.stdout.strip().splitlines()[-1])
    lit = child('literal')
    vec = child('vectorised')
    tr, mx = [lit['train_s_per_step']], [lit['maximize_s_per_step_scaled']]
    t = tr[0] + mx[0]
    tv = vec['train_s_per_step'] + vec['maximize_s_per_step_scaled']
    return dict(value=1.0 / t, unit=UNIT, cores=1, kind='port',
                sample=('1 BO iteration of the numpy/scipy oracle at target n=%d, K=%d: fit+weights at full size '
                        '(literal Python pair loops as the reference), EI+sort on the first %d of %d pool rows '
                        'scaled x%.0f' % (n_start, args.sources, args.cpu_sample_rows, args.pool,
                                          args.pool / args.cpu_sample_rows)) + _mcmc_sample_note(args),
                train_s=tr[0], maximize_s_scaled=mx[0],
                value_vectorised_pair_count=1.0 / tv, host_cores_available=os.cpu_count())


def _mcmc_sample_note(args):
    if args.surrogate != 'gp_mcmc':
        return ''
    return ('; gp_mcmc: chains of %d+%d emcee steps timed, fit time scaled to %d+%d steps (linear in the number '
            'of log-probability evaluations)' % (args.cpu_mcmc_steps, args.cpu_mcmc_steps, args.mcmc_steps[0],
                                                 args.mcmc_steps[1]))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    P = args.ref_procs or (os.cpu_count() or 1)
    n_start = args.n0
    # same trial selection as the CUDA arm: replicas of trial 0 unless --distinct-trials
    tid = (lambda t: t) if args.distinct_trials else (lambda t: 0)
    jobs = [(args.method, tid(t) % 30, 4465 + tid(t), args.pool, args.sources, n_start, args.steps, args.warmup,
             args.cpu_sample_rows, args.pair_loops == 'literal', args.surrogate,
             (args.cpu_mcmc_steps, args.cpu_mcmc_steps), tuple(args.mcmc_steps)) for t in range(P)]
    t0 = time.perf_counter()
    if P == 1:
        res = [_oracle_worker(jobs[0])]
    else:
        with mp.get_context('fork').Pool(P) as pool:
            res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    per_proc = [sum(r[0]) + sum(r[1]) for r in res]
    total = max(per_proc)
    value = P * args.steps / total
    sample = ('%d parallel single-thread processes (one BO trial each: -- TRIALS --; as the reference pins BLAS to 1 '
              'thread); per step: fit+weights at full size (PAIRLOOPS pair loops), EI+sort on the first %d of '
              '%d pool rows scaled x%.0f; target n=%d..%d'
              % (P, args.cpu_sample_rows, args.pool, args.pool / args.cpu_sample_rows, n_start + args.warmup,
                 n_start + args.warmup + args.steps - 1)).replace('PAIRLOOPS', args.pair_loops)
    sample = sample.replace('-- TRIALS --', 'distinct tasks / seeds' if args.distinct_trials else 'replicas of trial 0')
    sample += _mcmc_sample_note(args)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, 'cpu'),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': P, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': wall,
        'per_trial_ms_per_step': 1e3 * float(np.mean(per_proc)) / args.steps,
        'train_s_per_step': float(np.mean([np.mean(r[0]) for r in res])),
        'maximize_s_per_step_scaled': float(np.mean([np.mean(r[1]) for r in res])),
        'pair_loops': args.pair_loops,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, where):
    return {
        'workload': '%s offline BO, %d source %s x 50 obs, %d-candidate EI pool, d=9 random_forest space, '
                    'target n=%d..%d' % (args.method, args.sources,
                                         'GPs' if args.surrogate == 'gp' else 'MCMC-marginalised GPs (34 walkers)',
                                         args.pool, args.n0 + args.warmup, args.n0 + args.warmup + args.steps - 1),
        'method': args.method, 'sources': args.sources, 'pool': args.pool, 'd': 9,
        'surrogate': args.surrogate,
        'restarts_per_fit': 11 if args.surrogate == 'gp' else None,
        'fits_per_iter': 6 if args.method in ('rgpe', 'trans_rgpe', 'topo', 'topo_v3', 'obtl') else 1,
        'mcmc': (None if args.surrogate == 'gp' else
                 {'walkers': 34, 'burnin_steps': args.mcmc_steps[0], 'chain_steps': args.mcmc_steps[1],
                  'log_prob_evals_per_fit': 34 * (sum(args.mcmc_steps) + 2)}),
        'l2': ('flushed before every step (256 MiB device memset inside the timed region)' if where == 'gpu'
               else 'n/a (host)'),
        'multi_gpu': ('independent BO trials per rank, no data-path collective; ' +
                      ('distinct target task / seed per rank' if getattr(args, 'distinct_trials', False)
                       else 'every rank runs a replica of trial 0 (identical per-GPU work)')),
        'resident': 'candidate pool, fitted surrogates and the source surrogates\' pool posteriors stay in HBM',
    }


# ---------------------------------------------------------------------------------------------
def make_product(args, wl, shard=None):
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.pogpe import POGPE
    from tlbo_b200.facade.rgpe import RGPE, TransBO_RGPE
    from tlbo_b200.facade.topo_variant1 import OBTL, OBTLV, TOPO_V3
    from tlbo_b200.facade.tst import TST
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    cls = dict(rgpe=RGPE, trans_rgpe=TransBO_RGPE, topo=OBTLV, topo_v3=TOPO_V3, obtl=OBTL, tst=TST, pogpe=POGPE)[args.method]
    cs = TableSpace(wl['types'])
    kw = {}
    if args.surrogate == 'gp_mcmc':
        kw['mcmc_steps'] = tuple(args.mcmc_steps)
    fac = cls(cs, wl['sources'], None, wl['seed'], surrogate_type=args.surrogate, num_src_hpo_trial=50, **kw)
    smbo = SMBO_OFFLINE((wl['Xt'], wl['yt']), cs, fac, random_seed=wl['seed'], max_runs=10 ** 9, shard=shard)
    smbo.configurations = list(wl['rows0'])
    smbo.perfs = [float(wl['yt'][r]) for r in wl['rows0']]
    return fac, smbo


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tlbo_b200 import ops

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py (impl=ours) needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = x
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    per_rank = []
    # N > 1: independent BO trials, one per rank, no data-path collective.  By default every rank runs a
    # replica of trial 0 so that the per-GPU work is IDENTICAL to the N = 1 run (weak scaling then
    # measures the system, not the trial-to-trial spread of L-BFGS-B / SLSQP iteration counts, which is
    # +-15 % and would be charged to the slowest rank); --distinct-trials gives each rank its own task/seed.
    trial = rank if args.distinct_trials else 0
    wl = build_workload(trial % 30, 4465 + trial, args.pool, args.sources, args.n0)
    fac, smbo = make_product(args, wl)
    d_cont = int((wl['types'] == 0).sum())
    d_cat = int((wl['types'] != 0).sum())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dmma_tf = ops.fp64_peak_probe()
    K, W = args.steps, args.warmup

    def step(e2e):
        flush.zero_()
        if e2e:
            smbo.acq_optimizer.reupload()
        cfg, _, _, _ = smbo.iterate()
        return cfg.index

    # ---- warm-up (the clock sampler runs from here to the end of the last timed region)
    # one sampler for the whole job (rank 0's GPU): eight nvidia-smi pollers contend for the driver
    # lock with the ranks' own launches and synchronisations
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(W):
        step(False)
    snap = (list(smbo.configurations), list(smbo.perfs))

    def timed(e2e, record_flops):
        smbo.configurations, smbo.perfs = list(snap[0]), list(snap[1])
        ops.STATS.reset()
        ops.STATS.timing = record_flops
        flops = 0.0
        cached_reads = 0.0
        fitfl, fitev = 0.0, 0
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            step(e2e)
            if record_flops:
                n_t = len(smbo.configurations) - 1
                skip = list(getattr(fac, 'ignored_flag', [False] * fac.K))
                n_src_used = sum(1 for k in range(fac.K) if not skip[k])
                cached = bool(getattr(fac, '_pool_cache', None))
                ns = ([] if cached else [50] * n_src_used) + [n_t]
                if args.surrogate == 'gp_mcmc':
                    ns = [n_t] * fac._G          # the target's walker GPs (tlbo_predict_pool_posteriors)
                flops += fused_flops_per_candidate(ns, d_cont, d_cat) * args.pool
                cached_reads += (16.0 * n_src_used * args.pool) if cached else 0.0
                ff, fe = fit_flops(getattr(fac, 'last_fit_models', []), d_cont, d_cat)
                fitfl += ff
                fitev += fe
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        barrier()
        mine = e0.elapsed_time(e1)
        ms = max_over_ranks(mine)
        per_rank.append(all_ranks(mine))
        ops.STATS.timing = False
        return ms, wall, (flops, fitfl, fitev, cached_reads)

    # (1) headline region: no instrumentation; (2) the same K steps again with a CUDA-event pair
    # around every C-ABI call (kernel durations for the roofline / step breakdown)
    ms_res, wall_res, _ = timed(False, False)
    launches = ops.STATS.total_launches()
    launch_detail = dict(ops.STATS.launches)
    ms_ins, _, (flops, fitfl, fitev, cached_reads) = timed(False, True)
    kms = ops.STATS.kernel_ms()
    ms_e2e, wall_e2e, _ = timed(True, False)
    clocks = sampler.stop()
    h2d_step = ops.STATS.h2d / float(K)
    d2h_step = ops.STATS.d2h / float(K)

    # ---- pool-sharded variant of ONE trial (all ranks fit redundantly, each scores a row
    # range, NCCL all-gather of the per-rank arg-max)
    pool_sharded = None
    if world > 1 and not args.no_pool_sharded:
        from tlbo_b200.distributed import PoolShard
        wl0 = build_workload(0, 4465, args.pool * world, args.sources, args.n0)
        shard = PoolShard()
        fac_s, smbo_s = make_product(args, wl0, shard=shard)
        for _ in range(W):
            smbo_s.iterate()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            smbo_s.iterate()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        ms_s = max_over_ranks(e0.elapsed_time(e1))
        pool_sharded = {'value': K / (ms_s * 1e-3), 'unit': UNIT, 'pool_rows_total': args.pool * world,
                        'collective': 'all_gather of 4 doubles per rank (NCCL)', 'ms_per_step': ms_s / K,
                        'rows_selected': smbo_s.configurations[-3:]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the fused predict + EI + arg-max kernel
    fused_name = 'predict_ei_fused' if args.surrogate == 'gp' else 'predict_pool_posteriors'
    fused_calls, fused_ms = kms.get(fused_name, (0, 0.0))
    peaks = {}
    pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk_path):
        with open(pk_path) as f:
            peaks = json.load(f)
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    hbm_src = 'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    traffic = None
    tr_path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tr_path):
        with open(tr_path) as f:
            traffic = json.load(f).get('predict_ei_fused_dram_bytes_per_launch')
    ach_tf = flops / (fused_ms * 1e-3) / 1e12 if fused_ms > 0 else 0.0
    alg_bytes = args.pool * (8.0 * 9 + 8.0) + cached_reads / max(fused_calls, 1)
    ach_gbs = alg_bytes * fused_calls / (fused_ms * 1e-3) / 1e9 if fused_ms > 0 else 0.0
    roofline = {
        'kernel': 'predict_ei_fused_kernel', 'bound': 'tensor', 'achieved': ach_tf, 'peak': dmma_tf[1],
        'unit': 'TFLOP/s', 'frac': ach_tf / dmma_tf[1] if dmma_tf[1] else None, 'traffic': traffic,
        'peak_source': 'FP64 DMMA (mma.m8n8k4.f64) rate measured by tlbo_fp64_peak_probe in this run; '
                       'MEASURED_PEAKS.json carries no FP64 figure. FP64 FMA-pipe rate: %.2f TFLOP/s' % dmma_tf[0],
        'algorithmic_flops_per_launch': flops / max(fused_calls, 1),
        'algorithmic_bytes_per_launch': alg_bytes, 'avg_launch_ms': fused_ms / max(fused_calls, 1),
        'hbm_achieved_gbs': ach_gbs, 'hbm_peak_gbs': hbm_peak, 'hbm_frac': ach_gbs / hbm_peak,
        'hbm_peak_source': hbm_src,
        'note': 'flops by SURVEY.md 8(d) (triangular L^-1 form) over the surrogates actually EVALUATED per launch; '
                'the source surrogates\' posteriors over the fixed pool are computed once at maximiser construction '
                'and read back from HBM (16 B per candidate and source, counted in algorithmic_bytes_per_launch); '
                'arithmetic intensity ~%.0f flop/B. With every surrogate evaluated (no resident posteriors) the '
                'same kernel sustains 17-28%% of the FP64 peak: profiles/r1_predict_sweep.json' % (
                    flops / max(fused_calls, 1) / alg_bytes),
    }
    fit_calls, fit_ms = kms.get('gp_fit', (0, 0.0))
    # launches on side streams that run UNDER other kernels (the look-ahead tie-break key generator: one CTA,
    # enqueued behind the previous fused kernel and finished long before the next one needs the keys) are
    # reported but are not on the critical path of the step
    overlapped = {'mt19937_random_sample'}
    on_path = {k: v for k, v in kms.items() if k not in overlapped}
    dominant = max(on_path.items(), key=lambda kv: kv[1][1])[0] if on_path else None
    roofline_fit = {
        'kernel': 'gp_fit_kernel (+ gp_factorize_kernel)', 'bound': 'latency',
        'achieved': fitfl / (fit_ms * 1e-3) / 1e12 if fit_ms > 0 else 0.0, 'peak': dmma_tf[0], 'unit': 'TFLOP/s',
        'frac': (fitfl / (fit_ms * 1e-3) / 1e12 / dmma_tf[0]) if fit_ms > 0 else None,
        'nll_evaluations_per_launch': fitev / max(fit_calls, 1), 'avg_launch_ms': fit_ms / max(fit_calls, 1),
        'us_per_sequential_evaluation_note': 'one CTA per (surrogate, restart) runs its L-BFGS-B evaluations '
        'back to back; the launch lasts as long as the slowest restart (see DESIGN.md 3.1, tools/phase_fit.py)',
        'peak_source': 'FP64 FMA-pipe rate measured by tlbo_fp64_peak_probe in this run',
    }
    roofline_mcmc = None
    if args.surrogate == 'gp_mcmc':
        ch_calls, ch_ms = kms.get('gp_mcmc_chain', (0, 0.0))
        half = 2 * sum(args.mcmc_steps) + 2
        roofline_mcmc = {
            'kernel': 'gp_mcmc_kernel (persistent cooperative launch, one CTA per (surrogate, active walker))',
            'bound': 'latency', 'avg_launch_ms': ch_ms / max(ch_calls, 1),
            'sequential_log_prob_evaluations_per_launch': half,
            'us_per_half_step': 1e3 * ch_ms / max(ch_calls, 1) / half,
            'log_prob_evaluations_per_launch': 6 * fac._G * (sum(args.mcmc_steps) + 2),
            'note': 'one half-step = one kernel-matrix build + register-panel Cholesky + barrier among the '
                    'W/2 CTAs of a surrogate; SURVEY 8(d): latency-bound, report time not a roofline fraction',
        }
    step_ms = ms_res / K
    breakdown = {k: {'calls_per_step': c / float(K), 'ms_per_step': t / K, 'share_of_step': t / ms_ins}
                 for k, (c, t) in kms.items()}
    for k in overlapped:
        if k in breakdown:
            breakdown[k]['overlapped'] = ('side stream, concurrent with the host post-processing and the next fit kernel; '
                                          'the event pair also spans its wait for SM resources behind the previous '
                                          'fused kernel -- not part of the critical path')
    line = {
        'metric': METRIC, 'value': world * K / (ms_res * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, 'gpu'),
        'e2e': {'value': world * K / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d_step,
                'd2h_bytes_per_step': d2h_step, 'ms_per_step': ms_e2e / K,
                'api': 'SMBO_OFFLINE.iterate() with the candidate table re-uploaded from pinned host memory '
                       'every step and the selected row read back'},
        'gpu_launches': launches, 'gpu_launch_detail': launch_detail,
        'clocks': clocks, 'roofline': roofline, 'roofline_fit': roofline_fit, 'dominant_kernel': dominant,
        'step_breakdown': breakdown,
        'predict_candidates_per_s': args.pool * fused_calls / (fused_ms * 1e-3) if fused_ms > 0 else None,
        'host_wall_ms_per_step': 1e3 * wall_res / K, 'instrumented_ms_per_step': ms_ins / K,
        'per_rank_ms_per_step': [round(v / K, 4) for v in per_rank[0]],
    }
    if roofline_mcmc is not None:
        line['roofline_mcmc'] = roofline_mcmc
        line['roofline']['kernel'] = 'predict_ei_fused_kernel (per-walker posterior mode: tlbo_predict_pool_posteriors)'
    if pool_sharded is not None:
        line['pool_sharded'] = pool_sharded
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline_sample(args, args.n0 + W)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
