#!/usr/bin/env python
"""Headline benchmark of the surrogate-and-acquisition hot path.

Metric (BASELINE.json): BO iterations / second, one iteration = fit (target + 5 CV GPs, 11
L-BFGS-B restarts each) + ensemble weights (RGPE bootstrap ranking loss or TOPO SLSQP) +
fused predict / EI / arg-max over the whole candidate pool, at 29 source surrogates.
Workload = BASELINE.json configs[1]: 29 source GPs (50 observations each), 10^5-candidate
pool, random_forest-shaped space (d = 9), synthetic histories.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

Prints ONE JSON line (rank 0).  The headline fields are the RGPE run on configs[1]; the rest of
BASELINE.json's metric rides in the same line under "sub" (each record with its own value / e2e /
roofline / cpu_baseline): "topo" (OBTLV, what --methods topo runs), "pool_1e6" (the north-star pool
size), "predict_sweep" (configs[3]: 30 surrogates x n x N x d, nothing resident), "pool_strong" (ONE
trial over a FIXED 1e7-row pool sharded over the N ranks: strong scaling, NCCL all-gather of 4 doubles per
rank), "online" (configs[2]: 50 online-scenario trials sharded over the ranks), and for N > 1 "distinct_trials" (every rank its own task / seed, slowest rank charged) and --
at N = 8 -- "gp_mcmc" (configs[4]: TOPO over MCMC-marginalised GPs).  N > 1: one process per GPU
(torchrun); the headline shards by trial with no data-path collective ("weak" scaling).
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
    ap.add_argument('--cpu-sample-rows', type=int, default=0,
                    help='reference arm: score this many pool rows per step and scale (0 = the full pool, measured)')
    ap.add_argument('--passes', type=int, default=5, help='timed passes of exactly --steps steps each (median reported)')
    ap.add_argument('--sub', default='auto', help="sub-records: 'auto', 'none' or a comma list of topo,pool_1e6,predict_sweep,"
                    "pool_strong,online,distinct_trials,trials_per_gpu,gp_mcmc")
    ap.add_argument('--trials-per-gpu', type=int, default=4, help='concurrent trials (threads) of the trials_per_gpu record')
    ap.add_argument('--strong-pool', type=int, default=10000000, help='rows of the fixed pool of the pool_strong record')
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
    if pool <= 200000:
        yt = fam.evaluate(task, Xt, rng)
    else:
        # chunked: the random-feature map is a [rows, 64] temporary
        yt = np.concatenate([fam.evaluate(task, Xt[i:i + 200000], rng) for i in range(0, pool, 200000)])
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
    """Runs warmup + steps BO iterations of the oracle.  Per step: the full fit + weights (timed) and
    EI + ordering over the first ``sample_rows`` pool rows (timed; ``sample_rows`` = 0 or >= pool: the WHOLE
    pool, measured; otherwise scaled by pool / sample_rows -- the work is linear in the number of candidates).
    Returns (train_s[], maximize_s[], setup_s)."""
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
    if not sample_rows or sample_rows >= pool:
        sample_rows = pool
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


def _maximise_note(sample_rows, pool):
    if not sample_rows or sample_rows >= pool:
        return 'EI + sort over ALL %d pool rows (measured)' % pool
    return 'EI + sort on the first %d of %d pool rows scaled x%.0f' % (sample_rows, pool, pool / float(sample_rows))


def cpu_baseline_sample(args, n_start, method=None, pool=None, surrogate=None):
    """One reference-faithful oracle iteration on this box's host (1 core, BLAS pinned to one thread as the
    reference mandates), bounded sample.  Runs in a child process so the thread pins take effect before
    numpy loads."""
    method = method or args.method
    pool = pool or args.pool
    surrogate = surrogate or args.surrogate
    # bounded: the maximise leg is measured on the whole pool up to 2e5 rows, scaled beyond
    sample_rows = args.cpu_sample_rows if args.cpu_sample_rows else (0 if pool <= 200000 else 100000)

    def child(loops):
        cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--ref-procs', '1', '--steps', '1',
               '--warmup', '0', '--n0', str(n_start), '--method', method, '--pool', str(pool),
               '--sources', str(args.sources), '--cpu-sample-rows', str(sample_rows),
               '--pair-loops', loops, '--surrogate', surrogate, '--mcmc-steps', str(args.mcmc_steps[0]),
               str(args.mcmc_steps[1]), '--cpu-mcmc-steps', str(args.cpu_mcmc_steps)]
        env = dict(os.environ)
        for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
            env.pop(k, None)
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=1800)
        if out.returncode != 0:
            raise RuntimeError('cpu baseline child failed: ' + out.stderr[-2000:])
        return json.loads(out.stdout.strip().splitlines()[-1])
    lit = child('literal')
    t = lit['train_s_per_step'] + lit['maximize_s_per_step_scaled']
    rec = dict(value=1.0 / t, unit=UNIT, cores=1, kind='port',
               sample=('1 BO iteration of the numpy/scipy oracle at target n=%d, K=%d: fit+weights at full size '
                       '(literal Python pair loops as the reference), %s' % (n_start, args.sources,
                                                                             _maximise_note(sample_rows, pool)))
               + _mcmc_sample_note(args, surrogate),
               train_s=lit['train_s_per_step'], maximize_s=lit['maximize_s_per_step_scaled'],
               host_cores_available=os.cpu_count())
    if method in ('rgpe', 'trans_rgpe') and surrogate == 'gp':
        vec = child('vectorised')
        rec['value_vectorised_pair_count'] = 1.0 / (vec['train_s_per_step'] + vec['maximize_s_per_step_scaled'])
    return rec


def _mcmc_sample_note(args, surrogate=None):
    if (surrogate or args.surrogate) != 'gp_mcmc':
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
              'thread); per step: fit+weights at full size (PAIRLOOPS pair loops), %s; target n=%d..%d'
              % (P, _maximise_note(args.cpu_sample_rows, args.pool), n_start + args.warmup,
                 n_start + args.warmup + args.steps - 1)).replace('PAIRLOOPS', args.pair_loops)
    sample = sample.replace('-- TRIALS --', 'distinct tasks / seeds' if args.distinct_trials else 'replicas of trial 0')
    sample += _mcmc_sample_note(args)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': P, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': wall,
        'per_trial_ms_per_step': 1e3 * float(np.mean(per_proc)) / args.steps,
        'train_s_per_step': float(np.mean([np.mean(r[0]) for r in res])),
        'maximize_s_per_step_scaled': float(np.mean([np.mean(r[1]) for r in res])),
        'pair_loops': args.pair_loops,
    }
    emit(line)


def workload_config(args, method=None, pool=None, surrogate=None, distinct=None):
    """The workload description -- IDENTICAL in both arms (the driver compares the dicts)."""
    method = method or args.method
    pool = pool or args.pool
    surrogate = surrogate or args.surrogate
    distinct = getattr(args, 'distinct_trials', False) if distinct is None else distinct
    return {
        'workload': '%s offline BO, %d source %s x 50 obs, %d-candidate EI pool, d=9 random_forest space, '
                    'target n=%d..%d' % (method, args.sources,
                                         'GPs' if surrogate == 'gp' else 'MCMC-marginalised GPs (34 walkers)',
                                         pool, args.n0 + args.warmup, args.n0 + args.warmup + args.steps - 1),
        'method': method, 'sources': args.sources, 'pool': pool, 'd': 9,
        'surrogate': surrogate,
        'restarts_per_fit': 11 if surrogate == 'gp' else None,
        'fits_per_iter': 6 if method in ('rgpe', 'trans_rgpe', 'topo', 'topo_v3', 'obtl') else 1,
        'mcmc': (None if surrogate == 'gp' else
                 {'walkers': 34, 'burnin_steps': args.mcmc_steps[0], 'chain_steps': args.mcmc_steps[1],
                  'log_prob_evals_per_fit': 34 * (sum(args.mcmc_steps) + 2)}),
        'l2': 'CUDA arm: flushed before every step (256 MiB device memset inside the timed region); CPU arm: n/a (host)',
        'multi_gpu': ('independent BO trials per rank, no data-path collective; ' +
                      ('distinct target task / seed per rank' if distinct
                       else 'every rank runs a replica of trial 0 (identical per-GPU work)')),
        'resident': 'CUDA arm: candidate pool, fitted surrogates and the source surrogates\' pool posteriors stay in HBM',
    }


# ---------------------------------------------------------------------------------------------
def make_product(args, wl, shard=None, method=None, surrogate=None):
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.pogpe import POGPE
    from tlbo_b200.facade.rgpe import RGPE, TransBO_RGPE
    from tlbo_b200.facade.topo_variant1 import OBTL, OBTLV, TOPO_V3
    from tlbo_b200.facade.tst import TST
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    method = method or args.method
    surrogate = surrogate or args.surrogate
    cls = dict(rgpe=RGPE, trans_rgpe=TransBO_RGPE, topo=OBTLV, topo_v3=TOPO_V3, obtl=OBTL, tst=TST, pogpe=POGPE)[method]
    cs = TableSpace(wl['types'])
    kw = {}
    if surrogate == 'gp_mcmc':
        kw['mcmc_steps'] = tuple(args.mcmc_steps)
    fac = cls(cs, wl['sources'], None, wl['seed'], surrogate_type=surrogate, num_src_hpo_trial=50, **kw)
    smbo = SMBO_OFFLINE((wl['Xt'], wl['yt']), cs, fac, random_seed=wl['seed'], max_runs=10 ** 9, shard=shard)
    smbo.configurations = list(wl['rows0'])
    smbo.perfs = [float(wl['yt'][r]) for r in wl['rows0']]
    return fac, smbo


class Comm(object):
    """Process-group helpers: barrier + synchronize, max / all-gather of a scalar over the ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py (impl=ours) needs a CUDA device; there is no CPU fallback')
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(self, x):
        if self.world == 1:
            return [x]
        t = self.torch.zeros(self.world, dtype=self.torch.float64, device=self.dev)
        t[self.rank] = x
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _peaks():
    peaks = {}
    pk_path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk_path):
        with open(pk_path) as f:
            peaks = json.load(f)
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    hbm_src = 'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    return hbm_peak, hbm_src


def measure_offline(args, comm, flush, dmma_tf, method, pool, surrogate='gp', distinct=False, passes=1,
                    want_all_surrogates=False):
    """One offline-BO measurement: W warm-up steps, `passes` un-instrumented passes of EXACTLY K steps each
    over the same trajectory (median reported; every pass bracketed by barrier + synchronize, max over
    ranks), one pass with a CUDA-event pair around every C-ABI call (roofline / step breakdown), `passes`
    end-to-end passes (candidate table re-uploaded from pinned host memory every step).  Returns the record
    (every rank) -- the caller prints rank 0's."""
    import torch
    from tlbo_b200 import ops
    K, W = args.steps, args.warmup
    world = comm.world
    trial = comm.rank if distinct else 0
    wl = build_workload(trial % 30, 4465 + trial, pool, args.sources, args.n0)
    fac, smbo = make_product(args, wl, method=method, surrogate=surrogate)
    d_cont = int((wl['types'] == 0).sum())
    d_cat = int((wl['types'] != 0).sum())

    def step(e2e):
        flush.zero_()
        if e2e:
            smbo.acq_optimizer.reupload()
        cfg, _, _, _ = smbo.iterate()
        return cfg.index

    for _ in range(W):
        step(False)
    snap = (list(smbo.configurations), list(smbo.perfs))
    per_rank = []

    def timed(e2e, record_flops):
        smbo.configurations, smbo.perfs = list(snap[0]), list(snap[1])
        ops.STATS.reset()
        ops.STATS.timing = record_flops
        flops = 0.0
        cached_reads = 0.0
        fitfl, fitev = 0.0, 0
        comm.barrier()
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
                if surrogate == 'gp_mcmc':
                    ns = [n_t] * fac._G          # the target's walker GPs (tlbo_predict_pool_posteriors)
                flops += fused_flops_per_candidate(ns, d_cont, d_cat) * pool
                cached_reads += (16.0 * n_src_used * pool) if cached else 0.0
                ff, fe = fit_flops(getattr(fac, 'last_fit_models', []), d_cont, d_cat)
                fitfl += ff
                fitev += fe
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        comm.barrier()
        mine = e0.elapsed_time(e1)
        ms = comm.max_over_ranks(mine)
        per_rank.append(comm.all_ranks(mine))
        ops.STATS.timing = False
        return ms, wall, (flops, fitfl, fitev, cached_reads)

    res_passes = [timed(False, False) for _ in range(passes)]
    launches = ops.STATS.total_launches()
    launch_detail = dict(ops.STATS.launches)
    ms_ins, _, (flops, fitfl, fitev, cached_reads) = timed(False, True)
    kms = ops.STATS.kernel_ms()
    e2e_passes = [timed(True, False) for _ in range(passes)]
    h2d_step = ops.STATS.h2d / float(K)
    d2h_step = ops.STATS.d2h / float(K)
    order = sorted(range(passes), key=lambda i: res_passes[i][0])
    mid = order[(passes - 1) // 2]              # (lower) median pass
    ms_res, wall_res = res_passes[mid][0], res_passes[mid][1]
    ms_e2e = sorted(p[0] for p in e2e_passes)[(passes - 1) // 2]

    # ---- all K + 1 surrogates evaluated over the pool, nothing resident (the reference's predict shape):
    # the "GP predict candidates / sec" half of the metric on THIS workload
    all_sur = None
    if want_all_surrogates and surrogate == 'gp':
        acq = smbo.acq_optimizer
        saved = fac._pool_cache
        saved_flags = getattr(fac, 'ignored_flag', None)
        fac._pool_cache = {}
        if saved_flags is not None:
            fac.ignored_flag = [False] * fac.K       # weight-dilution pruning off: EVERY surrogate is evaluated
        try:
            keys = acq._keys_full[acq._key_cur][acq.lo:acq.hi] if acq._device_keys else acq._keys_dev
            for _ in range(2):
                smbo.acquisition_function.best_unevaluated(acq.X_dev, keys=keys, excluded=None, ws=acq._ws, sync=True)
            reps = 5
            comm.barrier()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                flush.zero_()
                smbo.acquisition_function.best_unevaluated(acq.X_dev, keys=keys, excluded=None, ws=acq._ws, sync=False)
            e1.record()
            torch.cuda.synchronize()
            e2 = torch.cuda.Event(enable_timing=True)
            e3 = torch.cuda.Event(enable_timing=True)
            e2.record()
            for _ in range(reps):
                flush.zero_()
            e3.record()
            torch.cuda.synchronize()
            ms_all = (e0.elapsed_time(e1) - e2.elapsed_time(e3)) / reps
        finally:
            fac._pool_cache = saved
            if saved_flags is not None:
                fac.ignored_flag = saved_flags
        n_t = len(smbo.configurations)
        n_used = fac.K
        fl = fused_flops_per_candidate([50] * n_used + [n_t], d_cont, d_cat) * pool
        all_sur = {'surrogates_evaluated': n_used + 1, 'ms_per_launch': ms_all,
                   'candidates_per_s': pool / (ms_all * 1e-3), 'tflops_algorithmic': fl / (ms_all * 1e-3) / 1e12,
                   'frac_of_dmma_peak': fl / (ms_all * 1e-3) / 1e12 / dmma_tf[1],
                   'note': 'fused predict + EI + arg-max with ALL K source surrogates (weight-dilution pruning off) AND '
                           'the target evaluated per candidate, no resident posteriors, L2 flushed between launches '
                           '(flush time subtracted)'}

    # ---- roofline of the fused predict + EI + arg-max kernel
    fused_name = 'predict_ei_fused' if surrogate == 'gp' else 'predict_pool_posteriors'
    fused_calls, fused_ms = kms.get(fused_name, (0, 0.0))
    hbm_peak, hbm_src = _peaks()
    traffic = None
    tr_path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tr_path) and pool == 100000 and method == 'rgpe':
        with open(tr_path) as f:
            traffic = json.load(f).get('predict_ei_fused_dram_bytes_per_launch')
    ach_tf = flops / (fused_ms * 1e-3) / 1e12 if fused_ms > 0 else 0.0
    alg_bytes = pool * (8.0 * 9 + 8.0) + cached_reads / max(fused_calls, 1)
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
                'arithmetic intensity ~%.0f flop/B. The all-surrogates figure of the same kernel is in '
                '"predict_all_surrogates" and the "predict_sweep" sub-record' % (
                    flops / max(fused_calls, 1) / max(alg_bytes, 1.0)),
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
    step_ms = ms_res / K
    breakdown = {k: {'calls_per_step': c / float(K), 'ms_per_step': t / K, 'share_of_step': t / ms_ins}
                 for k, (c, t) in kms.items()}
    for k in overlapped:
        if k in breakdown:
            breakdown[k]['overlapped'] = ('side stream, concurrent with the host post-processing and the next fit kernel; '
                                          'the event pair also spans its wait for SM resources behind the previous '
                                          'fused kernel -- not part of the critical path')
    rec = {
        'metric': METRIC, 'value': world * K / (ms_res * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': K,
        'warmup': W, 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, method=method, pool=pool, surrogate=surrogate, distinct=distinct),
        'e2e': {'value': world * K / (ms_e2e * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d_step,
                'd2h_bytes_per_step': d2h_step, 'ms_per_step': ms_e2e / K,
                'api': 'SMBO_OFFLINE.iterate() with the candidate table re-uploaded from pinned host memory '
                       'every step and the selected row read back'},
        'gpu_launches': launches, 'gpu_launch_detail': launch_detail,
        'timed_passes': {'passes': passes, 'steps_per_pass': K, 'statistic': 'median pass (max over ranks per pass)',
                         'ms_per_step': [round(p[0] / K, 4) for p in res_passes],
                         'e2e_ms_per_step': [round(p[0] / K, 4) for p in e2e_passes],
                         'timed_region_s_total': sum(p[0] for p in res_passes) * 1e-3},
        'roofline': roofline, 'roofline_fit': roofline_fit, 'dominant_kernel': dominant,
        'step_breakdown': breakdown,
        'host_wall_ms_per_step': 1e3 * wall_res / K, 'instrumented_ms_per_step': ms_ins / K,
        'per_rank_ms_per_step': [round(v / K, 4) for v in per_rank[mid]],
    }
    if all_sur is not None:
        rec['predict_all_surrogates'] = all_sur
        rec['predict_candidates_per_s'] = all_sur['candidates_per_s']
    if surrogate == 'gp_mcmc':
        ch_calls, ch_ms = kms.get('gp_mcmc_chain', (0, 0.0))
        half = 2 * sum(args.mcmc_steps) + 2
        rec['roofline_mcmc'] = {
            'kernel': 'gp_mcmc_kernel (persistent cooperative launch, one CTA per (surrogate, active walker))',
            'bound': 'latency', 'avg_launch_ms': ch_ms / max(ch_calls, 1),
            'sequential_log_prob_evaluations_per_launch': half,
            'us_per_half_step': 1e3 * ch_ms / max(ch_calls, 1) / half,
            'log_prob_evaluations_per_launch': 6 * fac._G * (sum(args.mcmc_steps) + 2),
            'note': 'one half-step = one kernel-matrix build + register-panel Cholesky + barrier among the '
                    'W/2 CTAs of a surrogate; SURVEY 8(d): latency-bound, report time not a roofline fraction',
        }
        rec['roofline']['kernel'] = 'predict_ei_fused_kernel (per-walker posterior mode: tlbo_predict_pool_posteriors)'
    del fac, smbo
    torch.cuda.empty_cache()
    return rec


def measure_trials_per_gpu(args, comm, n_threads=2, method='rgpe'):
    """SEVERAL independent trials per GPU, concurrently: one thread + one CUDA stream + one private copy of the
    "process-global" np.random stream (tlbo_b200.utils.global_stream) per trial, so that the trials' fit kernels (66
    CTAs each) share the 148 SMs.  Each trial is a distinct task / seed and selects exactly the rows it selects when it
    runs alone (tests/test_gpu_bo_loop.py::test_two_trials_in_one_process_...).  W warm-up steps per trial, then K
    timed steps per trial between two device-wide synchronisations, every step behind a 256 MiB L2 flush on the
    trial's own stream; value = trials x K / time, summed over the ranks (slowest rank charged).  A second leg (e2e)
    re-uploads every trial's candidate table from pinned host memory in every step.  The single-trial step time of
    the same run is reported beside it."""
    import threading
    import torch
    from tlbo_b200.utils.global_stream import thread_stream, waiting
    K, W = args.steps, args.warmup
    T = int(n_threads)
    turns = os.environ.get('TLBO_HOST_TURNS', '1') != '0'
    trials = [comm.rank * T + t for t in range(T)]
    reps = 3
    legs = (False, True) * reps          # resident / e2e legs alternate; the median repetition is reported
    state = {}
    errors = []
    ready = threading.Barrier(T + 1)
    go = threading.Barrier(T + 1)
    done = threading.Barrier(T + 1)

    def run(t, trial):
        try:
            torch.cuda.set_device(comm.local_rank)
            stream = torch.cuda.Stream()
            with thread_stream(turns=turns), torch.cuda.stream(stream):
                flush = torch.empty(256 << 20, dtype=torch.uint8, device=comm.dev)
                wl = build_workload(trial % 30, 4465 + trial, args.pool, args.sources, args.n0)
                fac, smbo = make_product(args, wl, method=method, surrogate='gp')
                for _ in range(W):
                    flush.zero_()
                    smbo.iterate()
                snap = (list(smbo.configurations), list(smbo.perfs))
                for leg, e2e in enumerate(legs):
                    smbo.configurations, smbo.perfs = list(snap[0]), list(snap[1])
                    with waiting():
                        stream.synchronize()
                        ready.wait()
                        go.wait()
                    e0 = torch.cuda.Event(enable_timing=True)
                    e1 = torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(K):
                        flush.zero_()
                        if e2e:
                            smbo.acq_optimizer.reupload()
                        smbo.iterate()
                    e1.record(stream)
                    with waiting():
                        stream.synchronize()
                    state.setdefault((t, e2e), []).append(e0.elapsed_time(e1))
                    with waiting():
                        done.wait()
        except Exception as ex:                               # pragma: no cover
            errors.append(repr(ex))
            for b in (ready, go, done):
                b.abort()

    # the trial threads hand the interpreter lock to each other at every C-ABI call; CPython's default 5 ms switch
    # interval would let a thread that is busy in pure Python hold the others off for two whole BO steps
    old_switch = sys.getswitchinterval()
    sys.setswitchinterval(float(os.environ.get('TLBO_SWITCH_INTERVAL', '1e-4')))
    threads = [threading.Thread(target=run, args=(t, tr), daemon=True) for t, tr in enumerate(trials)]
    for th in threads:
        th.start()
    ms = {}
    broken = False
    for e2e in legs:
        # every rank makes the SAME collective calls whatever happens in its trial threads (a failed thread must not
        # leave the other ranks waiting in a barrier)
        try:
            if not broken:
                ready.wait()
        except threading.BrokenBarrierError:
            broken = True
        comm.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        try:
            if not broken:
                go.wait()
                done.wait()
        except threading.BrokenBarrierError:
            broken = True
        torch.cuda.synchronize()
        ms.setdefault(e2e, []).append(comm.max_over_ranks((time.perf_counter() - t0) * 1e3))
    for th in threads:
        th.join(60)
    sys.setswitchinterval(old_switch)
    failed = comm.max_over_ranks(1.0 if (broken or errors) else 0.0)
    if failed:
        return {'metric': METRIC + ', %d concurrent trials per GPU' % T, 'value': None, 'unit': UNIT,
                'error': (errors[0] if errors else 'a trial thread failed on another rank')}
    pool_bytes = float(args.pool) * 9 * 8
    all_ms = {k: [round(x / (T * K), 4) for x in v] for k, v in ms.items()}
    pick = {k: sorted(range(reps), key=lambda i: v[i])[(reps - 1) // 2] for k, v in ms.items()}
    state = {(t, k): state[(t, k)][pick[k]] for t in range(T) for k in (False, True)}
    ms = {k: v[pick[k]] for k, v in ms.items()}
    return {'metric': METRIC + ', %d concurrent trials per GPU' % T, 'value': comm.world * T * K / (ms[False] * 1e-3),
            'unit': UNIT, 'n_gpus': comm.world, 'steps': K, 'warmup': W, 'trials_per_gpu': T,
            'ms_per_step_per_gpu': ms[False] / (T * K),
            'single_trial_ms_per_step_under_sharing': [state[(t, False)] / K for t in range(T)],
            'e2e': {'value': comm.world * T * K / (ms[True] * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': pool_bytes, 'ms_per_step_per_gpu': ms[True] / (T * K),
                    'single_trial_ms_per_step_under_sharing': [state[(t, True)] / K for t in range(T)],
                    'api': 'SMBO_OFFLINE.iterate() in every trial thread, that trial\'s candidate table re-uploaded '
                           'from pinned host memory every step and the selected row read back'},
            'host_turns': turns,
            'repetitions': {'n': reps, 'statistic': 'median repetition', 'ms_per_step_per_gpu': all_ms[False],
                            'e2e_ms_per_step_per_gpu': all_ms[True]},
            'higher_is_better': True, 'scaling': 'weak', 'dtype': 'f64', 'data': 'synthetic',
            'timing': 'wall clock between two device-wide synchronisations (the trials run on %d streams), max over '
                      'ranks; per-trial figures from CUDA events on each trial\'s stream' % T,
            'config': workload_config(args, method=method, pool=args.pool, surrogate='gp', distinct=True)}


def measure_pool_strong(args, comm, flush, pool_total):
    """ONE BO trial over a FIXED pool of `pool_total` rows sharded by rows over the ranks (strong scaling):
    every rank fits redundantly (deterministic), scores its row range, and the ranks all-gather the 32-byte
    (ei, key, row, #negative) quadruple and merge it on the device."""
    import torch
    from tlbo_b200.distributed import PoolShard
    K, W = args.steps, args.warmup
    wl0 = build_workload(0, 4465, pool_total, args.sources, args.n0)
    shard = PoolShard() if comm.world > 1 else None
    fac_s, smbo_s = make_product(args, wl0, shard=shard, method='rgpe', surrogate='gp')

    def step():
        flush.zero_()
        smbo_s.iterate()
    for _ in range(W):
        step()
    comm.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    comm.barrier()
    ms_s = comm.max_over_ranks(e0.elapsed_time(e1))
    rec = {'metric': METRIC, 'value': K / (ms_s * 1e-3), 'unit': UNIT, 'scaling': 'strong', 'n_gpus': comm.world,
           'pool_rows_total': pool_total, 'pool_rows_per_rank': (pool_total + comm.world - 1) // comm.world,
           'collective': ('all_gather of 4 doubles per rank (NCCL), merged on the device' if comm.world > 1 else 'none'),
           'ms_per_step': ms_s / K, 'rows_selected': [int(r) for r in smbo_s.configurations[-3:]],
           'note': 'fit + weights run redundantly on every rank (deterministic), so the strong-scaling limit of the '
                   'step is the fit; the tie-break keys rng.rand(N) of the WHOLE pool are generated on every rank '
                   '(MT19937 is sequential; see DESIGN.md)'}
    del fac_s, smbo_s
    torch.cuda.empty_cache()
    return rec


def _online_trial(K, t, n0, num_points):
    """One online-scenario trial (own target task / seed), built and advanced by ONE untimed iteration."""
    from tlbo_b200 import synthetic
    from tlbo_b200.config_space import Configuration, TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.framework.smbo_online import SMBO_ONLINE
    types, const = synthetic.space_types('rf')
    fam = synthetic.TaskFamily(types, const, n_tasks=30, seed=4466)
    task, seed = t % 30, 4465 + t
    ids = [i for i in range(30) if i != task]
    np.random.RandomState(seed).shuffle(ids)
    sources = []
    for k in ids[:K]:
        rng = np.random.RandomState(1000 + k)
        X = synthetic.sample_table(rng, 50, types, const)
        sources.append((X, fam.evaluate(k, X, rng)))

    def objective(config, task=task):
        x = config.get_array() if hasattr(config, 'get_array') else np.asarray(config)
        return float(fam.evaluate(task, x[np.newaxis, :])[0])
    cs = TableSpace(types, const_mask=const)
    cs.seed(seed)
    fac = RGPE(cs, sources, None, seed, surrogate_type='gp', num_src_hpo_trial=50)
    smbo = SMBO_ONLINE(objective, cs, fac, random_seed=seed, max_runs=10 ** 9, num_points=num_points)
    X0 = cs.sample_array(n0)
    for i in range(n0):
        c = Configuration(cs, X0[i], -1)
        smbo.configurations.append(c)
        smbo.perfs.append(objective(c))
        smbo.history_container.add(c, smbo.perfs[-1])
    smbo.iterate()
    return smbo


def measure_online(args, comm, n_trials=50, iters=4, n0=20, num_points=5000):
    """BASELINE configs[2]: the ONLINE scenario (tools/online_benchmark.py shape: SMBO_ONLINE with
    InterleavedLocalAndRandomSearch -- 10 hill-climbs + ~4990 random configurations sorted by EI -- and RGPE over
    29 sources), `n_trials` independent trials (own target task / seed) sharded round-robin over the ranks, `iters`
    timed BO iterations each after one untimed one.  No collective on the data path; the slowest rank is charged.
    Facade construction (the 29 source fits of a trial) is outside the timed loops and reported beside them.
    A second pass (``concurrent``) runs the rank's trials on --trials-per-gpu threads (a CUDA stream and a private copy
    of the global np.random stream each, as the trials_per_gpu record): wall clock over all timed loops."""
    import torch
    from tlbo_b200 import ops
    K = args.sources
    mine = [t for t in range(n_trials) if t % comm.world == comm.rank]
    timed_ms, build_s, launches = 0.0, 0.0, 0
    picked = {}
    comm.barrier()
    for t in mine:
        a = time.perf_counter()
        smbo = _online_trial(K, t, n0, num_points)
        torch.cuda.synchronize()
        build_s += time.perf_counter() - a
        ops.STATS.reset()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            smbo.iterate()
        e1.record()
        torch.cuda.synchronize()
        timed_ms += e0.elapsed_time(e1)
        launches += ops.STATS.total_launches()
        picked[t] = [tuple(np.asarray(c.get_array(), dtype=np.float64).tolist()) for c in smbo.configurations[-iters:]]
        del smbo
    comm.barrier()
    ms = comm.max_over_ranks(timed_ms)
    per_rank = comm.all_ranks(timed_ms)
    total_iters = n_trials * iters
    rec = {'metric': 'BO iters/sec, online scenario (fit + weights + interleaved local / random EI search) at %d sources' % K,
           'value': total_iters / (ms * 1e-3), 'unit': UNIT, 'n_gpus': comm.world, 'scaling': 'strong',
           'trials': n_trials, 'timed_iterations_per_trial': iters, 'ms_per_iteration_per_rank': ms / max(len(mine) * iters, 1),
           'per_rank_timed_ms': [round(v, 2) for v in per_rank], 'gpu_launches_rank0': launches,
           'construction_s_rank0': build_s,
           'config': {'workload': 'rgpe online BO (SMBO_ONLINE + InterleavedLocalAndRandomSearch, %d challengers per '
                                  'iteration), %d source GPs x 50 obs, d=9 random_forest space, %d trials (own task / seed) '
                                  'sharded over the ranks, target n=%d..%d' % (num_points, K, n_trials, n0 + 1, n0 + iters)},
           'note': 'host-bound: one hill-climb step = one neighbourhood launch + one synchronisation (DESIGN.md 3.5)'}
    T = int(getattr(args, 'trials_per_gpu', 0) or 0)
    if T > 1:
        rec['concurrent'] = _online_concurrent(comm, K, mine, T, iters, n0, num_points, total_iters, picked)
    return rec


def _online_concurrent(comm, K, mine, T, iters, n0, num_points, total_iters, picked):
    """The rank's online trials on T threads: every thread builds its trials (untimed, the private np.random state
    saved with each), then -- between two barriers -- runs their timed iterations.  Also checks that every trial
    selected the configurations of its sequential run."""
    import threading
    import torch
    from tlbo_b200.utils.global_stream import thread_stream
    errors, same = [], []
    ready = threading.Barrier(T + 1)
    done = threading.Barrier(T + 1)

    def run(w):
        try:
            torch.cuda.set_device(comm.local_rank)
            stream = torch.cuda.Stream()
            with thread_stream(), torch.cuda.stream(stream):
                built = []
                for t in mine[w::T]:
                    smbo = _online_trial(K, t, n0, num_points)
                    built.append((t, smbo, np.random.get_state()))
                stream.synchronize()
                ready.wait()
                for t, smbo, st in built:
                    np.random.set_state(st)
                    for _ in range(iters):
                        smbo.iterate()
                stream.synchronize()
                done.wait()
                for t, smbo, _ in built:
                    got = [tuple(np.asarray(c.get_array(), dtype=np.float64).tolist())
                           for c in smbo.configurations[-iters:]]
                    same.append(got == picked[t])
        except Exception as ex:                               # pragma: no cover
            errors.append(repr(ex))
            ready.abort()
            done.abort()

    threads = [threading.Thread(target=run, args=(w,), daemon=True) for w in range(T)]
    for th in threads:
        th.start()
    broken = False
    try:
        ready.wait()
    except threading.BrokenBarrierError:
        broken = True
    comm.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    try:
        if not broken:
            done.wait()
    except threading.BrokenBarrierError:
        broken = True
    torch.cuda.synchronize()
    ms = comm.max_over_ranks((time.perf_counter() - t0) * 1e3)
    for th in threads:
        th.join(60)
    if comm.max_over_ranks(1.0 if (broken or errors) else 0.0):
        return {'error': errors[0] if errors else 'a trial thread failed on another rank'}
    return {'value': total_iters / (ms * 1e-3), 'unit': UNIT, 'threads_per_gpu': T, 'wall_ms': ms,
            'trials_matching_their_sequential_run_rank0': '%d of %d' % (sum(same), len(same)),
            'timing': 'wall clock over the timed iterations of all trials of a rank (T threads, own streams), max over '
                      'ranks; construction outside'}


def predict_sweep(dmma_tf, max_rows):
    """BASELINE configs[3]: 30 surrogates x n in {50,100,200} x N candidates x d in {5,20}, every surrogate
    evaluated per candidate (nothing resident): candidates / s and the fraction of the FP64 DMMA peak."""
    import torch
    from tlbo_b200 import ops
    dev = torch.device('cuda')
    rows = []
    M = 30
    for d in (5, 20):
        types = np.zeros(d, int)
        for n in (50, 100, 200):
            rng = np.random.RandomState(100 * d + n)
            pack = ops.SurrogatePack(M, n, d)
            Xs, ys, ths, yms, yss = [], [], [], [], []
            for m in range(M):
                X = rng.rand(n, d)
                y = np.sin(X.sum(1) * (1 + 0.1 * m)) + 0.05 * rng.randn(n)
                th = np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.2, 0.05, d), [rng.uniform(-9, -5)]])
                Xs.append(X); ys.append((y - y.mean()) / y.std()); ths.append(th); yms.append(y.mean()); yss.append(y.std())
            st = ops.gp_factorize_batched(Xs, ys, types, ths, yms, yss, pack)
            assert not np.asarray(st).any()
            w = torch.full((M,), 1.0 / M, dtype=torch.float64, device=dev)
            flops_per_cand = 30.0 + M * (n * (3 * d + 12) + 2 * n + n * n + 2 * n)
            for N in (10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7):
                if N > max_rows:
                    continue
                Xc = torch.rand(N, d, dtype=torch.float64, device=dev)
                keys = torch.rand(N, dtype=torch.float64, device=dev)
                ws = ops.FusedWorkspace()
                ops.predict_ei_fused(pack, types, w, None, Xc, 0.1, keys=keys, ws=ws, nmax=n)
                reps = 2 if N >= 10 ** 7 else (3 if N >= 10 ** 6 else 10)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    ops.predict_ei_fused(pack, types, w, None, Xc, 0.1, keys=keys, ws=ws, nmax=n, sync=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                rows.append(dict(d=d, n=n, M=M, N=N, ms=ms, candidates_per_s=N / (ms * 1e-3),
                                 tflops_algorithmic=flops_per_cand * N / (ms * 1e-3) / 1e12,
                                 frac_of_dmma_peak=flops_per_cand * N / (ms * 1e-3) / 1e12 / dmma_tf[1],
                                 hbm_gbs_algorithmic=N * 8.0 * (d + 1) / (ms * 1e-3) / 1e9))
                del Xc, keys
            del pack
    torch.cuda.empty_cache()
    best = max(rows, key=lambda r: r['frac_of_dmma_peak'])
    return {'metric': 'GP predict candidates/sec (30 surrogates, fused mean / variance / EI / arg-max)',
            'unit': 'candidates/s', 'rows': rows, 'dmma_peak_tflops': dmma_tf[1],
            'best_frac_of_dmma_peak': best['frac_of_dmma_peak'],
            'flops': 'SURVEY.md 8(d), triangular form: 30 + M (n (3 d + 12) + 2 n + n^2 + 2 n) per candidate',
            'l2': 'pools >= 1e6 rows exceed L2 (126 MB); smaller pools are L2-resident between repetitions',
            'parity': 'tests/test_gpu_kernels.py and tools/predict_sweep.py check the same launches against the oracle'}


def sweep_cpu_baseline():
    """Oracle posterior over a bounded sample for one sweep cell (30 surrogates, n = 50, d = 5), one core."""
    from oracle import gp as ogp
    d, n, M, rows = 5, 50, 30, 2000
    types = np.zeros(d, int)
    spec = ogp.KernelSpec(types)
    rng = np.random.RandomState(100 * d + n)
    gs = []
    for m in range(M):
        X = rng.rand(n, d)
        y = np.sin(X.sum(1) * (1 + 0.1 * m)) + 0.05 * rng.randn(n)
        th = np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.2, 0.05, d), [rng.uniform(-9, -5)]])
        g = ogp.OracleGP(types, 1)
        g.X = X; g.y = (y - y.mean()) / y.std(); g.spec = spec; g.mean_y = y.mean(); g.std_y = y.std()
        g._fit_at(th); g.is_trained = True
        gs.append(g)
    Xq = rng.rand(rows, d)
    t0 = time.perf_counter()
    for g in gs:
        g.predict(Xq)
    dt = time.perf_counter() - t0
    return {'value': rows / dt, 'unit': 'candidates/s', 'cores': 1, 'kind': 'port',
            'sample': 'oracle GaussianProcess.predict (the reference\'s K_inv_ einsum form) of %d surrogates '
                      '(n = %d, d = %d) over %d candidates' % (M, n, d, rows)}


def run_ours(args):
    import torch
    from tlbo_b200 import ops
    comm = Comm()
    rank, world = comm.rank, comm.world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=comm.dev)
    dmma_tf = ops.fp64_peak_probe()
    if args.sub == 'auto':
        subs = ['pool_strong']
        if world == 1:
            subs = ['topo', 'pool_1e6', 'predict_sweep', 'pool_strong', 'online', 'trials_per_gpu']
        else:
            subs = ['distinct_trials', 'trials_per_gpu', 'pool_strong', 'online'] + (['gp_mcmc'] if world >= 8 else [])
    elif args.sub in ('none', ''):
        subs = []
    else:
        subs = [x.strip() for x in args.sub.split(',') if x.strip()]
    if args.method != 'rgpe' or args.surrogate != 'gp' or args.pool != 100000:
        subs = [x for x in subs if x in ('pool_strong', 'distinct_trials', 'online', 'trials_per_gpu')] if args.sub == 'auto' else subs

    # one clock sampler for the whole job (rank 0's GPU): eight nvidia-smi pollers contend for the driver
    # lock with the ranks' own launches and synchronisations
    sampler = ClockSampler(comm.local_rank)
    if rank == 0:
        sampler.start()
    line = measure_offline(args, comm, flush, dmma_tf, args.method, args.pool, args.surrogate,
                           distinct=args.distinct_trials, passes=max(1, args.passes), want_all_surrogates=True)
    sub = {}
    for name in subs:
        t0 = time.perf_counter()
        if name == 'topo':
            rec = measure_offline(args, comm, flush, dmma_tf, 'topo', args.pool, 'gp', passes=3)
        elif name == 'pool_1e6':
            rec = measure_offline(args, comm, flush, dmma_tf, 'rgpe', 1000000, 'gp', passes=3, want_all_surrogates=True)
        elif name == 'trials_per_gpu':
            rec = measure_trials_per_gpu(args, comm, args.trials_per_gpu)
        elif name == 'topo_trials_per_gpu':
            rec = measure_trials_per_gpu(args, comm, args.trials_per_gpu, method='topo')
        elif name == 'distinct_trials':
            rec = measure_offline(args, comm, flush, dmma_tf, args.method, args.pool, args.surrogate, distinct=True,
                                  passes=3)
            rec['note'] = ('every rank runs its OWN target task / seed; the slowest rank is charged (L-BFGS-B / SLSQP '
                           'iteration counts differ from trial to trial by about +-15 %)')
        elif name == 'gp_mcmc':
            rec = measure_offline(args, comm, flush, dmma_tf, 'topo', args.pool, 'gp_mcmc', distinct=True, passes=1)
        elif name == 'pool_strong':
            rec = measure_pool_strong(args, comm, flush, args.strong_pool)
        elif name == 'online':
            rec = measure_online(args, comm)
        elif name == 'predict_sweep':
            rec = predict_sweep(dmma_tf, 10 ** 7) if rank == 0 else None
            comm.barrier()
        else:
            continue
        if rec is not None:
            rec['record_wall_s'] = time.perf_counter() - t0
        sub[name] = rec
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        comm.close()
        return
    line['clocks'] = clocks
    if world == 1 and not args.no_cpu_baseline:
        n_start = args.n0 + args.warmup
        line['cpu_baseline'] = cpu_baseline_sample(args, n_start)
        if 'topo' in sub:
            sub['topo']['cpu_baseline'] = cpu_baseline_sample(args, n_start, method='topo')
        if 'pool_1e6' in sub:
            sub['pool_1e6']['cpu_baseline'] = cpu_baseline_sample(args, n_start, pool=1000000)
        if 'predict_sweep' in sub:
            sub['predict_sweep']['cpu_baseline'] = sweep_cpu_baseline()
    if sub:
        line['sub'] = sub
    tp = sub.get('trials_per_gpu')
    if tp and tp.get('value'):
        # per-GPU THROUGHPUT with several concurrent trials per GPU, next to the single-trial headline (latency)
        line['throughput_concurrent_trials'] = {
            'value': tp['value'], 'unit': UNIT, 'trials_per_gpu': tp['trials_per_gpu'], 'e2e_value': tp['e2e']['value'],
            'note': 'sub-record trials_per_gpu: the same BO steps, %d independent trials per GPU on their own threads '
                    'and streams; the headline "value" above is ONE trial per GPU' % tp['trials_per_gpu']}
    emit(line)
    comm.close()


_JSON_OUT = [None]


def emit(line):
    out = _JSON_OUT[0] or sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    args = parse_args()
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints
    # "NCCL version ..." there under NCCL_DEBUG=VERSION/WARN) are sent to stderr
    sys.stdout.flush()
    _JSON_OUT[0] = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
