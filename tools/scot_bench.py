"""SCoT at BASELINE scale: K = 29 sources x 50 observations + n target rows stacked into ONE GP (n_tot ~ 1.5 k rows,
d + m columns).  Times the parts of one BO iteration.  usage: scot_bench.py [K] [pool] [iters]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'efficient-tlbo_b200'))
import numpy as np, torch
import bench
from tlbo_b200.config_space import TableSpace
from tlbo_b200.facade.scot import SCoT
from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
import tlbo_b200.utils.rank_svm as rs

K = int(sys.argv[1]) if len(sys.argv) > 1 else 29
pool = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wl = bench.build_workload(0, 4465, pool, K, 40)
meta = np.random.RandomState(1).rand(K + 1, 4)
cs = TableSpace(wl['types'])
t0 = time.perf_counter()
fac = SCoT(cs, wl['sources'], None, wl['seed'], surrogate_type='gp', num_src_hpo_trial=50, metafeatures=[m for m in meta])
smbo = SMBO_OFFLINE((wl['Xt'], wl['yt']), cs, fac, random_seed=wl['seed'], max_runs=10 ** 9)
smbo.configurations = list(wl['rows0']); smbo.perfs = [float(wl['yt'][r]) for r in wl['rows0']]
print('construction %.2f s; stacked source rows %d' % (time.perf_counter() - t0, fac.X.shape[0]), flush=True)
orig_fit = rs.RankSVM.fit
svm_s = []
def timed_fit(self, X, y):
    a = time.perf_counter(); r = orig_fit(self, X, y); svm_s.append(time.perf_counter() - a); return r
rs.RankSVM.fit = timed_fit
for it in range(iters):
    torch.cuda.synchronize(); a = time.perf_counter()
    X = wl['Xt'][smbo.configurations]; Y = wl['yt'][smbo.configurations].copy()
    fac.train(X, Y)
    torch.cuda.synchronize(); b = time.perf_counter()
    cfg, _, _, _ = smbo.iterate()
    torch.cuda.synchronize(); c = time.perf_counter()
    info = fac.target_surrogate.fit_info_['restart_info'] if getattr(fac.target_surrogate, 'fit_info_', None) else None
    print('iter %d: n_tot=%d  train %.2f s (RankSVM %.2f s, GP fit %.2f s, nfev max %s)  full iterate %.2f s  row %d' % (
        it, fac.X.shape[0] + len(Y), b - a, svm_s[-2] if len(svm_s) > 1 else svm_s[-1], b - a - svm_s[-2 if len(svm_s) > 1 else -1],
        None if info is None else int(info[:, 0].max()), c - b, cfg.index), flush=True)
