"""Fit launch time (warp optimiser) against the target's row count n on the bench-shaped problem: shows the tile-size
steps of the register-panel path (n + 1 <= 16 T)."""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficient-tlbo_b200'))
import numpy as np, torch
import bench
from tlbo_b200 import ops
from tlbo_b200.model.gp import KernelInfo, GaussianProcess
wl = bench.build_workload(0, 4465, 1000, 29, 100)
types = wl['types']
ki = KernelInfo(types)
rng = np.random.RandomState(4465)
gp = GaussianProcess(None, types, None, seed=rng.randint(0, 10000), prior_rng=rng)
p0 = gp._start_points()
ns = [int(a) for a in sys.argv[1:]] or [56, 60, 62, 63, 64, 65, 70, 76, 79, 80, 90]
for n in ns:
    rows = wl['rows0'][:n]
    X = wl['Xt'][rows]; y = wl['yt'][rows]; y = (y - y.mean()) / y.std()
    Xs = [X] + [np.delete(X, slice(i * (n // 5), (i + 1) * (n // 5)), axis=0) for i in range(5)]
    ys = [y] + [np.delete(y, slice(i * (n // 5), (i + 1) * (n // 5))) for i in range(5)]
    ys = [(v - v.mean()) / v.std() for v in ys]
    pack = ops.SurrogatePack(6, ops.batch_cap(n), 9)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        st, rt, rf, ri = ops.gp_fit_batched(Xs, ys, types, p0, ki.bounds[:, 0], ki.bounds[:, 1], [0.] * 6, [1.] * 6,
                                            pack, return_restarts=True)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    nf = ri[:, :, 0]
    print('n=%3d cap=%3d fit ms %s  nfev max=%d (target max %d) mean=%.1f  us/eval of the slowest restart %.1f' % (
        n, ops.batch_cap(n), ['%.2f' % t for t in ts[1:]], nf.max(), nf[0].max(), nf.mean(), 1e3 * min(ts[1:]) / nf.max()))
