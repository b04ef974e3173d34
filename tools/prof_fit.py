import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficient-tlbo_b200'))
import numpy as np, torch
import bench
from tlbo_b200 import ops
from tlbo_b200.model.gp import KernelInfo, GaussianProcess
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
wl = bench.build_workload(0, 4465, 1000, 29, 64)
types = wl['types']
ki = KernelInfo(types)
rng = np.random.RandomState(4465)
gp = GaussianProcess(None, types, None, seed=rng.randint(0, 10000), prior_rng=rng)
p0 = gp._start_points()
rows = wl['rows0'][:n]
X = wl['Xt'][rows]; y = wl['yt'][rows]; y = (y - y.mean()) / y.std()
Xs = [X] + [np.delete(X, slice(i * (n // 5), (i + 1) * (n // 5)), axis=0) for i in range(5)]
ys = [y] + [np.delete(y, slice(i * (n // 5), (i + 1) * (n // 5))) for i in range(5)]
ys = [(v - v.mean()) / v.std() for v in ys]
pack = ops.SurrogatePack(6, n, 9)
for rep in range(3):
    st, rt, rf, ri = ops.gp_fit_batched(Xs, ys, types, p0, ki.bounds[:, 0], ki.bounds[:, 1], [0.] * 6, [1.] * 6, pack,
                                        return_restarts=True)
print('nfev', ri[:, :, 0].max(), ri[:, :, 0].sum())
