"""Do two fit launches on two CUDA streams overlap?  One launch alone, then two / three on their own streams
(own packs), timed between device-wide synchronisations."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficient-tlbo_b200'))
import numpy as np, torch
import bench
from tlbo_b200 import ops
from tlbo_b200.model.gp import KernelInfo, GaussianProcess
wl = bench.build_workload(0, 4465, 1000, 29, 60)
types = wl['types']
ki = KernelInfo(types)
rng = np.random.RandomState(4465)
gp = GaussianProcess(None, types, None, seed=rng.randint(0, 10000), prior_rng=rng)
p0 = gp._start_points()
n = 60
rows = wl['rows0'][:n]
X = wl['Xt'][rows]; y = wl['yt'][rows]; y = (y - y.mean()) / y.std()
Xs = [X] + [np.delete(X, slice(i * (n // 5), (i + 1) * (n // 5)), axis=0) for i in range(5)]
ys = [y] + [np.delete(y, slice(i * (n // 5), (i + 1) * (n // 5))) for i in range(5)]
ys = [(v - v.mean()) / v.std() for v in ys]
ops.set_fit_variant(1)
for S in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(S)]
    packs = [ops.SurrogatePack(6, n, 9) for _ in range(S)]
    ts = []
    for rep in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s, pk in zip(streams, packs):
            with torch.cuda.stream(s):
                ops.gp_fit_batched(Xs, ys, types, p0, ki.bounds[:, 0], ki.bounds[:, 1], [0.] * 6, [1.] * 6, pk, sync=False)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print('streams=%d  wall ms for all launches: %s' % (S, ['%.2f' % t for t in ts]))
