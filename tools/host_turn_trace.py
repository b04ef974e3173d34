"""How long does one RGPE trial hold its host turn per BO step (what bounds several trials per GPU), and where?
Solo trial with thread_stream(turns=True): the segments between its waiting() blocks are the host segments."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from tlbo_b200.utils import global_stream as gs  # noqa: E402


class A:
    pass


args = A()
args.method = 'rgpe'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
with gs.thread_stream(turns=True):
    wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
    fac, smbo = bench.make_product(args, wl)
    for _ in range(5):
        smbo.iterate()
    torch.cuda.synchronize()
    N = 40
    gs.trace = []
    marks = []
    t0 = time.perf_counter()
    for _ in range(N):
        marks.append(time.perf_counter())
        smbo.iterate()
    t1 = time.perf_counter()
    tr = gs.trace
    gs.trace = None
print('ms/step %.3f, waiting() blocks per step %.2f' % ((t1 - t0) / N * 1e3, len(tr) / N))
wait = sum(b - a for a, b in tr)
print('per step: waiting %.3f ms, turn held %.3f ms' % (wait / N * 1e3, (t1 - t0 - wait) / N * 1e3))
# per step: the segments in order
per = {}
for i in range(N):
    lo, hi = marks[i], marks[i + 1] if i + 1 < N else t1
    w = [(a, b) for a, b in tr if lo <= a < hi]
    cur = lo
    for j, (a, b) in enumerate(w):
        per.setdefault(('held', j), []).append(a - cur)
        per.setdefault(('wait', j), []).append(b - a)
        cur = b
    per.setdefault(('held', len(w)), []).append(hi - cur)
for k in sorted(per, key=lambda k: (k[1], k[0])):
    print('  %-5s %d  mean %7.1f us  (n=%d)' % (k[0], k[1], 1e6 * np.mean(per[k]), len(per[k])))
