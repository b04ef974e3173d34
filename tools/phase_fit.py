"""Per-phase cycle breakdown of the fit kernel (needs the -DTLBO_PHASE_TIMING build:
TLBO_B200_LIB=efficient-tlbo_b200/lib/libtlbo_b200_timing.so python tools/phase_fit.py 60)."""
import sys, os, ctypes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficient-tlbo_b200'))
import numpy as np, torch
import bench
from tlbo_b200 import ops, _cabi
from tlbo_b200.model.gp import KernelInfo, GaussianProcess
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
wl = bench.build_workload(0, 4465, 1000, 29, 64)
types = wl['types']; ki = KernelInfo(types)
rng = np.random.RandomState(4465)
gp = GaussianProcess(None, types, None, seed=rng.randint(0, 10000), prior_rng=rng)
p0 = gp._start_points()
rows = wl['rows0'][:n]
X = wl['Xt'][rows]; y = wl['yt'][rows]; y = (y - y.mean()) / y.std()
pack = ops.SurrogatePack(1, n, 9)
fn = _cabi.lib.tlbo_debug_phase_cycles
buf = (ctypes.c_longlong * 24)()
names = ['hyp+Xs scale', 'K build + tile load', 'symmetric sweep', 'log-det partials', 'nll_data reduce', '(unused)',
         'W publish', 'pair loop', 'grad reduce + finish', 'optimiser feed']
for rep in range(2):
    fn(buf, 1)
    st, rt, rf, ri = ops.gp_fit_batched([X], [y], types, p0, ki.bounds[:, 0], ki.bounds[:, 1], [0.], [1.], pack,
                                        return_restarts=True)
    torch.cuda.synchronize()
    fn(buf, 0)
nev = buf[15]
tot = sum(buf[i] for i in range(10))
print('n=%d  evals of CTA 0 (restart 0): %d  total %.1f cycles/eval = %.1f us @1.965GHz' % (n, nev, tot / nev, tot / nev / 1965.0))
pass
print('  feed: pre-rebuild %.0f  rebuild_B %.0f  cauchy %.0f  subsm %.0f  line-search entry %.0f (cycles/eval)' % tuple(buf[16 + i] / nev for i in (4, 0, 1, 2, 3)))
for i, nm in enumerate(names):
    print('  %-24s %9.0f cycles/eval  %5.1f%%' % (nm, buf[i] / nev, 100.0 * buf[i] / tot))
