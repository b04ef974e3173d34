import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'efficient-tlbo_b200'))
import numpy as np, torch
from scipy.stats import norm
import bench
from tlbo_b200 import ops
class A: pass
args = A(); args.method='rgpe'; args.pool=100000; args.sources=29; args.n0=40
wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
fac, smbo = bench.make_product(args, wl)
for it in range(30):
    X = smbo.X_table[smbo.configurations]; Y = np.array(smbo.perfs)
    fac.train(X, Y.copy())
    eta = np.min((Y - Y.mean())/Y.std())
    res, mu, var, ei = fac._fused(smbo.acq_optimizer.X_dev, eta=eta, var_floor=1e-10, want_rows=True)
    e = ei.cpu().numpy(); m = mu.cpu().numpy(); v = var.cpu().numpy()
    bad = np.nonzero(e < 0)[0]
    print(it, 'n', len(Y), 'neg', len(bad), 'nneg kernel', res[3], 'best', res[:3])
    if len(bad):
        for b in bad[:5]:
            s = np.sqrt(v[b]); z = (eta - m[b]) / s
            fo = (eta - m[b]) * norm.cdf(z) + s * norm.pdf(z)
            print('  row', b, 'mu', m[b], 'var', v[b], 'z', z, 'ei gpu', e[b], 'ei numpy', fo)
        break
    smbo.configurations.append(res[2]); smbo.perfs.append(float(wl['yt'][res[2]]))
