"""Mints tests/golden/hotpath_v1.npz from the CPU oracle with fixed seeds.

The reference ships no golden vectors for this path and cannot be executed
in this image, so the fixtures freeze the ORACLE's outputs (kernel matrix, NLL + gradient,
factorisation, posterior mean / variance, EI, bootstrap ranking losses, selection) at fixed
hyper-parameters.  `python tests/golden/make_golden.py` regenerates the file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gp as ogp                                          # noqa: E402
from oracle.acquisition import OracleEI, first_unevaluated, sort_by_acq_value   # noqa: E402
from oracle.facade import loss_der3, loss_func3, rank_loss_vectorised  # noqa: E402

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])


def data(rng, n, types):
    d = len(types)
    X = rng.rand(n, d)
    for j in range(d):
        if types[j] > 0:
            X[:, j] = rng.randint(0, types[j], size=n)
    if d == 9:
        X[:, [2, 4, 5, 8]] = 0.0
    y = np.sin(3 * X[:, 3 % d]) + 0.5 * X[:, 6 % d] ** 2 + 0.3 * X[:, 0] + 0.05 * rng.randn(n)
    return X, y


def build():
    out = {}
    rng = np.random.RandomState(20261019)
    for tag, types, n in (('rf', RF_TYPES, 24), ('c5', np.zeros(5, int), 40)):
        spec = ogp.KernelSpec(types)
        X, y = data(rng, n, types)
        theta = np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.5, 0.05, spec.H - 2), [rng.uniform(-9, -4)]])
        g = ogp.OracleGP(types, 1)
        g.train(X, y, do_optimize=False, theta=theta)
        f, grad = ogp.nll(theta, spec, X, g.y)
        Xq, _ = data(rng, 64, types)
        Xq[:5] = X[:5]
        mu, var = g.predict(Xq)
        eta = float(np.min(y))

        class One(object):
            def predict_marginalized_over_instances(self, Z):
                m, v = g.predict(Z)
                v[v < 1e-10] = 1e-10
                return m, v
        ei = OracleEI(One())
        ei.update(eta=eta)
        acq = ei(Xq)
        order, keys = sort_by_acq_value(acq, np.random.RandomState(5))
        out.update({tag + '_types': types, tag + '_X': X, tag + '_y': y, tag + '_theta': theta,
                    tag + '_K': ogp.kernel(theta, spec, X), tag + '_nll': f, tag + '_grad': grad,
                    tag + '_L': g.L, tag + '_alpha': g.alpha, tag + '_Kinv': g.K_inv, tag + '_ynorm': g.y,
                    tag + '_ymean': g.mean_y, tag + '_ystd': g.std_y, tag + '_Xq': Xq, tag + '_mu': mu, tag + '_var': var,
                    tag + '_eta': eta, tag + '_ei': acq, tag + '_keys': keys, tag + '_order': order,
                    tag + '_pick': first_unevaluated(order, list(order[:2]))})
    # integer ranking losses (with ties) and the pairwise logistic loss
    n, S, Mr = 31, 6, 5
    yv = np.round(rng.randn(n), 1)
    ys = np.round(rng.randn(S, Mr, n), 1)
    lo = np.array([0, 0, 6, 12, 25])
    hi = np.array([n, 6, 12, 25, n])
    counts = np.array([[rank_loss_vectorised(yv, ys[s, m], lo[m], hi[m]) for m in range(Mr)] for s in range(S)])
    A = rng.randn(n, 4)
    w = np.array([0.1, 0.2, 0.3, 0.4])
    out.update(dict(rl_y=yv, rl_ys=ys, rl_lo=lo, rl_hi=hi, rl_counts=counts, pl_A=A, pl_w=w,
                    pl_loss=loss_func3(yv, A.dot(w)), pl_grad=loss_der3(yv, A, w)))
    return out


if __name__ == '__main__':
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hotpath_v1.npz')
    np.savez_compressed(path, **build())
    print('wrote', path, os.path.getsize(path), 'bytes')
