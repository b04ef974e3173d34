"""Mints tests/golden/mcmc_v1.npz from the CPU oracle (oracle/gp_mcmc.py) with fixed seeds.

Neither the reference nor emcee ships golden vectors for the MCMC path and neither can run in this
image, so the fixture freezes the ORACLE's outputs: log-probabilities ``_ll`` at given walker
positions, the state of a 40-step stretch-move chain (positions, log-probabilities, acceptance
counts), the pre-drawn random tables of the product's host half, and the marginalised posterior of
the resulting walker GPs.  ``python tests/golden/make_golden_mcmc.py`` regenerates the file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)
from oracle import gp as ogp                       # noqa: E402
from oracle import gp_mcmc as omc                  # noqa: E402

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])
NSTEPS = 40


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
    from tlbo_b200.model.gp_mcmc import stretch_move_tables        # host-only code, no GPU needed
    out = {}
    rng = np.random.RandomState(20261020)
    for tag, types, n in (('rf', RF_TYPES, 28), ('c4', np.zeros(4, int), 36)):
        spec = ogp.KernelSpec(types)
        X, y = data(rng, n, types)
        m = omc.OracleGPMCMC(types, 4465, chain_length=NSTEPS, burnin_steps=0)
        W, H = m.n_mcmc_walkers, spec.H
        ymean, ystd = np.mean(y), np.std(y)
        yn = (y - ymean) / ystd
        p0 = m._sample_p0()
        lp0 = np.array([omc.ll(p0[i].copy(), spec, X, yn) for i in range(W)])
        pos, lp, nacc = omc.emcee_run(lambda th: omc.ll(th, spec, X, yn), p0, NSTEPS, np.random.RandomState(77))
        tabs = stretch_move_tables(np.random.RandomState(77), W, H, NSTEPS)
        # walker GPs at the final positions, marginalised posterior at 50 query rows
        Xq, _ = data(rng, 50, types)
        mus, vs = [], []
        for i in range(W):
            g = ogp.OracleGP(types, 0, normalize_y=False)
            g._train(X, yn, do_optimize=False, theta=np.clip(pos[i], -50, 50))
            g.normalize_y, g.mean_y, g.std_y = True, ymean, ystd
            mu_i, v_i = g.predict(Xq)
            mus.append(mu_i.ravel())
            vs.append(v_i.ravel())
        mus, vs = np.array(mus), np.array(vs)
        gm = mus.mean(axis=0)
        gv = np.clip(np.var(mus, axis=0) + np.mean(vs, axis=0), np.finfo(float).eps, np.inf)
        out.update({tag + '_types': types, tag + '_X': X, tag + '_y': y, tag + '_ynorm': yn, tag + '_ymean': ymean,
                    tag + '_ystd': ystd, tag + '_p0': p0, tag + '_lp0': lp0, tag + '_pos': pos, tag + '_lp': lp,
                    tag + '_nacc': nacc, tag + '_sidx': tabs[0], tag + '_cidx': tabs[1], tag + '_zz': tabs[2],
                    tag + '_fac': tabs[3], tag + '_logu': tabs[4], tag + '_Xq': Xq, tag + '_mu': gm, tag + '_var': gv})
    return out


if __name__ == '__main__':
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'mcmc_v1.npz')
    np.savez_compressed(path, **build())
    print(path, os.path.getsize(path))
