"""Mints tests/golden/reference_v5_cond.npz: the reference's kernels / GaussianProcess on a space WITH conditions
(``has_conditions``: the activity mask of tlbo/model/gp_kernels.py:21-29 multiplies every kernel), fed arrays
whose inactive entries are NaN (imputed to -1 by GaussianProcess._impute_inactive, base_gp.py:148-151).
Same method as make_reference_golden.py (the stand-in space reports one condition).  Build container only:
    python tests/golden/make_reference_golden_cond.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M      # noqa: E402
import numpy as np                     # noqa: E402


def main():
    rng = np.random.RandomState(515)
    types = M.RF_TYPES
    cs = M.rf_space()
    cs.get_conditions = lambda: ['child | parent == x']           # only the count is read (base_gp.py:135)
    n = 40
    X, y = M.data(rng, n, types)
    Xq, _ = M.data(rng, 96, types)
    # conditional columns: numeric column 6 is inactive when categorical 0 == 1; column 7 inactive for a third
    for A in (X, Xq):
        A[A[:, 0] == 1, 6] = np.nan
        A[rng.rand(A.shape[0]) < 0.33, 7] = np.nan
    model = M.build_model('gp', cs, np.random.RandomState(7))
    assert model.kernel.has_conditions
    H = len(model.kernel.theta)
    thetas = np.array([np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.5, 0.05, H - 2), [rng.uniform(-9, -4)]])
                       for _ in range(3)])
    Xi = model._impute_inactive(X)
    Xqi = model._impute_inactive(Xq)
    k = model.kernel.clone_with_theta(thetas[0])
    K, dK = k(Xi, eval_gradient=True)
    out = dict(types=types, X=X, y=y, Xq=Xq, thetas=thetas, K=K, dK=dK, Ks=k(Xqi, Xi))
    M.quiet(model._train, X, y.copy(), do_optimize=False)
    model._all_priors = model._get_all_priors(add_bound_priors=False)
    vals = [model._nll(th.copy()) for th in thetas]
    out['nll_f'] = np.array([v[0] for v in vals])
    out['nll_g'] = np.array([v[1] for v in vals])
    out['ynorm'] = model.gp.y_train_.copy()
    model = M.build_model('gp', cs, np.random.RandomState(7))
    M.quiet(model.train, X, y.copy())
    out['theta_star'] = np.array(model.hypers)
    mu, var = model.predict(Xq)
    out['mu'], out['var'] = mu, var
    ei = M.EI(model)
    ei.update(model=model, eta=float(np.min(y)), num_data=n)
    out['eta'] = float(np.min(y))
    out['acq'] = ei._compute(Xq)
    frac = 1.0 - float((K != 0).mean())
    print('masked-out fraction of K: %.2f; theta* %s' % (frac, np.round(out['theta_star'], 3)))
    path = os.path.join(HERE, 'reference_v5_cond.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path))


if __name__ == '__main__':
    main()
