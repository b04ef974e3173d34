"""Mints tests/golden/reference_v6_sgpr.npz: the reference's stacking-GP baseline run by the reference itself --
facade/stacking_gpr.py ``SGPR`` (one residual GP per source history in turn, then one on the target observations;
prior mean / "sigma" cached over the union of all configurations) with surrogate_type='gp' -- through SMBO_OFFLINE.
Same method as make_reference_golden.py.  Build container only:
    python tests/golden/make_reference_golden_sgpr.py"""
import collections
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M               # noqa: E402
import numpy as np                              # noqa: E402
from tlbo.facade.stacking_gpr import SGPR       # noqa: E402


def main():
    types = M.RF_TYPES
    N, K, n_src, seed, iters = 300, 3, 50, 4465, 8
    rng = np.random.RandomState(321)
    tables = [M.data(rng, N, types, task=t) for t in range(K + 1)]
    # the source tasks of the reference's benchmark are evaluated on the SAME configuration table as the target:
    # make source 2 share its first 20 configurations with source 1 and the target share 30 rows with source 1, so
    # that the de-duplicated configuration set (stacking_gpr.py:27-36) is exercised
    tables[2][0][:20] = tables[1][0][:20]
    tables[0][0][:30] = tables[1][0][10:40]
    out = dict(types=types, seed=seed, K=K)
    for t, (X, y) in enumerate(tables):
        out['task%d_X' % t], out['task%d_y' % t] = X, y
    cs = M.rf_space()
    as_dict = lambda X, y: collections.OrderedDict((M.Configuration(cs, vector=x), float(v)) for x, v in zip(X, y))
    sources = [as_dict(*tables[t]) for t in range(1, K + 1)]
    target = as_dict(*tables[0])
    thetas = []
    orig_train = M.ref_gp.GaussianProcess._train

    def spy_train(self, X, y, do_optimize=True):
        r = orig_train(self, X, y, do_optimize)
        thetas.append(np.array(self.hypers, dtype=np.float64))
        return r
    M.ref_gp.GaussianProcess._train = spy_train
    try:
        fac = M.quiet(SGPR, cs, sources, list(target.keys()), seed, surrogate_type='gp', num_src_hpo_trial=n_src)
        n_source_fits = len(thetas)
        out['n_configs'] = len(fac.configs_set)
        out['configs_X'] = np.array(fac.configs_X, dtype=np.float64)
        out['prior_mu'] = np.array(fac.cached_prior_mu, dtype=np.float64)
        out['prior_sigma'] = np.array(fac.cached_prior_sigma, dtype=np.float64)
        smbo = M.quiet(M.SMBO_OFFLINE, target, cs, fac, random_seed=seed, max_runs=iters, source_hpo_data=sources,
                       num_src_hpo_trial=n_src, surrogate_type='gp', initial_runs=3, logging_dir='/tmp/tlbo_ref_logs')
        keys = list(target.keys())
        rows, stack_mu, stack_sigma = [], [], []
        for _ in range(iters):
            cfg, _, _, _ = M.quiet(smbo.iterate)
            rows.append(keys.index(cfg))
            stack_mu.append(np.array(fac.cached_stacking_mu, dtype=np.float64))
            stack_sigma.append(np.array(fac.cached_stacking_sigma, dtype=np.float64))
    finally:
        M.ref_gp.GaussianProcess._train = orig_train
    out['rows'] = np.array(rows)
    out['perfs'] = np.array(smbo.perfs, dtype=np.float64)
    out['thetas'] = np.array(thetas)
    out['n_source_fits'] = n_source_fits
    out['stack_mu'] = np.array(stack_mu)
    out['stack_sigma'] = np.array(stack_sigma)
    mu, var = fac.predict(tables[0][0][:64])
    out['final_mu'], out['final_var'] = mu, var
    print('rows', rows, 'source fits', n_source_fits, 'configs', len(fac.configs_set), 'fits', len(thetas))
    path = os.path.join(HERE, 'reference_v6_sgpr.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
