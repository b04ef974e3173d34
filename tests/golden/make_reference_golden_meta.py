"""Mints tests/golden/reference_v4_meta.npz: the reference's meta-feature baselines run by the reference itself --
facade/scot.py ``SCoT`` (RankSVM over all within-task pairs, then ONE GP on the K x 50 + n stacked rows with the
scaled data-set meta-features as extra columns) and facade/tstm.py ``TSTM`` (TST with meta-feature distance
weights) -- through SMBO_OFFLINE.  Same method as make_reference_golden.py.  Build container only:
    python tests/golden/make_reference_golden_meta.py"""
import collections
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M      # noqa: E402
import numpy as np                     # noqa: E402
from tlbo.facade.scot import SCoT      # noqa: E402
from tlbo.facade.tstm import TSTM      # noqa: E402
import tlbo.utils.rank_svm as ref_svm  # noqa: E402


def main():
    types = M.RF_TYPES
    N, K, n_src, seed = 300, 3, 50, 4465
    rng = np.random.RandomState(123)
    tables = [M.data(rng, N, types, task=t) for t in range(K + 1)]
    meta = rng.rand(K + 1, 4) * np.array([1.0, 10.0, 100.0, 0.1])
    meta[1, 2] = np.nan                       # exercises the mean imputer of base_facade.py:193-208
    meta[K, 0] = 1.7                          # target outside the sources' range: clipped after scaling
    out = dict(types=types, seed=seed, K=K, meta=meta)
    for t, (X, y) in enumerate(tables):
        out['task%d_X' % t], out['task%d_y' % t] = X, y
    for method, cls, iters in (('scot', SCoT, 8), ('tstm', TSTM, 8)):
        cs = M.rf_space()
        as_dict = lambda X, y: collections.OrderedDict((M.Configuration(cs, vector=x), float(v)) for x, v in zip(X, y))
        sources = [as_dict(*tables[t]) for t in range(1, K + 1)]
        target = as_dict(*tables[0])
        thetas, coefs = [], []
        orig_train = M.ref_gp.GaussianProcess._train
        orig_fit = ref_svm.RankSVM.fit

        def spy_train(self, X, y, do_optimize=True):
            r = orig_train(self, X, y, do_optimize)
            thetas.append(np.array(self.hypers, dtype=np.float64))
            return r

        def spy_fit(self, X, y):
            r = orig_fit(self, X, y)
            coefs.append(np.array(self.coef, dtype=np.float64))
            return r
        M.ref_gp.GaussianProcess._train = spy_train
        ref_svm.RankSVM.fit = spy_fit
        try:
            fac = M.quiet(cls, cs, sources, list(target.keys()), seed, surrogate_type='gp', num_src_hpo_trial=n_src,
                          metafeatures=[m for m in meta])
            n_source_fits = len(thetas)
            smbo = M.quiet(M.SMBO_OFFLINE, target, cs, fac, random_seed=seed, max_runs=iters, source_hpo_data=sources,
                           num_src_hpo_trial=n_src, surrogate_type='gp', initial_runs=3,
                           logging_dir='/tmp/tlbo_ref_logs')
            keys = list(target.keys())
            rows, fits_per_iter, w_hist = [], [], []
            for _ in range(iters):
                before = len(thetas)
                cfg, _, _, _ = M.quiet(smbo.iterate)
                rows.append(keys.index(cfg))
                fits_per_iter.append(len(thetas) - before)
                w_hist.append(np.array(fac.w, dtype=np.float64).ravel() if fac.w is not None else np.zeros(K + 1))
        finally:
            M.ref_gp.GaussianProcess._train = orig_train
            ref_svm.RankSVM.fit = orig_fit
        p = method + '_'
        out[p + 'rows'] = np.array(rows)
        out[p + 'perfs'] = np.array(smbo.perfs, dtype=np.float64)
        out[p + 'thetas'] = np.array(thetas)
        out[p + 'n_source_fits'] = n_source_fits
        out[p + 'fits_per_iter'] = np.array(fits_per_iter)
        out[p + 'w'] = np.array(w_hist)
        out[p + 'metafeatures_scaled'] = np.array(fac.metafeatures)
        if coefs:
            out[p + 'svm_coef'] = np.array(coefs)
        # posterior of the final model over a probe set (un-standardised mean / variance)
        Xq = tables[0][0][:64]
        mu, var = fac.predict(Xq)
        out[p + 'final_mu'], out[p + 'final_var'] = mu, var
        print(method, 'rows', rows, 'fits', n_source_fits, fits_per_iter)
    path = os.path.join(HERE, 'reference_v4_meta.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
