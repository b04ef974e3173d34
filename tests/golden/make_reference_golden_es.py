"""Mints tests/golden/reference_v8_es.npz: the reference's ensemble-selection facade (facade/obtl_es.py ``ES``,
``--methods es``) run by the reference itself through SMBO_OFFLINE -- rows, perfs, every fitted theta, the weight
history and the global np.random state after every iteration.  Same method as make_reference_golden.py.
Build container only:   python tests/golden/make_reference_golden_es.py"""
import collections
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M               # noqa: E402
import numpy as np                              # noqa: E402
from tlbo.facade.obtl_es import ES              # noqa: E402


def main():
    types = M.RF_TYPES
    N, K, n_src, seed, iters = 300, 3, 50, 4465, 14
    rng = np.random.RandomState(77)
    tables = [M.data(rng, N, types, task=t) for t in range(K + 1)]
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
        fac = M.quiet(ES, cs, sources, list(target.keys()), seed, surrogate_type='gp', num_src_hpo_trial=n_src)
        n_source_fits = len(thetas)
        smbo = M.quiet(M.SMBO_OFFLINE, target, cs, fac, random_seed=seed, max_runs=iters, source_hpo_data=sources,
                       num_src_hpo_trial=n_src, surrogate_type='gp', initial_runs=3, logging_dir='/tmp/tlbo_ref_logs')
        keys = list(target.keys())
        rows, fits_per_iter, w_hist, states = [], [], [], []
        for _ in range(iters):
            before = len(thetas)
            cfg, _, _, _ = M.quiet(smbo.iterate)
            rows.append(keys.index(cfg))
            fits_per_iter.append(len(thetas) - before)
            w_hist.append(np.array(fac.w, dtype=np.float64).ravel())
            st = np.random.get_state()
            states.append(np.concatenate([st[1].astype(np.int64), [st[2]]]))
    finally:
        M.ref_gp.GaussianProcess._train = orig_train
    out.update(rows=np.array(rows), perfs=np.array(smbo.perfs, dtype=np.float64), thetas=np.array(thetas),
               n_source_fits=n_source_fits, fits_per_iter=np.array(fits_per_iter), w=np.array(w_hist),
               np_random_states=np.array(states))
    mu, var = fac.predict(tables[0][0][:64])
    out['final_mu'], out['final_var'] = mu, var
    print('rows', rows, 'fits', n_source_fits, fits_per_iter)
    path = os.path.join(HERE, 'reference_v8_es.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
