"""Mints tests/golden/reference_v3_default.npz: the reference's SMBO_OFFLINE initial design when the space's
default configuration IS a table row at a non-zero position (tlbo/framework/smbo_offline.py:210-219).
Same method as make_reference_golden.py.  Build container only:  python tests/golden/make_reference_golden_default.py"""
import collections
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M      # noqa: E402
import numpy as np                     # noqa: E402


def main():
    types = M.RF_TYPES
    N, seed, iters, drow = 300, 3822, 7, 17
    rng = np.random.RandomState(77)
    X, y = M.data(rng, N, types, task=0)
    cs = M.rf_space()
    cs.get_default_configuration = lambda: M.Configuration(cs, vector=X[drow].copy())
    target = collections.OrderedDict((M.Configuration(cs, vector=x), float(v)) for x, v in zip(X, y))
    fac = M.quiet(M.NoTL, cs, [], list(target.keys()), seed, surrogate_type='gp', num_src_hpo_trial=50)
    thetas = []
    orig_train = M.ref_gp.GaussianProcess._train

    def spy_train(self, X_, y_, do_optimize=True):
        r = orig_train(self, X_, y_, do_optimize)
        thetas.append(np.array(self.hypers, dtype=np.float64))
        return r
    M.ref_gp.GaussianProcess._train = spy_train
    try:
        smbo = M.quiet(M.SMBO_OFFLINE, target, cs, fac, random_seed=seed, max_runs=iters, source_hpo_data=[],
                       num_src_hpo_trial=50, surrogate_type='gp', initial_runs=3, logging_dir='/tmp/tlbo_ref_logs')
        keys = list(target.keys())
        rows = []
        for _ in range(iters):
            cfg, _, _, _ = M.quiet(smbo.iterate)
            rows.append(keys.index(cfg))
    finally:
        M.ref_gp.GaussianProcess._train = orig_train
    out = dict(X=X, y=y, types=types, seed=seed, default_row=drow, rows=np.array(rows), thetas=np.array(thetas),
               perfs=np.array(smbo.perfs, dtype=np.float64))
    path = os.path.join(HERE, 'reference_v3_default.npz')
    np.savez_compressed(path, **out)
    print('rows', rows, 'wrote', path)


if __name__ == '__main__':
    main()
