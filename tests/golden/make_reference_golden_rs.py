"""Mints tests/golden/reference_v7_rs.npz: the reference's random-search baseline (facade/random_surrogate.py,
``--methods rs``) run by the reference itself through SMBO_OFFLINE -- rows, perfs and the facade's generator state.
Same method as make_reference_golden.py.  Build container only:
    python tests/golden/make_reference_golden_rs.py"""
import collections
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M                        # noqa: E402
import numpy as np                                       # noqa: E402
from tlbo.facade.random_surrogate import RandomSearch    # noqa: E402


def main():
    types = M.RF_TYPES
    N, seed, iters = 700, 4465, 12
    rng = np.random.RandomState(99)
    X, y = M.data(rng, N, types, task=0)
    cs = M.rf_space()
    target = collections.OrderedDict((M.Configuration(cs, vector=x), float(v)) for x, v in zip(X, y))
    fac = M.quiet(RandomSearch, cs, [], list(target.keys()), seed, surrogate_type='gp')
    smbo = M.quiet(M.SMBO_OFFLINE, target, cs, fac, random_seed=seed, max_runs=iters, source_hpo_data=[],
                   num_src_hpo_trial=50, surrogate_type='gp', initial_runs=3, logging_dir='/tmp/tlbo_ref_logs')
    keys = list(target.keys())
    rows = []
    for _ in range(iters):
        cfg, _, _, _ = M.quiet(smbo.iterate)
        rows.append(keys.index(cfg))
    st = fac.rng.get_state()
    out = dict(types=types, seed=seed, X=X, y=y, rows=np.array(rows), perfs=np.array(smbo.perfs, dtype=np.float64),
               rng_key=st[1], rng_pos=st[2])
    print('rows', rows)
    path = os.path.join(HERE, 'reference_v7_rs.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
