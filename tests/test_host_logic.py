"""Host-side logic of the drop-in, checked on the CPU (no compute calls)."""
import numpy as np
import pytest

from oracle import gp as ogp
from oracle.facade import kfold_indices


def test_start_points_equal_the_oracle_restatement():
    """GaussianProcess._start_points (product) consumes the shared / own RNG streams exactly
    like the reference's _optimize (gp.py:211-239), restated by oracle.gp.restart_points."""
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.model.model_builder import build_model
    for types, seed in (([2, 2, 0, 0, 0, 0, 0, 0, 0], 4465), ([0] * 5, 3822), ([3, 0, 2], 7)):
        m = build_model('gp', TableSpace(types), np.random.RandomState(seed))
        p0 = m._start_points()
        spec = ogp.KernelSpec(np.array(types))
        want = np.vstack([spec.default_theta()[None, :], ogp.restart_points(spec, seed)])
        assert p0.shape == want.shape == (11, len(types) + 2)
        assert (p0 == want).all()
        np.testing.assert_array_equal(m.kernel.bounds, spec.log_bounds())


def test_kfold_slices_match_sklearn():
    from sklearn.model_selection import KFold
    from tlbo_b200.facade.topo_variant1 import _kfold_val_slices
    for n in (5, 7, 10, 23, 75):
        sl = _kfold_val_slices(n, 5)
        for (lo, hi), (tr, va), (otr, ova) in zip(sl, KFold(n_splits=5).split(np.zeros((n, 1))), kfold_indices(n)):
            assert list(range(lo, hi)) == list(va) == list(ova)
            assert list(tr) == list(otr)


def test_merge_best_is_the_reference_ordering():
    from tlbo_b200.distributed import merge_best
    rng = np.random.RandomState(0)
    for trial in range(50):
        N = 40
        ei = np.round(rng.rand(N), 1)            # many ties
        key = np.round(rng.rand(N), 1)
        order = np.lexsort((key, ei))[::-1]      # ties on (ei, key) resolve to the larger row
        cuts = sorted(rng.choice(np.arange(1, N), size=3, replace=False))
        quads = []
        for lo, hi in zip([0] + cuts, cuts + [N]):
            best = max(range(lo, hi), key=lambda r: (ei[r], key[r], r))
            quads.append((ei[best], key[best], best, 0))
        quads.append((-np.inf, -np.inf, -1, 2))  # an empty shard
        got = merge_best(quads)
        assert got[2] == order[0] and got[3] == 2


def test_configuration_objects():
    from tlbo_b200.config_space import Configuration, TableSpace, convert_configurations_to_array, get_types, \
        table_to_configurations
    cs = TableSpace([2, 0, 0])
    types, bounds = get_types(cs)
    assert list(types) == [2, 0, 0] and bounds[0][0] == 2
    X = np.array([[0, .1, .2], [1, .3, .4], [0, .1, .2]])
    cfgs = table_to_configurations(cs, X)
    assert cfgs[0] == cfgs[2] and cfgs[0] != cfgs[1] and len({cfgs[0], cfgs[1], cfgs[2]}) == 2
    np.testing.assert_array_equal(convert_configurations_to_array(cfgs), X)
    assert isinstance(cfgs[1], Configuration) and cfgs[1].index == 1


def test_unsupported_surrogates_raise():
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.model.model_builder import build_model
    cs = TableSpace([0, 0])
    for kind in ('rf', 'gp_rbf'):
        with pytest.raises(ValueError):
            build_model(kind, cs, np.random.RandomState(0))
    m = build_model('gp_mcmc', cs, np.random.RandomState(0))       # model_builder.py:74-89
    assert (m.n_mcmc_walkers, m.burnin_steps, m.chain_length) == (12, 250, 250)
    with pytest.raises(ValueError):
        build_model('nonsense', cs, np.random.RandomState(0))
    with pytest.raises(ValueError):
        RGPE(cs, [], None, 1, surrogate_type='rf')


def test_synthetic_benchmark_shape():
    from tlbo_b200 import synthetic
    b = synthetic.make_benchmark(n_tasks=4, T=200, seed=1)
    assert len(b['tables']) == 4 and b['tables'][0][0].shape == (200, 9)
    X, y = b['tables'][2]
    assert set(np.unique(X[:, 0])) <= {0.0, 1.0} and (X[:, [2, 4, 5, 8]] == 0).all()
    assert (y > 0).all() and (y < 1).all()
    src, ids = synthetic.leave_one_out_sources(b['tables'], 1, 2, 5)
    assert len(src) == 2 and 1 not in ids
