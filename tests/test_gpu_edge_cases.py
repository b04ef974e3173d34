"""Edge cases of the hot path through the drop-in interface: tiny and ragged inputs, inactive
(non-finite) values, constant targets, exhausted candidate tables."""
import numpy as np
import pytest

import oracle.facade as ofacade
from oracle import gp as ogp
from oracle.acquisition import OracleEI

pytestmark = pytest.mark.gpu

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])


def _rows(rng, n, types=RF_TYPES):
    X = rng.rand(n, len(types))
    for j, t in enumerate(types):
        if t > 0:
            X[:, j] = rng.randint(0, t, size=n)
    if len(types) == 9:
        X[:, [2, 4, 5, 8]] = 0.0
    return X


@pytest.mark.parametrize('n', [1, 2, 3])
def test_tiny_training_sets(n):
    """n = 1 (std_y = 0 -> 1), 2, 3 observations: fit at given theta and predict."""
    from tlbo_b200 import ops
    rng = np.random.RandomState(n)
    X = _rows(rng, n)
    y = rng.rand(n)
    g = ogp.OracleGP(RF_TYPES, 1)
    th = g.spec.default_theta()
    th[-1] = np.log(1e-4)
    g.train(X, y, do_optimize=False, theta=th)
    pack = ops.SurrogatePack(1, n, 9)
    st = ops.gp_factorize_batched([X], [g.y], RF_TYPES, [th], [g.mean_y], [g.std_y], pack)
    assert not st.any()
    Xq = _rows(rng, 33)
    Xq[0] = X[0]
    mu, var = ops.gp_predict(pack, RF_TYPES, Xq)
    mo, vo = g.predict(Xq)
    np.testing.assert_allclose(mu[0].cpu().numpy(), mo.ravel(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(var[0].cpu().numpy(), vo.ravel(), rtol=1e-6, atol=1e-13)


def test_nonfinite_inputs_are_imputed_like_the_reference():
    """Inactive (NaN) hyper-parameters become -1 on both the train and the predict side
    (reference tlbo/model/base_gp.py:148-151)."""
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.model.model_builder import build_model
    rng = np.random.RandomState(0)
    types = np.zeros(4, int)
    X = rng.rand(14, 4)
    X[::3, 2] = np.nan
    y = np.sin(3 * X[:, 0]) + np.nan_to_num(X[:, 2])
    m = build_model('gp', TableSpace(types), np.random.RandomState(5))
    m.train(X, y, do_optimize=False)
    g = ogp.OracleGP(types, 5)
    g.train(X, y, do_optimize=False)
    Xq = rng.rand(40, 4)
    Xq[::4, 2] = np.nan
    mu, var = m.predict(Xq)
    mo, vo = g.predict(Xq)
    np.testing.assert_allclose(mu, mo, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(var, vo, rtol=1e-6, atol=1e-13)


@pytest.mark.parametrize('N', [1, 5, 127, 129])
def test_small_and_ragged_candidate_pools(N):
    """Pools smaller than / not a multiple of the kernel's tile; odd row counts take the
    non-TMA staging path."""
    import torch
    from tlbo_b200 import ops
    rng = np.random.RandomState(N)
    X = _rows(rng, 20)
    y = rng.rand(20)
    g = ogp.OracleGP(RF_TYPES, 1)
    th = g.spec.default_theta()
    th[-1] = np.log(1e-3)
    g.train(X, y, do_optimize=False, theta=th)
    pack = ops.SurrogatePack(1, 20, 9)
    ops.gp_factorize_batched([X], [g.y], RF_TYPES, [th], [g.mean_y], [g.std_y], pack)
    Xc = _rows(rng, N)
    dev = torch.device('cuda')
    res, mu, var, ei = ops.predict_ei_fused(pack, RF_TYPES, torch.ones(1, dtype=torch.float64, device=dev), None,
                                            torch.from_numpy(Xc).to(dev), float(y.min()), want_rows=True, nmax=20)

    class One(object):
        def predict_marginalized_over_instances(self, Z):
            m, v = g.predict(Z)
            v[v < 1e-10] = 1e-10
            return m, v
    oe = OracleEI(One())
    oe.update(eta=float(y.min()))
    acq = oe(Xc)
    np.testing.assert_allclose(ei.cpu().numpy(), acq.ravel(), rtol=1e-6, atol=1e-12)
    assert res[2] == int(np.lexsort((np.arange(N), acq.ravel()))[-1]) or acq.ravel()[res[2]] == acq.max()
    # every row excluded -> no candidate left
    ex = torch.arange(N, dtype=torch.int64, device=dev)
    none = ops.predict_ei_fused(pack, RF_TYPES, torch.ones(1, dtype=torch.float64, device=dev), None,
                                torch.from_numpy(Xc).to(dev), float(y.min()), excluded=ex, nmax=20)
    assert none[2] == -1


def test_exhausted_table_raises_like_the_reference():
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.notl import NoTL
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    rng = np.random.RandomState(1)
    Xt = _rows(rng, 6)
    yt = rng.rand(6)
    cs = TableSpace(RF_TYPES)
    smbo = SMBO_OFFLINE((Xt, yt), cs, NoTL(cs, [], None, 3), random_seed=3, max_runs=10)
    for _ in range(6):
        smbo.iterate()
    assert sorted(smbo.configurations) == list(range(6))
    with pytest.raises(ValueError):
        smbo.iterate()


def test_constant_targets_take_the_reference_guard():
    """All-equal y: the facade bumps y[0] by 1e-4 in place before standardising
    (reference base_facade.py:96-99) -- same mutation, same model."""
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.notl import NoTL
    rng = np.random.RandomState(2)
    X = _rows(rng, 8)
    y = np.full(8, 0.25)
    yo = y.copy()
    cs = TableSpace(RF_TYPES)
    prod = NoTL(cs, [], None, 11)
    prod.train(X, y)
    assert y[0] == 0.25 + 1e-4 and (y[1:] == 0.25).all()
    keep = ofacade.OracleGP

    class Forced(ogp.OracleGP):
        def train(self, X_, Y_, do_optimize=True, theta=None):
            return super().train(X_, Y_, do_optimize=False, theta=prod._pack.theta[0].cpu().numpy())
    ofacade.OracleGP = Forced
    try:
        ora = ofacade.OracleNoTL(RF_TYPES, [], 11)
        ora.train(X, yo)
    finally:
        ofacade.OracleGP = keep
    assert (y == yo).all()
    Xq = _rows(rng, 50)
    mu, var = prod.predict(Xq)
    mo, vo = ora.predict(Xq)
    np.testing.assert_allclose(mu, mo, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(var, vo, rtol=1e-6, atol=1e-13)


@pytest.mark.parametrize('method', ['rgpe', 'topo'])
def test_native_host_prep_equals_numpy_host_prep_on_device(method, monkeypatch):
    """The same facade trained through ``tlbo_host_prepare_jobs`` and through the numpy statement of the host
    half: identical fitted hyper-parameters, loss tables, weights and (mutated) targets, bit for bit --
    including a target vector whose CV folds trigger the all-equal guards."""
    import tlbo_b200.facade.base_facade as BF
    from tlbo_b200 import synthetic
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.facade.topo_variant1 import OBTLV
    b = synthetic.make_benchmark(n_tasks=3, T=400, kind='rf', seed=5)
    sources, _ = synthetic.leave_one_out_sources(b['tables'], 0, 2, 5)
    Xt, yt = b['tables'][0]
    cls = dict(rgpe=RGPE, topo=OBTLV)[method]
    out = []
    for native in (True, False):
        if not native:
            monkeypatch.setattr(BF.BaseFacade, '_native_prep_ok', lambda self, X, y: False)
        fac = cls(TableSpace(b['types']), sources, None, 11, surrogate_type='gp', num_src_hpo_trial=50)
        res = []
        for n, flat in ((23, False), (12, True)):
            X = np.ascontiguousarray(Xt[:n]); y = np.array(yt[:n], dtype=np.float64)
            if flat:
                y[2:] = 0.25              # every training set that drops rows 0-1 is all-equal
            np.random.seed(3)
            fac.train(X, y)
            res.append((fac._pack.theta[fac.K:fac.K + 6].cpu().numpy().copy(), np.array(fac.w, dtype=np.float64).copy(),
                        y.copy(), None if method != 'rgpe' else fac.last_losses.copy()))
        out.append(res)
    for (t1, w1, y1, l1), (t2, w2, y2, l2) in zip(*out):
        assert np.array_equal(t1, t2) and np.array_equal(w1, w2) and np.array_equal(y1, y2)
        if l1 is not None:
            assert np.array_equal(l1, l2)
