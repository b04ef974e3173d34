"""End-to-end parity of the drop-in (facade + EI + maximiser + offline loop on the GPU)
against the oracle's restatement of the reference loop, iteration by iteration.

Hyper-parameter optimisation is checked separately (tests/test_gpu_fit.py); here the
oracle's GPs are teacher-forced with the hyper-parameters the device fit selected, so that
everything downstream -- source / CV predictions, the bootstrap draws taken from the global
np.random stream, the integer ranking losses, weights, dilution flags, EI, tie-broken
ordering and the selected row -- must agree: counts and indices bit-exact, floats to 1e-6."""
import numpy as np
import pytest

import oracle.facade as ofacade
from oracle.acquisition import OracleSMBOOffline
from oracle.gp import OracleGP

pytestmark = pytest.mark.gpu


class ForcedGP(OracleGP):
    queue = []

    def train(self, X, Y, do_optimize=True, theta=None):
        th = ForcedGP.queue.pop(0)
        return super().train(X, Y, do_optimize=False, theta=th)


@pytest.fixture
def forced(monkeypatch):
    monkeypatch.setattr(ofacade, 'OracleGP', ForcedGP)
    ForcedGP.queue = []
    yield ForcedGP
    ForcedGP.queue = []


def _bench(K, N, seed, kind='rf', d=None):
    from tlbo_b200 import synthetic
    b = synthetic.make_benchmark(n_tasks=K + 1, T=N, kind=kind, d=d, seed=seed)
    sources, _ = synthetic.leave_one_out_sources(b['tables'], 0, K, seed)
    return b['types'], sources, b['tables'][0]


def _make(method, types, sources, seed, **kw):
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.notl import NoTL
    from tlbo_b200.facade.pogpe import POGPE
    from tlbo_b200.facade.rgpe import RGPE, TransBO_RGPE
    from tlbo_b200.facade.topo_variant1 import OBTLV
    from tlbo_b200.facade.tst import TST
    cs = TableSpace(types)
    cls = dict(rgpe=RGPE, trans_rgpe=TransBO_RGPE, topo=OBTLV, tst=TST, notl=NoTL, pogpe=POGPE)[method]
    prod = cls(cs, sources, None, seed, surrogate_type='gp', num_src_hpo_trial=50, **kw)
    return cs, prod


def _make_oracle(method, types, sources, seed, prod, forced):
    K = len(sources)
    if method != 'notl':
        forced.queue = [prod._pack.theta[k].cpu().numpy() for k in range(K)]
    if method == 'rgpe':
        return ofacade.OracleRGPE(types, sources, seed)
    if method == 'trans_rgpe':
        return ofacade.OracleRGPE(types, sources, seed, smooth=0.7)
    if method == 'topo':
        return ofacade.OracleOBTLV(types, sources, seed)
    if method == 'tst':
        return ofacade.OracleTST(types, sources, seed)
    if method == 'pogpe':
        return ofacade.OraclePOGPE(types, sources, seed)
    return ofacade.OracleNoTL(types, sources, seed)


@pytest.mark.parametrize('method,K,N,iters,kind,d', [
    ('rgpe', 4, 3000, 16, 'rf', None),
    ('rgpe_regrow', 3, 1500, 70, 'rf', None),
    ('trans_rgpe', 3, 2000, 14, 'rf', None),
    ('topo', 4, 3000, 14, 'rf', None),
    ('tst', 3, 2000, 9, 'rf', None),
    ('pogpe', 3, 2000, 9, 'rf', None),
    ('notl', 2, 2000, 8, 'cont', 5),
])
def test_offline_loop_matches_oracle(method, K, N, iters, kind, d, forced):
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    seed = 4465
    types, sources, (Xt, yt) = _bench(K, N, seed, kind, d)
    kw = {}
    if method == 'rgpe_regrow':
        # long run from a deliberately small pack: exercises pack growth (sources copied to the new
        # pack while their pool posteriors stay resident) and the larger panel tiles of the fit
        method, kw = 'rgpe', dict(max_target_rows=8)
    cs, prod = _make(method, types, sources, seed, **kw)
    ora = _make_oracle(method, types, sources, seed, prod, forced)
    assert not forced.queue
    smbo = SMBO_OFFLINE((Xt, yt), cs, prod, random_seed=seed, max_runs=iters)      # seeds np.random
    state = np.random.get_state()
    osmbo = OracleSMBOOffline(Xt, yt, ora, seed, max_runs=iters)
    np.random.set_state(state)
    if method in ('rgpe', 'trans_rgpe', 'topo'):
        for k in range(K):
            assert abs(prod.eta_list[k] - ora.eta_list[k]) < 1e-12
    for it in range(iters):
        s0 = np.random.get_state()
        cfg, st, perf, _ = smbo.iterate()
        s1 = np.random.get_state()
        n_prev = len(osmbo.evaluated)
        if n_prev >= 3:
            n_fit = 1
            if method in ('rgpe', 'trans_rgpe', 'topo') and n_prev >= 5:
                n_fit = 6
            forced.queue = [prod._pack.theta[K + j].cpu().numpy() for j in range(n_fit)]
        np.random.set_state(s0)
        idx = osmbo.iterate()
        s2 = np.random.get_state()
        assert not forced.queue
        # identical consumption of the global stream
        assert s1[2] == s2[2] and (s1[1] == s2[1]).all() and s1[3] == s2[3] and s1[4] == s2[4]
        assert cfg.index == idx, (it, cfg.index, idx)
        assert smbo.configurations == osmbo.evaluated
        if n_prev >= 3 and method != 'notl':
            np.testing.assert_allclose(np.asarray(prod.w, float), np.asarray(ora.w, float), rtol=0, atol=2e-5)
            if method != 'pogpe':                      # pogpe.py never logs a target weight
                np.testing.assert_allclose(prod.target_weight[-1], ora.target_weight[-1], rtol=0, atol=2e-5)
        if n_prev >= 3 and method in ('rgpe', 'trans_rgpe'):
            assert (prod.last_losses == ora.last_losses).all()
            assert list(prod.ignored_flag) == list(ora.ignored_flag)
        if n_prev >= 3 and osmbo.last_acq is not None:
            ei_o = float(osmbo.last_acq[idx, 0])
            ei_p = smbo.acq_optimizer.last_best[0]
            assert abs(ei_p - ei_o) <= 1e-6 * max(abs(ei_o), 1e-12) + 1e-13
            # the tie-break key of the picked row: the device generator's rng.rand(N)[idx]
            assert smbo.acq_optimizer.last_best[1] == osmbo.last_keys[idx]
    assert abs(smbo.get_adtm() - osmbo.get_adtm()) < 1e-15
    # the maximiser's MT19937 state lives on the device (one look-ahead draw ahead): brought back to the host
    # it equals the reference maximiser's rng after the same number of rng.rand(N) draws
    a, b = smbo.acq_optimizer.sync_host_rng().get_state(), osmbo.acq_rng.get_state()
    assert (a[1] == b[1]).all() and a[2] == b[2]


def test_facade_predict_and_ei_call_interface(forced):
    """predict / predict_marginalized_over_instances / EI.__call__ shapes, floors, errors."""
    from tlbo_b200.acquisition_function.acquisition import EI
    from tlbo_b200.config_space import table_to_configurations
    seed = 3822
    types, sources, (Xt, yt) = _bench(3, 500, seed)
    cs, prod = _make('rgpe', types, sources, seed)
    ora = _make_oracle('rgpe', types, sources, seed, prod, forced)
    X, y = Xt[:12].copy(), yt[:12].copy()
    np.random.seed(1)
    prod.train(X, y.copy())
    forced.queue = [prod._pack.theta[3 + j].cpu().numpy() for j in range(6)]
    np.random.seed(1)
    ora.train(X, y.copy())
    Xq = Xt[100:400]
    mu, var = prod.predict(Xq)
    mo, vo = ora.predict(Xq)
    assert mu.shape == (300, 1) and var.shape == (300, 1)
    np.testing.assert_allclose(mu, mo, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(var, vo, rtol=1e-6, atol=1e-13)
    m2, v2 = prod.predict_marginalized_over_instances(Xq)
    assert (v2 >= 1e-10).all()
    with pytest.raises(ValueError):
        prod.predict_marginalized_over_instances(Xq[:, :3])
    with pytest.raises(ValueError):
        prod.predict_marginalized_over_instances(Xq[0])
    ei = EI(prod)
    with pytest.raises(ValueError):
        ei(table_to_configurations(cs, Xq[:4]))
    ei.update(model=prod, eta=float(np.min((y - y.mean()) / y.std())), num_data=12)
    acq = ei(table_to_configurations(cs, Xq))
    assert acq.shape == (300, 1) and (acq >= 0).all()
    from oracle.acquisition import OracleEI
    oe = OracleEI(ora)
    oe.update(eta=ei.eta)
    np.testing.assert_allclose(acq, oe(Xq), rtol=1e-6, atol=1e-12)
    # a single surrogate behind the same interface (plain SMBO: 1e-5 floor)
    gp = prod.source_surrogates[0]
    m, v = gp.predict_marginalized_over_instances(Xq)
    og = ora.source_surrogates[0]
    mo, vo = og.predict_marginalized_over_instances(Xq)
    np.testing.assert_allclose(m, mo, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(v, vo, rtol=1e-6, atol=1e-13)
    e1 = EI(gp)
    e1.update(eta=-0.5)
    o1 = OracleEI(og)
    o1.update(eta=-0.5)
    np.testing.assert_allclose(e1._compute(Xq), o1._compute(Xq), rtol=1e-6, atol=1e-12)


def test_gp_train_validation_errors():
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.model.model_builder import build_model
    cs = TableSpace([2, 0, 0])
    m = build_model('gp', cs, np.random.RandomState(1))
    X = np.random.RandomState(0).rand(10, 3)
    X[:, 0] = np.round(X[:, 0])
    y = X[:, 1] * 2
    with pytest.raises(ValueError):
        m.train(X[:, :2], y)
    with pytest.raises(ValueError):
        m.train(X, y[:5])
    with pytest.raises(ValueError):
        m.train(X[0], y)
    with pytest.raises(Exception):
        m.predict(X)
    with pytest.raises(ValueError):
        build_model('rf', cs, np.random.RandomState(1))
    m.train(X, y)
    assert m.is_trained and m.hypers.shape == (5,)
    mu, var = m.predict(X)
    assert mu.shape == (10, 1) and var.shape == (10, 1)
    assert np.abs(mu.ravel() - y).max() < 0.05


def test_resident_pool_posteriors_are_bit_identical(forced):
    """The fused call with the source surrogates' posteriors read from the HBM-resident cache
    returns exactly the bits of the uncached call (per-row mu / var / EI and the selected row)."""
    import torch
    seed = 2854
    types, sources, (Xt, yt) = _bench(5, 4000, seed)
    cs, prod = _make('topo', types, sources, seed)
    X, y = Xt[:15].copy(), yt[:15].copy()
    prod.train(X, y.copy())
    from tlbo_b200 import ops
    Xd = ops.h2d(Xt)
    keys = ops.h2d(np.random.RandomState(3).rand(Xt.shape[0]))
    ex = ops.h2d(np.arange(15, dtype=np.int64))
    eta = float(np.min((y - y.mean()) / y.std()))
    r0, m0, v0, e0 = prod._fused(Xd, eta=eta, var_floor=1e-10, keys=keys, excluded=ex, want_rows=True)
    prod.register_static_pool(Xd)
    assert prod._pool_cache
    r1, m1, v1, e1 = prod._fused(Xd, eta=eta, var_floor=1e-10, keys=keys, excluded=ex, want_rows=True)
    assert r0 == r1
    assert torch.equal(m0, m1) and torch.equal(v0, v1) and torch.equal(e0, e1)
    # cached values are the per-surrogate predictions
    cmu, cvar = list(prod._pool_cache.values())[0]
    mo, vo = prod.source_surrogates[2].predict(Xt[:300])
    np.testing.assert_allclose(cmu[2, :300].cpu().numpy(), mo.ravel(), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(cvar[2, :300].cpu().numpy(), vo.ravel(), rtol=1e-10, atol=1e-16)


def test_offline_search_full_order_and_head_share_one_key_stream(forced):
    """``OfflineSearch.maximize``: the default head-only call and the reference-shaped ``full_order=True`` call
    consume the SAME ``rng.rand(N)`` stream (device generator with a look-ahead buffer): a head call, a full
    ordering and another head call equal the oracle's three consecutive ``_sort_configs_by_acq_value`` draws."""
    from oracle.acquisition import OracleEI, sort_by_acq_value
    from tlbo_b200.acquisition_function.acquisition import EI
    from tlbo_b200.optimizer.ei_offline_optimizer import OfflineSearch
    seed = 4465
    types, sources, (Xt, yt) = _bench(2, 700, seed)
    cs, prod = _make('rgpe', types, sources, seed)
    ora = _make_oracle('rgpe', types, sources, seed, prod, forced)
    X, y = Xt[:9].copy(), yt[:9].copy()
    np.random.seed(2)
    prod.train(X, y.copy())
    forced.queue = [prod._pack.theta[2 + j].cpu().numpy() for j in range(6)]
    np.random.seed(2)
    ora.train(X, y.copy())
    eta = float(np.min((y - y.mean()) / y.std()))
    ei = EI(prod)
    ei.update(model=prod, eta=eta, num_data=9)
    oe = OracleEI(ora)
    oe.update(eta=eta)
    acq_o = oe(Xt)
    search = OfflineSearch(Xt, ei, cs, rng=np.random.RandomState(seed), prefetch_keys=True)
    orng = np.random.RandomState(seed)
    # 1) head only
    head = search.maximize(None, 5000)
    order, keys = sort_by_acq_value(acq_o, orng)
    assert head[0].index == int(order[0]) and search.last_best[1] == keys[order[0]]
    # 2) full ordering (reference semantics)
    full = search.maximize(None, 5000, full_order=True)
    order, keys = sort_by_acq_value(acq_o, orng)
    got = np.array([c.index for c in full])
    top = 50                                    # beyond the head, EI ties at the 1e-6 level may swap neighbours
    assert len(got) == len(order) and (got[:top] == order[:top]).all()
    # 3) head again, with two evaluated rows excluded
    ex = [int(order[0]), int(order[1])]
    idx = search.best_unevaluated(ex)
    order, keys = sort_by_acq_value(acq_o, orng)
    want = [int(i) for i in order if int(i) not in ex][0]
    assert idx == want and search.last_best[1] == keys[want]
    a, b = search.sync_host_rng().get_state(), orng.get_state()
    assert (a[1] == b[1]).all() and a[2] == b[2]


def test_two_trials_in_one_process_select_the_rows_of_their_solo_runs():
    """Two RGPE trials (different tasks / seeds) run CONCURRENTLY in one process -- a thread, a CUDA stream and a
    private copy of the "process-global" np.random stream each (utils/global_stream.py) -- pick exactly the rows each
    of them picks when it runs alone in the process."""
    import threading
    import torch
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    from tlbo_b200.utils.global_stream import thread_stream
    K, N, iters = 4, 1500, 14

    def make(trial):
        types, sources, (Xt, yt) = _bench(K, N, 100 + trial)
        cs = TableSpace(types)
        fac = RGPE(cs, sources, None, 7 + trial, surrogate_type='gp', num_src_hpo_trial=50)
        return SMBO_OFFLINE((Xt, yt), cs, fac, random_seed=7 + trial, max_runs=10 ** 6)

    solo = []
    for trial in (0, 1):
        smbo = make(trial)
        for _ in range(iters):
            smbo.iterate()
        solo.append(list(smbo.configurations))
    got, errs = {}, []
    bar = threading.Barrier(2)

    def run(trial):
        try:
            with thread_stream(), torch.cuda.stream(torch.cuda.Stream()):
                smbo = make(trial)
                bar.wait()
                for _ in range(iters):
                    smbo.iterate()
                torch.cuda.current_stream().synchronize()
                got[trial] = list(smbo.configurations)
        except Exception as ex:                      # pragma: no cover
            errs.append(repr(ex))
            bar.abort()
    ts = [threading.Thread(target=run, args=(t,)) for t in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert got[0] == solo[0] and got[1] == solo[1]
    assert solo[0] != solo[1]
