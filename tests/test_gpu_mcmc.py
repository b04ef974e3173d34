"""GPU parity tests of the MCMC-marginalised GP path (BASELINE config 5):
``tlbo_gp_mcmc_chain`` / ``tlbo_mcmc_marginalize`` and the ``GaussianProcessMCMC`` mirror against
the oracle's restatement of ``tlbo/model/gp_mcmc.py`` + emcee's stretch move.

Bars: log-probabilities 1e-9 relative (``-inf`` exactly); given the same random tables the
device chain makes the SAME accept / reject decisions as the oracle (positions 1e-9, acceptance
counts equal); marginalised mean / variance 1e-6 relative (the kernel itself is bit-exact
against numpy's reductions)."""
import numpy as np
import pytest

from oracle import gp as ogp
from oracle import gp_mcmc as omc

pytestmark = pytest.mark.gpu

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])


def _data(rng, n, types, const_cols=(2, 4, 5, 8)):
    d = len(types)
    X = rng.rand(n, d)
    for j in range(d):
        if types[j] > 0:
            X[:, j] = rng.randint(0, types[j], size=n)
    if len(types) == 9:
        X[:, list(const_cols)] = 0.0
    y = np.sin(3 * X[:, 3 % d]) + 0.5 * X[:, 6 % d] ** 2 + 0.3 * X[:, 0] + 0.05 * rng.randn(n)
    return X, (y - y.mean()) / y.std()


def _bounds_raw(spec):
    return omc.prior_bounds_raw(spec)


@pytest.mark.parametrize('n,types', [(6, RF_TYPES), (50, RF_TYPES), (75, RF_TYPES), (40, np.zeros(5, int)),
                                     (150, np.zeros(5, int)), (200, np.zeros(5, int)), (100, np.zeros(20, int))])
def test_log_probability_matches_oracle(n, types):
    from tlbo_b200 import ops
    rng = np.random.RandomState(n)
    spec = ogp.KernelSpec(types)
    H = spec.H
    B, W = 3, 8
    lb = spec.log_bounds()
    Xs, ys, pos = [], [], np.empty((B, W, H))
    for b in range(B):
        X, y = _data(rng, n - b, types)
        Xs.append(X)
        ys.append(y)
        for w in range(W):
            pos[b, w] = rng.uniform(lb[:, 0], lb[:, 1])
            pos[b, w, -1] = rng.uniform(-14, -2)
    pos[0, 1, 1] = 0.2            # above the length-scale bound -> Tophat -inf
    pos[1, 2, -1] = -60.0         # clamped to -50, outside the noise bound
    pos[2, 3, 0] = 3.0            # amplitude above exp(2)
    batch = ops.upload_batch(Xs, ys)
    h = ops.gp_mcmc_chain(batch, types, W, 0, pos, None, None, _bounds_raw(spec), spec.default_theta())
    st, pos_out, lp, nacc = h.finish()
    assert np.all(st == 0) and np.all(nacc == 0)
    assert np.array_equal(pos_out, pos)
    for b in range(B):
        for w in range(W):
            want = omc.ll(pos[b, w].copy(), spec, Xs[b], ys[b])
            if np.isfinite(want):
                assert lp[b, w] == pytest.approx(want, rel=1e-9, abs=1e-9), (b, w)
            else:
                assert lp[b, w] == -np.inf, (b, w)
    assert lp[0, 1] == -np.inf and lp[1, 2] == -np.inf and lp[2, 3] == -np.inf


def _run_both(types, n, W, nsteps, seed, B=2):
    from tlbo_b200 import ops
    from tlbo_b200.model.gp_mcmc import stretch_move_tables
    rng = np.random.RandomState(seed)
    spec = ogp.KernelSpec(types)
    H = spec.H
    lb = spec.log_bounds()
    Xs, ys = [], []
    for b in range(B):
        X, y = _data(rng, n - 2 * b, types)
        Xs.append(X)
        ys.append(y)
    p0 = np.empty((B, W, H))
    for b in range(B):
        for w in range(W):
            p0[b, w] = rng.uniform(lb[:, 0], lb[:, 1])
            p0[b, w, -1] = rng.uniform(-12, -1)
    p0[0, 0, 1] = 0.3             # one walker starts outside the box (log-probability -inf)
    tabs = [stretch_move_tables(np.random.RandomState(100 + b), W, H, nsteps) for b in range(B)]
    batch = ops.upload_batch(Xs, ys)
    h = ops.gp_mcmc_chain(batch, types, W, nsteps, p0, ops.McmcTables(tabs), list(range(B)), _bounds_raw(spec),
                          spec.default_theta())
    st, pos, lp, nacc = h.finish()
    assert np.all(st == 0)
    for b in range(B):
        want = omc.emcee_run(lambda th: omc.ll(th, spec, Xs[b], ys[b]), p0[b], nsteps,
                             np.random.RandomState(100 + b), check_independent=False)
        assert np.array_equal(nacc[b], want[2]), b
        assert np.allclose(pos[b], want[0], rtol=1e-9, atol=1e-12), b
        fin = np.isfinite(want[1])
        assert np.array_equal(np.isfinite(lp[b]), fin)
        assert np.allclose(lp[b][fin], want[1][fin], rtol=1e-9, atol=1e-9), b
    return nacc


@pytest.mark.parametrize('types,n,W,nsteps', [(RF_TYPES, 30, 34, 25), (np.zeros(5, int), 50, 22, 25),
                                              (np.zeros(3, int), 140, 16, 6), (np.zeros(3, int), 200, 16, 4),
                                              (np.zeros(20, int), 60, 66, 5)])
def test_chain_makes_the_oracles_decisions(types, n, W, nsteps):
    nacc = _run_both(types, n, W, nsteps, seed=n + W)
    assert nacc.sum() > 0


def test_full_length_chain_reference_shape():
    """250 burn-in + 250 chain steps, 34 walkers, n = 50, d = 9 (model_builder.py:74-89)."""
    nacc = _run_both(RF_TYPES, 50, 34, 500, seed=5, B=1)
    frac = nacc.sum() / (34 * 500.0)
    assert 0.02 < frac < 0.9


def test_many_surrogates_span_several_cooperative_launches():
    """29 source surrogates x 17 active walkers exceed one co-resident wave (148 SMs)."""
    from tlbo_b200 import ops
    from tlbo_b200.model.gp_mcmc import stretch_move_tables
    types = RF_TYPES
    spec = ogp.KernelSpec(types)
    H, W, B, nsteps = spec.H, 34, 29, 3
    rng = np.random.RandomState(1)
    lb = spec.log_bounds()
    Xs, ys = [], []
    for b in range(B):
        X, y = _data(rng, 20, types)
        Xs.append(X)
        ys.append(y)
    p0 = rng.uniform(lb[:, 0], lb[:, 1], size=(B, W, H))
    p0[:, :, -1] = rng.uniform(-12, -1, size=(B, W))
    tab = stretch_move_tables(np.random.RandomState(9), W, H, nsteps)
    h = ops.gp_mcmc_chain(ops.upload_batch(Xs, ys), types, W, nsteps, p0, ops.McmcTables([tab]), None,
                          _bounds_raw(spec), spec.default_theta())
    st, pos, lp, nacc = h.finish()
    for b in (0, 8, 9, 28):
        want = omc.emcee_run(lambda th: omc.ll(th, spec, Xs[b], ys[b]), p0[b], nsteps, np.random.RandomState(9),
                             check_independent=False)
        assert np.array_equal(nacc[b], want[2]) and np.allclose(pos[b], want[0], rtol=1e-9, atol=1e-12)


def test_marginalize_is_numpy_bit_exact():
    import torch
    from tlbo_b200 import ops
    rng = np.random.RandomState(0)
    G, W, N = 3, 34, 1000
    mu = rng.randn(G * W, N)
    var = rng.rand(G * W, N) * 1e-3
    mu[W:2 * W] = mu[W]                  # identical walkers: var(mu) == 0
    var[W:2 * W, :10] = 0.0              # -> clipped at eps
    gmu, gvar = ops.mcmc_marginalize(ops.h2d(mu), ops.h2d(var), G, W)
    gmu, gvar = gmu.cpu().numpy(), gvar.cpu().numpy()
    for g in range(G):
        m = mu[g * W:(g + 1) * W]
        v = var[g * W:(g + 1) * W]
        want_v = np.clip(np.var(m, axis=0) + np.mean(v, axis=0), np.finfo(float).eps, np.inf)
        assert np.array_equal(gmu[g], m.mean(axis=0))
        assert np.array_equal(gvar[g], want_v)
    assert np.all(gvar[1, :10] == np.finfo(float).eps)


@pytest.mark.parametrize('types,n', [(RF_TYPES, 25), (np.array([0, 0, 0, 2, 2]), 18)])
def test_model_train_predict_matches_oracle(types, n):
    from tlbo_b200.model.gp_mcmc import GaussianProcessMCMC
    rng = np.random.RandomState(n)
    X, y = _data(rng, n, types)
    y = 0.3 + 2.0 * y                                  # exercise the y normalisation
    Xq, _ = _data(rng, 300, types)
    seed = 4465
    want = omc.OracleGPMCMC(types, seed, chain_length=12, burnin_steps=12).train(X, y)
    prior_rng = np.random.RandomState(seed)
    got = GaussianProcessMCMC(None, types, None, n_mcmc_walkers=want.n_mcmc_walkers, chain_length=12,
                              burnin_steps=12, seed=prior_rng.randint(low=0, high=10000), prior_rng=prior_rng)
    got.train(X, y)
    assert got.is_trained and len(got.models) == want.n_mcmc_walkers
    assert np.array_equal(got.n_accepted_, want.n_accepted)
    assert np.allclose(got.hypers, want.hypers, rtol=1e-9, atol=1e-12)
    assert got._n_ll_evals == want.n_ll_evals
    mu, var = got.predict(Xq)
    wmu, wvar = want.predict(Xq)
    assert mu.shape == (300, 1) and var.shape == (300, 1)
    assert np.allclose(mu, wmu, rtol=1e-6, atol=1e-9) and np.allclose(var, wvar, rtol=1e-6, atol=1e-12)
    # one walker GP through the per-walker view
    m3, v3 = got.models[3].predict(Xq[:7])
    w3, wv3 = want.models[3].predict(Xq[:7])
    assert np.allclose(m3, w3, rtol=1e-6) and np.allclose(v3, wv3, rtol=1e-6, atol=1e-12)
    # second train on the same object: no burn-in, continues from p0 with the advanced model rng
    want.train(X, y)
    got.train(X, y)
    assert np.array_equal(got.n_accepted_, want.n_accepted)
    assert np.allclose(got.hypers, want.hypers, rtol=1e-9, atol=1e-12)
    mu, var = got.predict_marginalized_over_instances(Xq)
    wmu, wvar = want.predict_marginalized_over_instances(Xq)
    assert np.allclose(mu, wmu, rtol=1e-6, atol=1e-9) and np.allclose(var, wvar, rtol=1e-6, atol=1e-12)


def test_build_model_gp_mcmc_shape():
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.model.model_builder import build_model
    cs = TableSpace(RF_TYPES)
    m = build_model('gp_mcmc', cs, np.random.RandomState(1))
    assert m.n_mcmc_walkers == 34 and m.chain_length == 250 and m.burnin_steps == 250
    with pytest.raises(ValueError):
        m.train(np.zeros((4, 3)), np.zeros(4))
    with pytest.raises(Exception):
        m.predict(np.zeros((4, 9)))


# ---- whole BO loops with MCMC-marginalised surrogates (config 5: TOPO / RGPE / TST over gp_mcmc) ----

def _loop_case(method, K, N, seed, steps):
    import oracle.facade as ofacade
    from oracle.acquisition import OracleSMBOOffline
    from tlbo_b200 import synthetic
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.notl import NoTL
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.facade.topo_variant1 import OBTLV
    from tlbo_b200.facade.tst import TST
    b = synthetic.make_benchmark(n_tasks=K + 1, T=N, seed=seed)
    sources, _ = synthetic.leave_one_out_sources(b['tables'], 0, K, seed)
    types, (Xt, yt) = b['types'], b['tables'][0]
    cs = TableSpace(types)
    cls = dict(rgpe=RGPE, topo=OBTLV, tst=TST, notl=NoTL)[method]
    prod = cls(cs, sources, None, seed, surrogate_type='gp_mcmc', num_src_hpo_trial=20, mcmc_steps=steps)
    obase = dict(rgpe=ofacade.OracleRGPE, topo=ofacade.OracleOBTLV, tst=ofacade.OracleTST,
                 notl=ofacade.OracleNoTL)[method]
    ocls = type('Mcmc' + obase.__name__, (obase,), dict(surrogate_type='gp_mcmc', mcmc_steps=steps))
    ora = ocls(types, sources, seed, num_src_hpo_trial=20)
    return cs, prod, ora, Xt, yt, OracleSMBOOffline


@pytest.mark.parametrize('method,K,iters', [('topo', 3, 9), ('rgpe', 2, 8), ('tst', 2, 6), ('notl', 1, 6)])
def test_offline_loop_with_mcmc_surrogates_matches_oracle(method, K, iters):
    """No teacher forcing: given the same seeds the device sampler makes the oracle's accept /
    reject decisions, so hyper-parameter samples, weights, EI and the selected rows all agree."""
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    seed = 4465
    cs, prod, ora, Xt, yt, OSMBO = _loop_case(method, K, 1200, seed, (5, 5))
    if method != 'notl':
        for k in range(K):
            assert np.allclose(prod.source_surrogates[k].hypers, ora.source_surrogates[k].hypers, rtol=1e-9,
                               atol=1e-12)
            assert np.array_equal(prod.source_surrogates[k].n_accepted_, ora.source_surrogates[k].n_accepted)
    smbo = SMBO_OFFLINE((Xt, yt), cs, prod, random_seed=seed, max_runs=iters)
    state = np.random.get_state()
    osmbo = OSMBO(Xt, yt, ora, seed, max_runs=iters)
    np.random.set_state(state)
    for it in range(iters):
        s0 = np.random.get_state()
        cfg, _, _, _ = smbo.iterate()
        s1 = np.random.get_state()
        n_prev = len(osmbo.evaluated)
        np.random.set_state(s0)
        idx = osmbo.iterate()
        s2 = np.random.get_state()
        assert s1[2] == s2[2] and (s1[1] == s2[1]).all()
        assert cfg.index == idx, (it, cfg.index, idx)
        if n_prev >= 3:
            assert np.allclose(prod.target_surrogate.hypers, ora.target_surrogate.hypers, rtol=1e-8, atol=1e-11)
            if method != 'notl':
                np.testing.assert_allclose(np.asarray(prod.w, float), np.asarray(ora.w, float), rtol=0, atol=2e-5)
            if method == 'rgpe':
                assert (prod.last_losses == ora.last_losses).all()
            if osmbo.last_acq is not None:
                ei_o = float(osmbo.last_acq[idx, 0])
                ei_p = smbo.acq_optimizer.last_best[0]
                assert abs(ei_p - ei_o) <= 1e-6 * max(abs(ei_o), 1e-12) + 1e-13
    assert abs(smbo.get_adtm() - osmbo.get_adtm()) < 1e-15


def test_mcmc_facade_predict_rows_match_oracle():
    seed = 3822
    cs, prod, ora, Xt, yt, _ = _loop_case('topo', 2, 400, seed, (4, 4))
    X, y = Xt[:11].copy(), yt[:11].copy()
    prod.train(X, y.copy())
    ora.train(X, y.copy())
    Xq = Xt[50:250]
    mu, var = prod.predict(Xq)
    wmu, wvar = ora.predict(Xq)
    assert np.allclose(mu, wmu, rtol=1e-6, atol=1e-9) and np.allclose(var, wvar, rtol=1e-6, atol=1e-13)
    mu, var = prod.predict_marginalized_over_instances(Xq)
    wmu, wvar = ora.predict_marginalized_over_instances(Xq)
    assert np.allclose(mu, wmu, rtol=1e-6, atol=1e-9) and np.allclose(var, wvar, rtol=1e-6, atol=1e-13)


def test_chain_samples_the_hyperparameter_posterior():
    """Statistical contract of the sampler, independent of any stream order: the stationary distribution of the
    stretch-move chain is the hyper-parameter posterior exp(_ll).  A one-column problem (H = 3: amplitude, length
    scale, noise) is small enough for quadrature: posterior mean and spread from a two-level grid (coarse over the
    prior box, fine over the region holding the mass; grid values from the device's own batched ``_ll``, which
    ``test_log_probability_matches_oracle`` pins to the oracle at 1e-9) against the final ensembles of 64
    independent 16-walker chains after 400 steps (1 024 draws; a wrong proposal density, e.g. a missing
    ``(ndim - 1) log z`` factor, or a broken accept rule shifts these moments by far more than the tolerances)."""
    from tlbo_b200 import ops
    from tlbo_b200.model.gp_mcmc import stretch_move_tables
    types = np.zeros(1, int)
    spec = ogp.KernelSpec(types)
    H = spec.H
    assert H == 3
    rng = np.random.RandomState(77)
    X = rng.rand(14, 1)
    y = np.sin(4 * X[:, 0]) + 0.15 * rng.randn(14)
    y = (y - y.mean()) / y.std()
    braw = _bounds_raw(spec)
    lo, hi = np.log(braw[:, 0]), np.log(braw[:, 1])
    GB, GW = 8, 128                                   # 1 024 grid points per evaluation call
    rep = ops.upload_batch([X] * GB, [y] * GB)

    def grid_lp(axes):
        G = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1).reshape(-1, H)
        out = np.empty(len(G))
        for s in range(0, len(G), GB * GW):
            blk = G[s:s + GB * GW]
            pad = np.vstack([blk, np.repeat(blk[-1:], GB * GW - len(blk), axis=0)])
            h = ops.gp_mcmc_chain(rep, types, GW, 0, pad.reshape(GB, GW, H), None, None, braw, spec.default_theta())
            st, _, lp, _ = h.finish()
            assert np.all(st == 0)
            out[s:s + len(blk)] = lp.reshape(-1)[:len(blk)]
        return G, out

    eps = 1e-9
    coarse = [np.linspace(lo[k] + eps, hi[k] - eps, 28) for k in range(H)]
    G, lp = grid_lp(coarse)
    keep = lp > lp.max() - 30.0
    step = np.array([c[1] - c[0] for c in coarse])
    blo = np.maximum(G[keep].min(axis=0) - step, lo + eps)
    bhi = np.minimum(G[keep].max(axis=0) + step, hi - eps)
    fine = [np.linspace(blo[k], bhi[k], 44) for k in range(H)]
    G, lp = grid_lp(fine)
    w = np.exp(lp - lp.max())
    # trapezoid weights on the uniform fine grid
    tw = np.ones((44,) * H)
    for k in range(H):
        sl = [slice(None)] * H
        for edge in (0, -1):
            sl[k] = edge
            tw[tuple(sl)] *= 0.5
    w = w * tw.reshape(-1)
    w /= w.sum()
    mean_q = (w[:, None] * G).sum(axis=0)
    std_q = np.sqrt((w[:, None] * (G - mean_q) ** 2).sum(axis=0))

    B, W, nsteps = 64, 16, 400
    p0 = np.empty((B, W, H))
    for k in range(H):
        p0[:, :, k] = rng.uniform(blo[k], bhi[k], size=(B, W))
    tabs = [stretch_move_tables(np.random.RandomState(500 + b), W, H, nsteps) for b in range(B)]
    batch = ops.upload_batch([X] * B, [y] * B)
    h = ops.gp_mcmc_chain(batch, types, W, nsteps, p0, ops.McmcTables(tabs), list(range(B)), braw, spec.default_theta())
    st, pos, lpc, nacc = h.finish()
    assert np.all(st == 0) and np.all(np.isfinite(lpc))
    acc = nacc.sum() / float(B * W * nsteps)
    S = pos.reshape(-1, H)
    mean_c, std_c = S.mean(axis=0), S.std(axis=0)
    print('posterior mean quadrature %s chain %s; std quadrature %s chain %s; acceptance %.2f' % (
        np.round(mean_q, 3), np.round(mean_c, 3), np.round(std_q, 3), np.round(std_c, 3), acc))
    assert 0.2 < acc < 0.8
    assert np.all(np.abs(mean_c - mean_q) <= 0.25 * std_q), (mean_c, mean_q, std_q)
    assert np.all(std_c / std_q > 0.75) and np.all(std_c / std_q < 1.3), (std_c, std_q)
