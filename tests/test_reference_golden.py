"""The oracle pinned against the REFERENCE'S OWN outputs (tests/golden/reference_v1.npz).

The fixture was produced by tests/golden/make_reference_golden.py, which imports the unmodified
reference from /root/reference in the build container and runs its kernels, ``GaussianProcess``,
``EI``, ``scipy_solve``, facades and ``SMBO_OFFLINE`` loop (absent third-party imports stood in for by
tests/golden/refshims/).  These tests run everywhere (no /root/reference needed): the numpy/scipy
oracle must reproduce the reference's numbers -- kernels / likelihood / posterior to rounding,
integer ranking losses, weights, selected rows and the global np.random stream state exactly."""
import os

import numpy as np
import pytest

import oracle.facade as ofacade
from oracle import gp as ogp
from oracle.acquisition import OracleEI, OracleSMBOOffline

HERE = os.path.dirname(os.path.abspath(__file__))
R = np.load(os.path.join(HERE, 'golden', 'reference_v1.npz'))
TAGS = ['rf', 'c5']


@pytest.mark.parametrize('tag', TAGS)
def test_kernel_matches_reference(tag):
    """gp_kernels.py Matern / Hamming / Constant / White / Product / Sum ``_call`` bodies."""
    types, X, Xq, theta = R['kern_%s_types' % tag], R['kern_%s_X' % tag], R['kern_%s_Xq' % tag], R['kern_%s_thetas' % tag][0]
    spec = ogp.KernelSpec(types)
    K, dK = ogp.kernel(theta, spec, X, eval_gradient=True)
    np.testing.assert_allclose(K, R['kern_%s_K' % tag], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(dK, R['kern_%s_dK' % tag], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(ogp.kernel(theta, spec, Xq, X), R['kern_%s_Ks' % tag], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(ogp.kernel_diag(theta, spec, Xq), R['kern_%s_diag' % tag], rtol=1e-14)


@pytest.mark.parametrize('tag', TAGS)
def test_nll_matches_reference(tag):
    """GaussianProcess._nll: sklearn LML + lognormal / horseshoe priors, value and gradient."""
    types, X = R['kern_%s_types' % tag], R['kern_%s_X' % tag]
    spec = ogp.KernelSpec(types)
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy(), do_optimize=False, theta=spec.default_theta())
    np.testing.assert_allclose(g.y, R['nll_%s_ynorm' % tag], rtol=1e-14, atol=1e-15)
    for i, theta in enumerate(R['kern_%s_thetas' % tag]):
        f, grad = ogp.nll(theta, spec, X, g.y)
        assert abs(f - R['nll_%s_f' % tag][i]) <= 1e-10 * abs(R['nll_%s_f' % tag][i])
        np.testing.assert_allclose(grad, R['nll_%s_g' % tag][i], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize('tag', TAGS)
def test_fit_matches_reference(tag):
    """GaussianProcess._optimize: identical start points (prior / uniform draws of the shared and the
    model's own RandomState), the same optimum restart by restart, the same theta*."""
    types, X = R['kern_%s_types' % tag], R['kern_%s_X' % tag]
    spec = ogp.KernelSpec(types)
    p0 = R['fit_%s_p0' % tag]
    np.testing.assert_array_equal(p0[0], spec.default_theta())
    np.testing.assert_allclose(ogp.restart_points(spec, 7), p0[1:], rtol=1e-15, atol=0)
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy())
    f_ref = R['fit_%s_f' % tag]
    f_got = np.array([r[1] for r in g.restart_results])
    np.testing.assert_allclose(f_got, f_ref, rtol=1e-6, atol=1e-6)
    assert int(np.argmin(f_got)) == int(np.argmin(f_ref))
    np.testing.assert_allclose(g.hypers, R['fit_%s_theta_star' % tag], rtol=1e-4, atol=1e-4)
    # posterior of the fully trained model (each side with its own optimiser run)
    mu, var = g.predict(R['kern_%s_Xq' % tag])
    np.testing.assert_allclose(mu, R['pred_%s_mu' % tag], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(var, R['pred_%s_var' % tag], rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize('tag', TAGS)
def test_predict_and_ei_match_reference(tag):
    """GaussianProcess.predict at a fixed theta (the regressor's K_inv_ variance form), the 1e-5 floor of
    predict_marginalized_over_instances, EI._compute and EI.__call__."""
    types, X, Xq = R['kern_%s_types' % tag], R['kern_%s_X' % tag], R['kern_%s_Xq' % tag]
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy(), do_optimize=False, theta=R['kern_%s_thetas' % tag][1])
    np.testing.assert_allclose(g.L, R['pred_%s_fixed_L' % tag], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(g.alpha, R['pred_%s_fixed_alpha' % tag], rtol=1e-7, atol=1e-9)
    scale = np.abs(R['pred_%s_fixed_Kinv' % tag]).max()
    np.testing.assert_allclose(g.K_inv, R['pred_%s_fixed_Kinv' % tag], rtol=1e-7, atol=1e-10 * scale)
    mu, var = g.predict(Xq)
    np.testing.assert_allclose(mu, R['pred_%s_fixed_mu' % tag], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var, R['pred_%s_fixed_var' % tag], rtol=1e-6, atol=1e-13)
    # trained model: teacher-forced to the reference's theta*
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy(), do_optimize=False, theta=R['fit_%s_theta_star' % tag])
    mu, var = g.predict(Xq)
    np.testing.assert_allclose(mu, R['pred_%s_mu' % tag], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, R['pred_%s_var' % tag], rtol=1e-6, atol=1e-13)
    mu2, var2 = g.predict_marginalized_over_instances(Xq)
    np.testing.assert_allclose(var2, R['pred_%s_marg_var' % tag], rtol=1e-6, atol=1e-13)
    ei = OracleEI(g)
    ei.update(model=g, eta=float(R['ei_%s_eta' % tag]), num_data=X.shape[0])
    np.testing.assert_allclose(ei._compute(Xq), R['ei_%s_acq' % tag], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(ei(Xq), R['ei_%s_acq_call' % tag], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_slsqp_weights_match_reference(tag):
    """utils/scipy_solver.py scipy_solve with loss_type=3 (pairwise logistic)."""
    w, status = ofacade.scipy_solve(R['slsqp_%s_A' % tag], R['slsqp_%s_y' % tag])
    assert bool(status) == bool(R['slsqp_%s_status' % tag])
    np.testing.assert_allclose(w, R['slsqp_%s_w' % tag], rtol=1e-7, atol=1e-9)


class _ForcedGP(ogp.OracleGP):
    """Teacher-forced to the reference's fitted hyper-parameters, in the reference's fit order."""
    queue = []

    def train(self, X, Y, do_optimize=True, theta=None):
        return super().train(X, Y, do_optimize=False, theta=_ForcedGP.queue.pop(0))


def _state_vec():
    st = np.random.get_state()
    return np.concatenate([st[1].astype(np.int64), [st[2], st[3]], [np.float64(st[4]).view(np.int64)]])


LOOPS = {'rgpe': lambda t, s, seed: ofacade.OracleRGPE(t, s, seed, literal_loops=True),
         'trans_rgpe': lambda t, s, seed: ofacade.OracleRGPE(t, s, seed, smooth=0.7),
         'topo': lambda t, s, seed: ofacade.OracleOBTLV(t, s, seed),
         'topo_v3': lambda t, s, seed: ofacade.OracleTOPOV3(t, s, seed),
         'obtl': lambda t, s, seed: ofacade.OracleOBTL(t, s, seed, literal_loops=True),
         'tst': lambda t, s, seed: ofacade.OracleTST(t, s, seed),
         'pogpe': lambda t, s, seed: ofacade.OraclePOGPE(t, s, seed),
         'notl': lambda t, s, seed: ofacade.OracleNoTL(t, s, seed)}


@pytest.mark.parametrize('method', sorted(LOOPS))
def test_offline_loop_matches_reference(method, monkeypatch):
    """SMBO_OFFLINE + facade + EI + OfflineSearch, iteration by iteration, against the reference's own
    run: selected rows, weights, dilution flags, integer ranking-loss tables and the global np.random
    stream state are EXACT; the oracle's GPs take the reference's fitted theta (the optimiser itself is
    covered by test_fit_matches_reference)."""
    monkeypatch.setattr(ofacade, 'OracleGP', _ForcedGP)
    K, seed, types = int(R['loop_K']), int(R['loop_seed']), R['loop_types']
    p = 'loop_%s_' % method
    tables = [(R['loop_task%d_X' % t], R['loop_task%d_y' % t]) for t in range(K + 1)]
    thetas = [th for th in R[p + 'thetas']]
    nsrc = int(R[p + 'n_source_fits'])
    _ForcedGP.queue = thetas[:nsrc]
    fac = LOOPS[method](types, [(X.copy(), y.copy()) for X, y in tables[1:]], seed)
    assert not _ForcedGP.queue
    np.testing.assert_allclose(fac.eta_list, R[p + 'eta_list'], rtol=1e-13)
    smbo = OracleSMBOOffline(tables[0][0], tables[0][1], fac, seed, initial_runs=3)
    used, loss_rows = nsrc, 0
    ref_losses = R[p + 'losses']
    for it, row in enumerate(R[p + 'rows']):
        nfit = int(R[p + 'fits_per_iter'][it])
        _ForcedGP.queue = thetas[used:used + nfit]
        used += nfit
        got = smbo.iterate()
        assert not _ForcedGP.queue, 'iteration %d: the oracle fitted fewer GPs than the reference' % it
        assert got == int(row), 'iteration %d: oracle picked row %d, reference %d' % (it, got, row)
        np.testing.assert_array_equal(_state_vec(), R[p + 'np_random_state'][it], err_msg='np.random stream, it %d' % it)
        if fac.w is not None:
            np.testing.assert_allclose(np.asarray(fac.w, dtype=np.float64).ravel(), R[p + 'w'][it], rtol=1e-6,
                                       atol=1e-8, err_msg='weights, iteration %d' % it)
        if method in ('rgpe', 'trans_rgpe') and nfit:
            np.testing.assert_array_equal(np.asarray(fac.ignored_flag, dtype=bool), R[p + 'ignored'][it])
            if nfit == 6:
                np.testing.assert_array_equal(fac.last_losses, ref_losses[loss_rows:loss_rows + 30])
            loss_rows += 30
        if method == 'obtl' and nfit == 6:
            # obtl.py:111-117: the 50 x (K + 1) sampled soft ranking losses (captured from the reference's np.argmin)
            np.testing.assert_allclose(fac.last_sample_losses, ref_losses[loss_rows:loss_rows + 50], rtol=1e-9)
            loss_rows += 50
    np.testing.assert_allclose(smbo.perfs, R[p + 'perfs'], rtol=0, atol=0)
    np.testing.assert_allclose(fac.target_weight, R[p + 'target_weight'], rtol=1e-6, atol=1e-8)
    if p + 'hist_ws' in R.files and len(R[p + 'hist_ws']):
        np.testing.assert_allclose(np.array([np.asarray(w, dtype=np.float64) for w in fac.hist_ws]), R[p + 'hist_ws'],
                                   rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize('tag', ['rf', 'c4'])
def test_gp_mcmc_matches_reference(tag):
    """model/gp_mcmc.py GaussianProcessMCMC as run by the reference itself (short chains; the stretch-move
    sampler on both sides is the restated one -- see tests/golden/refshims/emcee.py -- so this pins _ll with
    its Tophat bound priors, the prior-sampled initial ensemble, the burn-in / no-burn-in flow, the walker
    models and the marginalised prediction, not emcee)."""
    from oracle import gp_mcmc as omc
    types, X, y, Xq = (R['mcmc_%s_%s' % (tag, k)] for k in ('types', 'X', 'y', 'Xq'))
    g = omc.OracleGPMCMC(types, 11, chain_length=12, burnin_steps=12)
    assert g.n_mcmc_walkers == int(R['mcmc_%s_walkers' % tag])
    g.train(X, y.copy())
    np.testing.assert_allclose(np.array(g.hypers), R['mcmc_%s_hypers1' % tag], rtol=1e-9, atol=1e-11)
    mu, var = g.predict(Xq)
    np.testing.assert_allclose(mu, R['mcmc_%s_mu1' % tag], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, R['mcmc_%s_var1' % tag], rtol=1e-6, atol=1e-13)
    want = R['mcmc_%s_ll' % tag]
    got = np.array([omc.ll(th.copy(), g.spec, g.X, g.y) for th in R['mcmc_%s_ll_thetas' % tag]], dtype=np.float64)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin) and fin.any() and (~fin).any()
    np.testing.assert_allclose(got[fin], want[fin], rtol=1e-10)
    g.train(R['mcmc_%s_X2' % tag], R['mcmc_%s_y2' % tag].copy())
    np.testing.assert_allclose(np.array(g.hypers), R['mcmc_%s_hypers2' % tag], rtol=1e-9, atol=1e-11)
    mu, var = g.predict_marginalized_over_instances(Xq)
    np.testing.assert_allclose(mu, R['mcmc_%s_mu2' % tag], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, R['mcmc_%s_var2' % tag], rtol=1e-6, atol=1e-13)


def _online_objective(x):
    return float(np.sin(3 * x[3]) + 0.5 * x[6] ** 2 + 0.3 * x[0] + 0.2 * np.cos(5 * x[7]) * x[1])


RF_CONST = np.array([0, 0, 1, 0, 1, 1, 0, 0, 1], dtype=bool)


@pytest.mark.parametrize('method', ['rgpe', 'topo'])
def test_online_loop_matches_reference(method, monkeypatch):
    """SMBO_ONLINE + InterleavedLocalAndRandomSearch (LocalSearch scoring ONE neighbour per acquisition
    call, sorted RandomSearch over ~4990 configurations, ChallengerList) as run by the reference itself:
    the oracle picks the same configurations with the same np.random stream consumption.  (The space's
    sampling / neighbourhood streams are the restated ones on both sides: tests/golden/refshims/ConfigSpace.)"""
    from oracle.ei_optimization import OracleSMBOOnline, OracleSpace
    monkeypatch.setattr(ofacade, 'OracleGP', _ForcedGP)
    K, types, seed = int(R['loop_K']), R['loop_types'], int(R['online_seed'])
    p = 'online_%s_' % method
    sources = [(R['loop_task%d_X' % t].copy(), R['loop_task%d_y' % t].copy()) for t in range(1, K + 1)]
    thetas = [th for th in R[p + 'thetas']]
    nsrc = int(R[p + 'n_source_fits'])
    _ForcedGP.queue = thetas[:nsrc]
    fac = LOOPS[method](types, sources, seed)
    smbo = OracleSMBOOnline(_online_objective, OracleSpace(types, RF_CONST), fac, seed, initial_runs=3,
                            first_config=R['online_first'])
    used = nsrc
    for it, pick in enumerate(R[p + 'picks']):
        nfit = int(R[p + 'fits_per_iter'][it])
        _ForcedGP.queue = thetas[used:used + nfit]
        used += nfit
        got = smbo.iterate()
        assert not _ForcedGP.queue
        np.testing.assert_array_equal(got, pick, err_msg='iteration %d' % it)
        np.testing.assert_array_equal(_state_vec(), R[p + 'np_random_state'][it], err_msg='np.random stream, it %d' % it)
        np.testing.assert_allclose(np.asarray(fac.w, dtype=np.float64).ravel(), R[p + 'w'][it], rtol=1e-6, atol=1e-8)
    np.testing.assert_array_equal(np.asarray(smbo.perfs), R[p + 'perfs'])


def test_default_configuration_row_matches_reference(monkeypatch):
    """smbo_offline.py:210-219: the first evaluation is the space's default configuration when it is a table
    row (here row 17 of 300), fixture reference_v3_default.npz from the reference's own NoTL run."""
    D = np.load(os.path.join(HERE, 'golden', 'reference_v3_default.npz'))
    monkeypatch.setattr(ofacade, 'OracleGP', _ForcedGP)
    _ForcedGP.queue = [th for th in D['thetas']]
    fac = ofacade.OracleNoTL(D['types'], [], int(D['seed']))
    smbo = OracleSMBOOffline(D['X'], D['y'], fac, int(D['seed']), initial_runs=3, default=D['X'][int(D['default_row'])])
    got = [smbo.iterate() for _ in D['rows']]
    assert got[0] == int(D['default_row']) == 17
    assert got == D['rows'].tolist()
    np.testing.assert_array_equal(np.asarray(smbo.perfs), D['perfs'])


META = {'scot': lambda t, s, seed, m: ofacade.OracleSCoT(t, s, seed, metafeatures=m),
        'tstm': lambda t, s, seed, m: ofacade.OracleTSTM(t, s, seed, metafeatures=m)}


@pytest.mark.parametrize('method', sorted(META))
def test_meta_feature_baselines_match_reference(method, monkeypatch):
    """facade/scot.py (RankSVM + one stacked GP with meta-feature columns) and facade/tstm.py as run by the
    reference itself (fixture reference_v4_meta.npz): scaled meta-features, RankSVM coefficients, selected rows,
    weights and the final posterior; the oracle's GPs take the reference's fitted theta."""
    D = np.load(os.path.join(HERE, 'golden', 'reference_v4_meta.npz'))
    monkeypatch.setattr(ofacade, 'OracleGP', _ForcedGP)
    K, seed, types = int(D['K']), int(D['seed']), D['types']
    p = method + '_'
    tables = [(D['task%d_X' % t], D['task%d_y' % t]) for t in range(K + 1)]
    thetas = [th for th in D[p + 'thetas']]
    nsrc = int(D[p + 'n_source_fits'])
    _ForcedGP.queue = thetas[:nsrc]
    fac = META[method](types, [(X.copy(), y.copy()) for X, y in tables[1:]], seed, [m for m in D['meta']])
    assert not _ForcedGP.queue
    np.testing.assert_allclose(fac.metafeatures, D[p + 'metafeatures_scaled'], rtol=1e-14, atol=1e-15)
    smbo = OracleSMBOOffline(tables[0][0], tables[0][1], fac, seed, initial_runs=3)
    used = nsrc
    for it, row in enumerate(D[p + 'rows']):
        nfit = int(D[p + 'fits_per_iter'][it])
        _ForcedGP.queue = thetas[used:used + nfit]
        used += nfit
        got = smbo.iterate()
        assert not _ForcedGP.queue
        assert got == int(row), 'iteration %d: oracle picked row %d, reference %d' % (it, got, row)
        if fac.w is not None:
            np.testing.assert_allclose(np.asarray(fac.w, dtype=np.float64).ravel(), D[p + 'w'][it], rtol=1e-12)
    if method == 'scot':
        np.testing.assert_allclose(np.array(fac.svm_coefs), D[p + 'svm_coef'], rtol=1e-9, atol=1e-12)
    mu, var = fac.predict(tables[0][0][:64])
    np.testing.assert_allclose(mu, D[p + 'final_mu'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var, D[p + 'final_var'], rtol=1e-6, atol=1e-12)


def test_conditional_activity_mask_matches_reference():
    """A space WITH conditions (fixture reference_v5_cond.npz, the reference's own kernels / GaussianProcess with
    ``has_conditions``): the activity mask of gp_kernels.py:21-29 in K, dK/dtheta, K(X*, X), the likelihood and
    the posterior; inactive entries arrive as NaN and are imputed to -1 (base_gp.py:148-151)."""
    D = np.load(os.path.join(HERE, 'golden', 'reference_v5_cond.npz'))
    types = D['types']
    spec = ogp.KernelSpec(types, has_conditions=True)
    Xi, Xqi = ogp.impute_inactive(D['X']), ogp.impute_inactive(D['Xq'])
    K, dK = ogp.kernel(D['thetas'][0], spec, Xi, eval_gradient=True)
    np.testing.assert_allclose(K, D['K'], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(dK, D['dK'], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(ogp.kernel(D['thetas'][0], spec, Xqi, Xi), D['Ks'], rtol=1e-13, atol=1e-15)
    assert (D['K'] == 0).mean() > 0.5
    for i, theta in enumerate(D['thetas']):
        f, g = ogp.nll(theta, spec, Xi, D['ynorm'])
        assert abs(f - D['nll_f'][i]) <= 1e-10 * abs(D['nll_f'][i])
        np.testing.assert_allclose(g, D['nll_g'][i], rtol=1e-8, atol=1e-10)
    g = ogp.OracleGP(types, 7, has_conditions=True)
    g.train(D['X'], D['y'].copy(), do_optimize=False, theta=D['theta_star'])
    mu, var = g.predict(D['Xq'])
    np.testing.assert_allclose(mu, D['mu'], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, D['var'], rtol=1e-6, atol=1e-13)


def test_sgpr_matches_reference(monkeypatch):
    """The stacking-GP baseline (tlbo/facade/stacking_gpr.py) against a run of the reference itself (fixture
    reference_v6_sgpr.npz; sources and target share configurations, so the de-duplicated configuration set is
    exercised): the prior after the K source regressors, the stacked mean / "sigma" after every iteration, the
    selected rows and the final prediction.  The oracle's GPs take the reference's fitted theta (two fits per
    regressor: on y, then on the residual)."""
    D = np.load(os.path.join(HERE, 'golden', 'reference_v6_sgpr.npz'))
    monkeypatch.setattr(ofacade, 'OracleGP', _ForcedGP)
    K, seed, types = int(D['K']), int(D['seed']), D['types']
    tables = [(D['task%d_X' % t], D['task%d_y' % t]) for t in range(K + 1)]
    thetas = [th for th in D['thetas']]
    nsrc = int(D['n_source_fits'])
    assert nsrc == 2 * K
    _ForcedGP.queue = thetas[:nsrc]
    fac = ofacade.OracleSGPR(types, [(X.copy(), y.copy()) for X, y in tables[1:]], seed, tables[0][0])
    assert not _ForcedGP.queue
    assert len(fac.configs_X) == int(D['n_configs']) < (K + 1) * 50 + 300 - 40
    np.testing.assert_array_equal(fac.configs_X, D['configs_X'])
    np.testing.assert_allclose(fac.cached_prior_mu, D['prior_mu'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(fac.cached_prior_sigma, D['prior_sigma'], rtol=1e-6, atol=1e-12)
    smbo = OracleSMBOOffline(tables[0][0], tables[0][1], fac, seed, initial_runs=3)
    used, trained = nsrc, 0
    for it, row in enumerate(D['rows']):
        if it >= 3:
            _ForcedGP.queue = thetas[used:used + 2]
            used += 2
        got = smbo.iterate()
        assert not _ForcedGP.queue
        assert got == int(row), 'iteration %d: oracle picked row %d, reference %d' % (it, got, row)
        if it >= 3:
            trained += 1
            np.testing.assert_allclose(fac.cached_stacking_mu, D['stack_mu'][it], rtol=1e-7, atol=1e-9)
            np.testing.assert_allclose(fac.cached_stacking_sigma, D['stack_sigma'][it], rtol=1e-6, atol=1e-12)
    assert trained == 5 and used == len(thetas)
    mu, var = fac.predict(tables[0][0][:64])
    np.testing.assert_allclose(mu, D['final_mu'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var, D['final_var'], rtol=1e-6, atol=1e-12)


def test_random_search_baseline_matches_reference():
    """``--methods rs`` (facade/random_surrogate.py) against a run of the reference itself (fixture
    reference_v7_rs.npz): selected rows, perfs and the facade generator's final state."""
    D = np.load(os.path.join(HERE, 'golden', 'reference_v7_rs.npz'))
    fac = ofacade.OracleRandomSearch(D['types'], [], int(D['seed']))
    smbo = OracleSMBOOffline(D['X'], D['y'], fac, int(D['seed']), initial_runs=3)
    for it, row in enumerate(D['rows']):
        assert smbo.iterate() == int(row), it
    st = fac.rng.get_state()
    assert np.array_equal(st[1], D['rng_key']) and st[2] == int(D['rng_pos'])


def test_ensemble_selection_facade_matches_reference(monkeypatch):
    """``--methods es`` (facade/obtl_es.py: greedy ensemble selection of the source weights, target weight by
    sampling) against a run of the reference itself (fixture reference_v8_es.npz): rows, the weight history to
    1e-12, the global np.random state after every iteration and the final posterior; GPs at the reference's theta."""
    D = np.load(os.path.join(HERE, 'golden', 'reference_v8_es.npz'))
    monkeypatch.setattr(ofacade, 'OracleGP', _ForcedGP)
    K, seed, types = int(D['K']), int(D['seed']), D['types']
    tables = [(D['task%d_X' % t], D['task%d_y' % t]) for t in range(K + 1)]
    thetas = [th for th in D['thetas']]
    nsrc = int(D['n_source_fits'])
    _ForcedGP.queue = thetas[:nsrc]
    fac = ofacade.OracleES(types, [(X.copy(), y.copy()) for X, y in tables[1:]], seed, literal_loops=True)
    assert not _ForcedGP.queue
    smbo = OracleSMBOOffline(tables[0][0], tables[0][1], fac, seed, initial_runs=3)
    used = nsrc
    for it, row in enumerate(D['rows']):
        nfit = int(D['fits_per_iter'][it])
        _ForcedGP.queue = thetas[used:used + nfit]
        used += nfit
        got = smbo.iterate()
        assert not _ForcedGP.queue
        assert got == int(row), 'iteration %d: oracle picked row %d, reference %d' % (it, got, row)
        np.testing.assert_allclose(np.asarray(fac.w, dtype=np.float64).ravel(), D['w'][it], rtol=1e-12, atol=1e-15)
        st = np.random.get_state()
        assert np.array_equal(np.concatenate([st[1].astype(np.int64), [st[2]]]), D['np_random_states'][it]), it
    mu, var = fac.predict(tables[0][0][:64])
    np.testing.assert_allclose(mu, D['final_mu'], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var, D['final_var'], rtol=1e-6, atol=1e-12)
