"""The oracle pinned against the REFERENCE'S OWN outputs at BASELINE sizes (tests/golden/reference_v2_large.npz,
produced by tests/golden/make_reference_golden_large.py from the unmodified reference in the build container):
20 continuous columns (configs[3], d = 20), n = 75 and n = 200 observations, and whole SMBO_OFFLINE loops with
K = 29 source surrogates x 50 observations over a 20 000-row table (configs[0] / [1] shape), n up to 75."""
import os

import numpy as np
import pytest

import oracle.facade as ofacade
from oracle import gp as ogp
from oracle.acquisition import OracleEI, OracleSMBOOffline

HERE = os.path.dirname(os.path.abspath(__file__))
R = np.load(os.path.join(HERE, 'golden', 'reference_v2_large.npz'))
TAGS = ['c20', 'rf75', 'rf200']


@pytest.mark.parametrize('tag', TAGS)
def test_kernel_and_nll_match_reference(tag):
    types, X, Xq = R['kern_%s_types' % tag], R['kern_%s_X' % tag], R['kern_%s_Xq' % tag]
    spec = ogp.KernelSpec(types)
    thetas = R['kern_%s_thetas' % tag]
    np.testing.assert_allclose(ogp.kernel(thetas[0], spec, X), R['kern_%s_K' % tag], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(ogp.kernel(thetas[0], spec, Xq[:32], X), R['kern_%s_Ks' % tag], rtol=1e-13, atol=1e-15)
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy(), do_optimize=False, theta=spec.default_theta())
    np.testing.assert_allclose(g.y, R['nll_%s_ynorm' % tag], rtol=1e-14, atol=1e-15)
    for i, theta in enumerate(thetas):
        f, grad = ogp.nll(theta, spec, X, g.y)
        assert abs(f - R['nll_%s_f' % tag][i]) <= 1e-10 * abs(R['nll_%s_f' % tag][i])
        np.testing.assert_allclose(grad, R['nll_%s_g' % tag][i], rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize('tag', TAGS)
def test_posterior_and_ei_match_reference(tag):
    types, X, Xq = R['kern_%s_types' % tag], R['kern_%s_X' % tag], R['kern_%s_Xq' % tag]
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy(), do_optimize=False, theta=R['kern_%s_thetas' % tag][1])
    mu, var = g.predict(Xq)
    np.testing.assert_allclose(mu, R['pred_%s_fixed_mu' % tag], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(var, R['pred_%s_fixed_var' % tag], rtol=1e-6, atol=1e-12)
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy(), do_optimize=False, theta=R['fit_%s_theta_star' % tag])
    mu, var = g.predict(Xq)
    np.testing.assert_allclose(mu, R['pred_%s_mu' % tag], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(var, R['pred_%s_var' % tag], rtol=1e-6, atol=1e-12)
    ei = OracleEI(g)
    ei.floor = 1e-5
    ei.update(model=g, eta=float(R['ei_%s_eta' % tag]), num_data=X.shape[0])
    np.testing.assert_allclose(ei._compute(Xq), R['ei_%s_acq' % tag], rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize('tag', ['c20', 'rf75'])
def test_fit_matches_reference(tag):
    types, X = R['kern_%s_types' % tag], R['kern_%s_X' % tag]
    spec = ogp.KernelSpec(types)
    p0 = R['fit_%s_p0' % tag]
    np.testing.assert_allclose(ogp.restart_points(spec, 7), p0[1:], rtol=1e-15, atol=0)
    g = ogp.OracleGP(types, 7)
    g.train(X, R['kern_%s_y' % tag].copy())
    f_ref = R['fit_%s_f' % tag]
    f_got = np.array([r[1] for r in g.restart_results])
    close = np.abs(f_got - f_ref) <= 1e-5 * np.maximum(1.0, np.abs(f_ref))
    assert close.sum() >= len(f_ref) - 1, (f_got, f_ref)
    assert abs(f_got.min() - f_ref.min()) <= 1e-6 * max(1.0, abs(f_ref.min()))


class _ForcedGP(ogp.OracleGP):
    queue = []

    def train(self, X, Y, do_optimize=True, theta=None):
        return super().train(X, Y, do_optimize=False, theta=_ForcedGP.queue.pop(0))


def _state_vec():
    st = np.random.get_state()
    return np.concatenate([st[1].astype(np.int64), [st[2], st[3]], [np.float64(st[4]).view(np.int64)]])


BIG = {'rgpe': (lambda t, s, seed: ofacade.OracleRGPE(t, s, seed, literal_loops=False), 14),
       'topo': (lambda t, s, seed: ofacade.OracleOBTLV(t, s, seed), 10)}


@pytest.mark.parametrize('method', sorted(BIG))
def test_big_offline_loop_matches_reference(method, monkeypatch):
    """K = 29 sources x 50 observations, 20 000-row table: rows, weights, dilution flags, the integer 30 x 30
    ranking-loss tables and the np.random stream states of the reference's own run, iteration by iteration
    (the oracle's GPs take the reference's fitted theta; pair counts by the vectorised statement, which the
    small fixtures hold equal to the literal loops).  The first iterations only (the CPU suite's time budget); the
    GPU test runs the CUDA path over the reference's FULL trajectories (76 / 40 iterations)."""
    monkeypatch.setattr(ofacade, 'OracleGP', _ForcedGP)
    make, iters = BIG[method]
    K, seed, types = int(R['big_K']), int(R['big_seed']), R['big_types']
    p = 'big_%s_' % method
    sources = [(R['big_source_X'][k].copy(), R['big_source_y'][k].copy()) for k in range(K)]
    thetas = [th for th in R[p + 'thetas']]
    nsrc = int(R[p + 'n_source_fits'])
    _ForcedGP.queue = thetas[:nsrc]
    fac = make(types, sources, seed)
    assert not _ForcedGP.queue
    np.testing.assert_allclose(fac.eta_list, R[p + 'eta_list'], rtol=1e-13)
    smbo = OracleSMBOOffline(R['big_target_X'], R['big_target_y'], fac, seed, initial_runs=3)
    used, loss_rows = nsrc, 0
    for it, row in enumerate(R[p + 'rows'][:iters]):
        nfit = int(R[p + 'fits_per_iter'][it])
        _ForcedGP.queue = thetas[used:used + nfit]
        used += nfit
        got = smbo.iterate()
        assert not _ForcedGP.queue
        assert got == int(row), 'iteration %d: oracle picked row %d, reference %d' % (it, got, row)
        np.testing.assert_array_equal(_state_vec(), R[p + 'np_random_state'][it], err_msg='np.random stream, it %d' % it)
        np.testing.assert_allclose(np.asarray(fac.w, dtype=np.float64).ravel(), R[p + 'w'][it], rtol=1e-6, atol=1e-8,
                                   err_msg='weights, iteration %d' % it)
        if method == 'rgpe' and nfit:
            np.testing.assert_array_equal(np.asarray(fac.ignored_flag, dtype=bool), R[p + 'ignored'][it])
            if nfit == 6:
                np.testing.assert_array_equal(fac.last_losses, R[p + 'losses'][loss_rows:loss_rows + 30])
            loss_rows += 30
