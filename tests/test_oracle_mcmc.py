"""CPU tests of the GP-MCMC oracle (oracle/gp_mcmc.py) and of the host half of the product's
sampler (``stretch_move_tables``).

emcee is absent here and the reference ships no golden vectors for this path, so the sampler's
restatement is pinned against first principles:
  * the oracle's log-probability equals sklearn's ``log_marginal_likelihood`` + the priors;
  * the restated stretch move samples a known Gaussian target correctly;
  * the product's pre-drawn random tables, replayed on the host with the oracle's
    log-probability, reproduce the oracle's chain BIT FOR BIT (same stream consumption, same
    arithmetic) -- the contract the device kernel is then tested against on the GPU."""
import numpy as np
import pytest

from oracle import gp as ogp
from oracle import gp_mcmc as omc

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])


def _data(rng, n, types):
    d = len(types)
    X = rng.rand(n, d)
    for j in range(d):
        if types[j] > 0:
            X[:, j] = rng.randint(0, types[j], size=n)
    y = np.sin(3 * X[:, 3 % d]) + 0.5 * X[:, 6 % d] ** 2 + 0.3 * X[:, 0] + 0.05 * rng.randn(n)
    return X, (y - y.mean()) / y.std()


def test_ll_is_sklearn_lml_plus_priors():
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel
    types = np.zeros(4, dtype=int)
    spec = ogp.KernelSpec(types)
    rng = np.random.RandomState(0)
    X, y = _data(rng, 30, types)
    k = ConstantKernel(2.0, (np.exp(-10), np.exp(2))) * Matern(np.ones(4), [ogp._LS_BOUNDS] * 4, nu=2.5) \
        + WhiteKernel(1e-8, (np.exp(-25), np.exp(2)))
    gpr = GaussianProcessRegressor(kernel=k, optimizer=None, alpha=0).fit(X, y)
    for _ in range(5):
        th = np.array([rng.uniform(-1, 1.5)] + list(rng.uniform(-2, 0.08, 4)) + [rng.uniform(-12, -3)])
        want = gpr.log_marginal_likelihood(th) + ogp.lognormal_lnprob(th[0]) + ogp.horseshoe_lnprob(th[-1])
        assert omc.ll(th.copy(), spec, X, y) == pytest.approx(want, rel=1e-10)
    # outside a Tophat bound -> -inf; clamping to [-50, 50] mutates the argument
    th = np.array([0.0, 0.2, -1, -1, -1, -5.0])          # length scale above exp(0.0858)
    assert omc.ll(th, spec, X, y) == -np.inf
    th = np.array([0.0, -1, -1, -1, -1, -60.0])
    assert omc.ll(th, spec, X, y) == -np.inf and th[-1] == -50


def test_stretch_move_samples_a_gaussian():
    rs = np.random.RandomState(7)
    cov = np.array([[2.0, 0.6], [0.6, 0.5]])
    ic = np.linalg.inv(cov)

    def f(x):
        return -0.5 * x @ ic @ x
    c, _, _ = omc.emcee_run(f, rs.randn(20, 2), 200, rs)
    samples = []
    for _ in range(300):
        c, _, na = omc.emcee_run(f, c, 5, rs)
        samples.append(c.copy())
    S = np.concatenate(samples)
    assert np.allclose(np.cov(S.T), cov, atol=0.25)
    assert np.allclose(S.mean(0), 0, atol=0.2)


def test_walkers_independent_rejects_degenerate_ensembles():
    rs = np.random.RandomState(0)
    assert omc.walkers_independent(rs.randn(10, 3))
    bad = rs.randn(10, 3)
    bad[:, 2] = 1.0
    assert not omc.walkers_independent(bad)
    with pytest.raises(ValueError):
        omc.emcee_run(lambda x: 0.0, bad, 1, rs)


def _replay(tables, p0, logprob):
    """The device kernel's algorithm on the host: pre-drawn tables + a log-probability."""
    sidx, cidx, zz, fac, logu = tables
    pos = np.array(p0)
    lp = np.array([logprob(pos[i]) for i in range(len(pos))])
    nacc = np.zeros(len(pos), dtype=np.int64)
    for h in range(sidx.shape[0]):
        new = []
        for k in range(sidx.shape[1]):
            s, c = sidx[h, k], cidx[h, k]
            q = pos[c] - (pos[c] - pos[s]) * zz[h, k]
            new.append((s, q, logprob(q)))
        for k, (s, q, lq) in enumerate(new):
            with np.errstate(invalid='ignore'):
                if logu[h, k] < fac[h, k] + lq - lp[s]:
                    pos[s], lp[s] = q, lq
                    nacc[s] += 1
    return pos, lp, nacc


@pytest.mark.parametrize('types,n', [(RF_TYPES, 12), (np.zeros(3, dtype=int), 20)])
def test_pre_drawn_tables_reproduce_the_oracle_chain(types, n):
    from tlbo_b200.model.gp_mcmc import stretch_move_tables
    spec = ogp.KernelSpec(types)
    rng = np.random.RandomState(3)
    X, y = _data(rng, n, types)
    m = omc.OracleGPMCMC(types, 4465)
    W, H = m.n_mcmc_walkers, spec.H
    assert W == 3 * H + (3 * H) % 2
    p0 = m._sample_p0()

    def fn(th):
        return omc.ll(th, spec, X, y)
    r1 = np.random.RandomState(11)
    r2 = np.random.RandomState(11)
    want = omc.emcee_run(fn, p0, 12, r1)
    tables = stretch_move_tables(r2, W, H, 12)
    got = _replay(tables, p0, fn)
    assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1]) and np.array_equal(want[2], got[2])
    assert r1.get_state()[2] == r2.get_state()[2] and np.array_equal(r1.get_state()[1], r2.get_state()[1])
    assert want[2].sum() > 0          # the chain moved


def test_oracle_model_marginalises_over_walkers():
    types = np.array([0, 0, 0, 2, 2])
    rng = np.random.RandomState(0)
    X, y = _data(rng, 18, types)
    m = omc.OracleGPMCMC(types, 4465, chain_length=8, burnin_steps=8)
    m.train(X, y)
    assert len(m.models) == m.n_mcmc_walkers == 22
    assert m.n_ll_evals == 22 * (8 + 8 + 2)
    mu, var = m.predict(X[:5])
    assert mu.shape == (5, 1) and var.shape == (5, 1) and np.all(var >= np.finfo(float).eps)
    mus = np.array([g.predict(X[:5])[0].ravel() for g in m.models])
    vs = np.array([g.predict(X[:5])[1].ravel() for g in m.models])
    assert np.allclose(mu.ravel(), mus.mean(0)) and np.allclose(var.ravel(), mus.var(0) + vs.mean(0))
    # second call on the same object: no burn-in, chain continues from p0
    e0 = m.n_ll_evals
    m.train(X, y)
    assert m.n_ll_evals - e0 == 22 * (8 + 1)
