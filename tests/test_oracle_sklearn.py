"""Pins the oracle (numpy/scipy restatement of the reference GP path) against the one
independent implementation available offline: scikit-learn's GaussianProcessRegressor
(the reference wraps sklearn 0.21.3 through scikit-optimize).  The Hamming
kernel has no counterpart there and is checked from first principles, including the
reference's 1/l**3 gradient scaling (tlbo/model/gp_kernels.py:734)."""
import numpy as np
import pytest
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import ConstantKernel, Matern, WhiteKernel

from oracle import gp as ogp
from oracle.acquisition import OracleEI
from scipy.stats import norm


def _sk_kernel(theta, d):
    e = np.exp(theta)
    return ConstantKernel(e[0], (np.exp(-10), np.exp(2))) * Matern(e[1:1 + d], [(1e-3, 2.0)] * d, nu=2.5) + \
        WhiteKernel(e[-1], (np.exp(-25), np.exp(2)))


@pytest.mark.parametrize('n,d,seed', [(20, 3, 0), (50, 7, 1)])
def test_lml_gradient_and_predict_match_sklearn(n, d, seed):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(X.sum(1)) + 0.1 * rng.randn(n)
    types = np.zeros(d, int)
    spec = ogp.KernelSpec(types)
    theta = np.concatenate([[0.3], rng.uniform(-1.5, 0.05, d), [-6.0]])
    gpr = GaussianProcessRegressor(kernel=_sk_kernel(theta, d), optimizer=None, alpha=0.0).fit(X, y)
    lml_s, grad_s = gpr.log_marginal_likelihood(theta, eval_gradient=True)
    lml_o, grad_o = ogp.log_marginal_likelihood(theta, spec, X, y, eval_gradient=True)
    assert abs(lml_s - lml_o) <= 1e-10 * abs(lml_s)
    np.testing.assert_allclose(grad_o, grad_s, rtol=1e-8, atol=1e-10)
    g = ogp.OracleGP(types, 1, normalize_y=False)
    g.train(X, y, do_optimize=False, theta=theta)
    Xq = rng.rand(500, d)
    mu_s, sd_s = gpr.predict(Xq, return_std=True)
    mu_o, var_o = g.predict(Xq)
    np.testing.assert_allclose(mu_o.ravel(), mu_s, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(var_o.ravel(), np.maximum(sd_s ** 2, 1e-10), rtol=1e-6, atol=1e-12)


def test_hamming_kernel_and_its_quirky_gradient():
    rng = np.random.RandomState(3)
    types = np.array([3, 0, 2, 0])
    spec = ogp.KernelSpec(types)
    X = rng.rand(12, 4)
    X[:, 0] = rng.randint(0, 3, 12)
    X[:, 2] = rng.randint(0, 2, 12)
    theta = np.array([0.2, -0.3, -0.8, -0.5, -0.1, -5.0])      # [c, l_cont x2, l_cat x2, noise]
    K, dK = ogp.kernel(theta, spec, X, eval_gradient=True)
    c, lc, lk, noise = spec.split(theta)
    # value: first principles
    for i in range(12):
        for j in range(12):
            r = np.sqrt((((X[i, [1, 3]] - X[j, [1, 3]]) / lc) ** 2).sum()) * np.sqrt(5)
            ham = np.exp(-((X[i, [0, 2]] != X[j, [0, 2]]) / (2 * lk ** 2)).sum())
            want = c * (1 + r + r * r / 3) * np.exp(-r) * ham + (noise if i == j else 0.0)
            assert abs(K[i, j] - want) <= 1e-13 * max(1, abs(want))
    # gradient: continuous and amplitude / noise components are true derivatives (finite differences);
    # the categorical components are the true d/dlog(l) scaled by 1/l (the reference's quirk)
    eps = 1e-6
    for h in range(spec.H):
        tp, tm = theta.copy(), theta.copy()
        tp[h] += eps
        tm[h] -= eps
        fd = (ogp.kernel(tp, spec, X) - ogp.kernel(tm, spec, X)) / (2 * eps)
        if h in (3, 4):
            np.testing.assert_allclose(dK[:, :, h], fd / lk[h - 3], rtol=1e-5, atol=1e-8)
        else:
            np.testing.assert_allclose(dK[:, :, h], fd, rtol=1e-5, atol=1e-8)


def test_ei_closed_form_and_edge_cases():
    class M(object):
        def predict_marginalized_over_instances(self, X):
            return X[:, :1].copy(), np.maximum(X[:, 1:2].copy(), 0.0)
    ei = OracleEI(M())
    with pytest.raises(ValueError):
        ei(np.zeros((2, 2)))
    ei.update(eta=0.3)
    X = np.array([[0.1, 0.04], [0.5, 1e-10], [0.3, 0.0], [2.0, 0.25]])
    f = ei(X)
    s = np.sqrt(X[:, 1])
    for i in range(4):
        if s[i] == 0:
            assert f[i, 0] == 0.0
        else:
            z = (0.3 - X[i, 0]) / s[i]
            assert abs(f[i, 0] - ((0.3 - X[i, 0]) * norm.cdf(z) + s[i] * norm.pdf(z))) < 1e-15
