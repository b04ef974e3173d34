"""Oracle: the reference's Gaussian-process surrogate, restated in numpy/scipy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations are relative
to ``/root/reference``.

Input contract (replaces ConfigSpace): ``X`` is ``float64[n, d]``; ``types`` is
an integer vector of length ``d`` with 0 for a continuous / integer / constant
dimension and the number of choices for a categorical one, exactly what
``tlbo/model/util_funcs.py:14-55`` (``get_types``) produces.

theta layout (log space), as produced by the reference's kernel tree
``cov_amp * (matern * hamming) + noise`` (``tlbo/model/model_builder.py:28-72``):
``[log c, log l_cont (x d_cont), log l_cat (x d_cat), log sigma2]``.
"""
import math

import numpy as np
import scipy.linalg
import scipy.optimize
import scipy.spatial.distance

VERY_SMALL_NUMBER = 1e-10  # tlbo/utils/constants.py:11

# tlbo/model/model_builder.py:32-33,44-45,52-53,58-59
_C_BOUNDS = (np.exp(-10), np.exp(2))
_LS_BOUNDS = (np.exp(-6.754111155189306), np.exp(0.0858637988771976))
_NOISE_BOUNDS = (np.exp(-25), np.exp(2))
_C_INIT, _LS_INIT, _NOISE_INIT = 2.0, 1.0, 1e-8


class KernelSpec(object):
    """Shape of the kernel tree built by ``create_gp_model``
    (tlbo/model/model_builder.py:28-72)."""

    def __init__(self, types, has_conditions=False):
        self.has_conditions = bool(has_conditions or getattr(types, 'has_conditions', False))
        types = np.asarray(types)
        self.types = types
        self.d = len(types)
        self.cont_dims = np.nonzero(types == 0)[0]   # model_builder.py:38
        self.cat_dims = np.nonzero(types != 0)[0]    # model_builder.py:39
        self.d_cont = len(self.cont_dims)
        self.d_cat = len(self.cat_dims)
        if self.d_cont == 0 and self.d_cat == 0:
            raise ValueError()                       # model_builder.py:71-72
        self.H = 1 + self.d_cont + self.d_cat + 1

    def default_theta(self):
        return np.log(np.array([_C_INIT] + [_LS_INIT] * (self.d_cont + self.d_cat) + [_NOISE_INIT]))

    def log_bounds(self):
        """``kernel.bounds`` = log of the hyper-parameter bounds, in theta order."""
        b = [_C_BOUNDS] + [_LS_BOUNDS] * (self.d_cont + self.d_cat) + [_NOISE_BOUNDS]
        return np.log(np.array(b, dtype=np.float64))

    def split(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        e = np.exp(theta)
        c = e[0]
        ls_cont = e[1:1 + self.d_cont]
        ls_cat = e[1 + self.d_cont:1 + self.d_cont + self.d_cat]
        noise = e[-1]
        return c, ls_cont, ls_cat, noise


def impute_inactive(X):
    """tlbo/model/base_gp.py:148-151."""
    X = X.copy()
    X[~np.isfinite(X)] = -1
    return X


def _matern52(Xc, Yc, ls, eval_gradient):
    """Matern nu=2.5: tlbo/model/gp_kernels.py:433-496."""
    if Yc is None:
        dists = scipy.spatial.distance.pdist(Xc / ls, metric='euclidean')
    else:
        dists = scipy.spatial.distance.cdist(Xc / ls, Yc / ls, metric='euclidean')
    K = dists * math.sqrt(5)
    K = (1. + K + K ** 2 / 3.0) * np.exp(-K)
    if Yc is None:
        K = scipy.spatial.distance.squareform(K)
        np.fill_diagonal(K, 1)
    if not eval_gradient:
        return K
    # gp_kernels.py:468-496.  anisotropic <=> more than one length scale
    if len(ls) > 1:
        D = (Xc[:, np.newaxis, :] - Xc[np.newaxis, :, :]) ** 2 / (ls ** 2)
    else:
        D = scipy.spatial.distance.squareform(dists ** 2)[:, :, np.newaxis]
    tmp = np.sqrt(5 * D.sum(-1))[..., np.newaxis]
    K_gradient = 5.0 / 3.0 * D * (tmp + 1) * np.exp(-tmp)
    if len(ls) <= 1:
        K_gradient = K_gradient[:, :].sum(-1)[:, :, np.newaxis]
    return K, K_gradient


def _hamming(Xk, Yk, ls, eval_gradient):
    """HammingKernel: tlbo/model/gp_kernels.py:706-736 (gradient carries the
    reference's 1/l**3 scaling, a deliberate reproduction of its quirk)."""
    if Yk is None:
        Yk = Xk
    indicator = np.expand_dims(Xk, axis=1) != Yk
    K = (-1 / (2 * ls ** 2) * indicator).sum(axis=2)
    K = np.exp(K)
    if not eval_gradient:
        return K
    if len(ls) > 1:
        grad = (np.expand_dims(K, axis=-1) * np.array(indicator, dtype=np.float32))
    else:
        grad = np.expand_dims(K * np.sum(indicator, axis=2), axis=-1)
    grad = grad * (1 / ls ** 3)
    return K, grad


def kernel(theta, spec, X, Y=None, eval_gradient=False, _plain=False):
    """k(X, Y) of ``c * (matern * hamming) + white`` and, optionally, dK/dtheta.

    Composition rules: Product ``gp_kernels.py:251-256``, Sum ``:307-313``,
    ConstantKernel ``:367-383``, WhiteKernel ``:631-650``; column selection via
    ``operate_on`` ``:41-79``; the conditional-activity mask (``:21-29``) when
    ``spec.has_conditions``.
    """
    c, ls_cont, ls_cat, noise = spec.split(theta)
    X = np.atleast_2d(X)
    n = X.shape[0]
    m = n if Y is None else Y.shape[0]
    if Y is not None and eval_gradient:
        raise ValueError("Gradient can only be evaluated when Y is None.")
    if getattr(spec, 'has_conditions', False) and not _plain:
        # gp_kernels.py:21-29 + :49-56: the top-level kernel computes the activity mask over ALL columns and hands
        # it to every leaf, each of which multiplies its K (and dK) by it (:465, :566, :639, :720) -- the mask is
        # 0/1, so the product carries it once: K = active * (c * matern * hamming + white)
        X_cond = X <= -1
        Y_cond = X_cond if Y is None else (np.atleast_2d(Y) <= -1)
        active = ~((np.expand_dims(X_cond, axis=1) != Y_cond).any(axis=2))
        r = kernel(theta, spec, X, Y, eval_gradient, _plain=True)
        if eval_gradient:
            return r[0] * active, r[1] * active[:, :, np.newaxis]
        return r * active

    K2 = np.ones((n, m))
    grads2 = []
    mat = ham = None
    if spec.d_cont > 0:
        Xc = X[:, spec.cont_dims].reshape([-1, spec.d_cont])
        Yc = None if Y is None else Y[:, spec.cont_dims].reshape([-1, spec.d_cont])
        r = _matern52(Xc, Yc, ls_cont, eval_gradient)
        mat, mat_g = r if eval_gradient else (r, None)
    if spec.d_cat > 0:
        Xk = X[:, spec.cat_dims].reshape([-1, spec.d_cat])
        Yk = None if Y is None else Y[:, spec.cat_dims].reshape([-1, spec.d_cat])
        r = _hamming(Xk, Yk, ls_cat, eval_gradient)
        ham, ham_g = r if eval_gradient else (r, None)

    if mat is not None and ham is not None:
        K2 = mat * ham
        if eval_gradient:
            grads2 = np.dstack((mat_g * ham[:, :, np.newaxis], ham_g * mat[:, :, np.newaxis]))
    elif mat is not None:
        K2 = mat
        grads2 = mat_g
    else:
        K2 = ham
        grads2 = ham_g

    K1 = np.full((n, m), c)
    K = K1 * K2
    if Y is None:
        K = K + noise * np.eye(n)
    else:
        K = K + np.zeros((n, m))
    if not eval_gradient:
        return K
    K1_g = np.full((n, n, 1), c)
    prod_g = np.dstack((K1_g * K2[:, :, np.newaxis], grads2 * K1[:, :, np.newaxis]))
    white_g = noise * np.eye(n)[:, :, np.newaxis]
    return K, np.dstack((prod_g, white_g))


def kernel_diag(theta, spec, X):
    """``kernel_.diag(X)``: product diag = c, white diag = sigma2 (noise stays in
    ``kernel_`` because the reference passes ``noise=None``, gp.py:160)."""
    c, _, _, noise = spec.split(theta)
    return np.full(X.shape[0], c) + noise


def log_marginal_likelihood(theta, spec, X, y, eval_gradient=True):
    """sklearn 0.21.3 ``GaussianProcessRegressor.log_marginal_likelihood``
    (alpha = 0, single output).  A failed Cholesky returns ``-inf`` (and a zero
    gradient) exactly as sklearn does; ``gp.py:183-186`` then maps it to 1e25."""
    if eval_gradient:
        K, K_gradient = kernel(theta, spec, X, eval_gradient=True)
    else:
        K = kernel(theta, spec, X)
    try:
        L = scipy.linalg.cholesky(K, lower=True)
    except np.linalg.LinAlgError:
        return (-np.inf, np.zeros_like(theta)) if eval_gradient else -np.inf
    y_train = y[:, np.newaxis]
    alpha = scipy.linalg.cho_solve((L, True), y_train)
    lml = -0.5 * np.einsum("ik,ik->k", y_train, alpha)
    lml -= np.log(np.diag(L)).sum()
    lml -= K.shape[0] / 2 * np.log(2 * np.pi)
    lml = lml.sum(-1)
    if not eval_gradient:
        return lml
    tmp = np.einsum("ik,jk->ijk", alpha, alpha)
    tmp -= scipy.linalg.cho_solve((L, True), np.eye(K.shape[0]))[:, :, np.newaxis]
    grad = 0.5 * np.einsum("ijl,ijk->kl", tmp, K_gradient).sum(-1)
    return lml, grad


# ---- priors (tlbo/model/gp_base_prior.py) -----------------------------------

def lognormal_lnprob(theta_log, sigma=1.0):
    """LognormalPrior(mean=0): gp_base_prior.py:389-407 via Prior.lnprob :35-51."""
    v = np.exp(theta_log)
    if v <= 0:
        return -1e25
    return -(math.log(v)) ** 2 / (2 * sigma ** 2) - math.log(np.sqrt(2 * np.pi) * sigma * v)


def lognormal_grad(theta_log, sigma=1.0):
    """gp_base_prior.py:427-449 via Prior.gradient :108-123."""
    v = np.exp(theta_log)
    if v <= 0:
        return 0
    s2 = sigma ** 2
    return -(s2 + math.log(v)) / (s2 * v) * v


def horseshoe_lnprob(theta_log, scale=0.1):
    """HorseshoePrior: gp_base_prior.py:289-311."""
    v = np.exp(theta_log)
    if v == 0:
        return np.inf
    a = math.log(1 + 3.0 * (scale ** 2 / v ** 2))
    return math.log(a + VERY_SMALL_NUMBER)


def horseshoe_grad(theta_log, scale=0.1):
    """gp_base_prior.py:333-355 (not chain-ruled into log space -- reference quirk)."""
    v = np.exp(theta_log)
    if v == 0:
        return np.inf
    s2 = scale ** 2
    a = -(6 * s2)
    b = 3 * s2 + v ** 2
    b *= math.log(3 * s2 * v ** (-2) + 1)
    b = max(b, 1e-14)
    return a / b


def nll(theta, spec, X, y):
    """``GaussianProcess._nll`` (tlbo/model/gp.py:164-197) with the priors of
    ``_get_all_priors(add_bound_priors=False)`` (base_gp.py:92-132): lognormal on
    theta[0], horseshoe(0.1) on theta[-1], nothing on the length scales."""
    theta = np.asarray(theta, dtype=np.float64)
    lml, grad = log_marginal_likelihood(theta, spec, X, y, eval_gradient=True)
    grad = np.array(grad, dtype=np.float64)
    lml += lognormal_lnprob(theta[0])
    grad[0] += lognormal_grad(theta[0])
    lml += horseshoe_lnprob(theta[-1])
    grad[-1] += horseshoe_grad(theta[-1])
    if not np.isfinite(lml).all() or not np.all(np.isfinite(grad)):
        return 1e25, np.zeros(theta.shape)
    return -lml, -grad


def restart_points(spec, facade_seed, n_opt_restarts=10):
    """The 10 sampled L-BFGS-B start points of ``GaussianProcess._optimize``
    (gp.py:211-239) for a model built by
    ``build_model('gp', cs, np.random.RandomState(facade_seed))``.

    Draw order: ``model_builder.py:97`` consumes one ``randint(0, 10000)`` from
    the shared rng for the GP's own seed; then per theta dimension, in order:
    lognormal prior -> ``rng.lognormal`` (shared rng, gp_base_prior.py:424),
    no prior -> ``gp.rng.uniform(lo, hi, 10)`` (GP rng), horseshoe ->
    ``abs(rng.standard_cauchy(10))`` then ``abs(rng.randn() * lamda * scale)``
    (shared rng, gp_base_prior.py:327-329).  All returned in log space.
    """
    rng = np.random.RandomState(facade_seed)
    gp_seed = rng.randint(low=0, high=10000)
    gp_rng = np.random.RandomState(gp_seed)
    lb = spec.log_bounds()
    dim_samples = []
    for dim in range(spec.H):
        if dim == 0:
            s = np.log(rng.lognormal(mean=0, sigma=1.0, size=n_opt_restarts))
        elif dim == spec.H - 1:
            lamda = np.abs(rng.standard_cauchy(size=n_opt_restarts))
            s = np.log(np.abs(rng.randn() * lamda * 0.1))
        else:
            s = gp_rng.uniform(low=lb[dim][0], high=lb[dim][1], size=(n_opt_restarts,))
        if np.any(~np.isfinite(s)):
            raise ValueError('Sample %s from prior contains infinite values!' % s)
        dim_samples.append(s.flatten())
    return np.vstack(dim_samples).transpose()


class OracleGP(object):
    """``GaussianProcess`` (tlbo/model/gp.py) + ``BaseGP`` (base_gp.py) +
    ``AbstractModel.train/predict`` (base_model.py:93-200,244-249) over skopt's
    GaussianProcessRegressor semantics."""

    def __init__(self, types, facade_seed, normalize_y=True, n_opt_restarts=10, has_conditions=False):
        self.types = np.asarray(types)
        self.spec = KernelSpec(self.types, has_conditions or getattr(types, 'has_conditions', False))
        self.facade_seed = facade_seed
        self.normalize_y = normalize_y
        self.n_opt_restarts = n_opt_restarts
        self.var_threshold = 10 ** -5     # base_model.py:86
        self.is_trained = False
        self.hypers = np.empty((0,))
        self.n_ll_evals = 0
        self.theta0 = self.spec.default_theta()

    # -- AbstractModel.train checks: base_model.py:108-114
    def train(self, X, Y, do_optimize=True, theta=None):
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Feature mismatch: X should have %d features, but has %d' % (len(self.types), X.shape[1]))
        if X.shape[0] != Y.shape[0]:
            raise ValueError('X.shape[0] (%s) != y.shape[0] (%s)' % (X.shape[0], Y.shape[0]))
        return self._train(X, Y, do_optimize, theta)

    def _fit_at(self, theta):
        """skopt ``GaussianProcessRegressor.fit`` with ``optimizer=None``:
        L = chol(K), alpha = K^-1 y, K_inv = L^-T L^-1."""
        K = kernel(theta, self.spec, self.X)
        self.L = scipy.linalg.cholesky(K, lower=True)
        self.alpha = scipy.linalg.cho_solve((self.L, True), self.y)
        L_inv = scipy.linalg.solve_triangular(self.L.T, np.eye(self.L.shape[0]))
        self.K_inv = L_inv.dot(L_inv.T)
        self.theta = np.array(theta, dtype=np.float64)

    def _train(self, X, y, do_optimize=True, theta=None):
        """gp.py:98-151."""
        X = impute_inactive(X)
        if self.normalize_y:
            # base_gp.py:48-64
            self.mean_y = np.mean(y)
            self.std_y = np.std(y)
            if self.std_y == 0:
                self.std_y = 1
            y = (y - self.mean_y) / self.std_y
        y = y.flatten()
        self.X, self.y = X, y
        theta0 = np.array(self.theta0 if theta is None else theta, dtype=np.float64)
        n_tries = 10
        for i in range(n_tries):
            try:
                self._fit_at(theta0)
                break
            except np.linalg.LinAlgError:
                # gp.py:133-140: add 1 to the noise level and retry
                e = np.exp(theta0)
                e[-1] += 1
                theta0 = np.log(e)
        if do_optimize:
            self.hypers = self._optimize(theta0)
            self._fit_at(self.hypers)
        else:
            self.hypers = theta0
        self.is_trained = True
        return self

    def _optimize(self, theta0):
        """gp.py:199-248: 1 + 10 L-BFGS-B runs with scipy defaults; first strict
        minimum of f wins."""
        log_bounds = [(b[0], b[1]) for b in self.spec.log_bounds()]
        p0 = [theta0]
        if self.n_opt_restarts > 0:
            p0 += list(restart_points(self.spec, self.facade_seed, self.n_opt_restarts))
        theta_star, f_opt_star = None, np.inf
        self.restart_results = []
        for start_point in p0:
            def fun(t):
                self.n_ll_evals += 1
                return nll(t, self.spec, self.X, self.y)
            theta, f_opt, info = scipy.optimize.fmin_l_bfgs_b(fun, start_point, bounds=log_bounds)
            self.restart_results.append((np.array(theta), float(f_opt), info))
            if f_opt < f_opt_star:
                f_opt_star = f_opt
                theta_star = theta
        return theta_star

    def predict(self, X_test):
        """base_model.py:156-200 -> gp.py:250-305 -> skopt ``predict(return_std=True)``."""
        if len(X_test.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X_test.shape))
        if X_test.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X_test.shape[1]))
        if not self.is_trained:
            raise Exception('Model has to be trained first!')
        X_test = impute_inactive(X_test)
        K_trans = kernel(self.theta, self.spec, X_test, self.X)
        mu = K_trans.dot(self.alpha)
        y_var = kernel_diag(self.theta, self.spec, X_test)
        y_var = y_var - np.einsum("ki,kj,ij->k", K_trans, K_trans, self.K_inv)
        y_var[y_var < 0] = 0.0
        std = np.sqrt(y_var)
        var = std ** 2                                     # gp.py:293
        var = np.clip(var, VERY_SMALL_NUMBER, np.inf)      # gp.py:297
        if self.normalize_y:
            mu = mu * self.std_y + self.mean_y             # base_gp.py:86-90
            var = var * self.std_y ** 2
        return mu.reshape((-1, 1)), var.reshape((-1, 1))

    def predict_marginalized_over_instances(self, X):
        """base_model.py:244-249 (no instance features)."""
        mean, var = self.predict(X)
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var
