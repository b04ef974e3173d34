"""Gaussian-process surrogate with the reference's ``train`` / ``predict`` interface
(reference tlbo/model/gp.py ``GaussianProcess``, base_gp.py ``BaseGP``, base_model.py
``AbstractModel``), backed by the batched CUDA fit / predict kernels.

The fitted state (scaled inputs, alpha, L^-1, theta) lives in a slot of a device-resident
``SurrogatePack``; nothing numerical runs on the host.  ``fit_models`` trains any number of
surrogates in ONE batched launch (one CTA per surrogate x restart) -- facades use it for
the K source GPs and for the target + cross-validation GPs of an iteration."""
import numpy as np

from tlbo_b200.utils.constants import VERY_SMALL_NUMBER

# kernel hyper-parameter bounds / defaults: reference tlbo/model/model_builder.py:28-72
C_BOUNDS = (np.exp(-10), np.exp(2))
LS_BOUNDS = (np.exp(-6.754111155189306), np.exp(0.0858637988771976))
NOISE_BOUNDS = (np.exp(-25), np.exp(2))
C_INIT, LS_INIT, NOISE_INIT = 2.0, 1.0, 1e-8
HORSESHOE_SCALE = 0.1


class KernelInfo(object):
    """Descriptor of the kernel tree ``cov_amp * (matern52 * hamming) + white``.
    Exposes ``theta`` / ``bounds`` (log space) like an sklearn kernel object."""

    def __init__(self, types):
        types = np.asarray(types)
        self.cont_dims = np.nonzero(types == 0)[0].astype(int)
        self.cat_dims = np.nonzero(types != 0)[0].astype(int)
        n_ls = len(self.cont_dims) + len(self.cat_dims)
        if n_ls == 0:
            raise ValueError()
        self.n_dims = 1 + n_ls + 1
        self.theta = np.log(np.array([C_INIT] + [LS_INIT] * n_ls + [NOISE_INIT], dtype=np.float64))
        self.bounds = np.log(np.array([C_BOUNDS] + [LS_BOUNDS] * n_ls + [NOISE_BOUNDS], dtype=np.float64))

    def __repr__(self):
        return 'ConstantKernel * (Matern(nu=2.5)[%d] * Hamming[%d]) + WhiteKernel, theta=%s' % (
            len(self.cont_dims), len(self.cat_dims), np.array2string(self.theta, precision=4))


def _impute_inactive(X):
    """base_gp.py:148-151."""
    X = np.array(X, dtype=np.float64)
    X[~np.isfinite(X)] = -1
    return X


class GaussianProcess(object):
    """Drop-in for the reference's ``GaussianProcess`` on the B200 path.

    Parameters follow reference gp.py:62-75; ``prior_rng`` is the RandomState shared with
    the priors by ``create_gp_model`` (model_builder.py:35,59) and ``seed`` the model's own.
    """

    def __init__(self, configspace, types, bounds, kernel=None, alpha=0, normalize_y=True, n_opt_restarts=10,
                 seed=42, prior_rng=None):
        self.configspace = configspace
        self.types = np.asanyarray(types)        # keeps TypesArray.has_conditions
        self.bounds = bounds
        self.kernel = kernel if kernel is not None else KernelInfo(self.types)
        self._theta0 = self.kernel.theta.copy()
        self.alpha = alpha
        self.normalize_y = normalize_y
        self.n_opt_restarts = n_opt_restarts
        self.seed = seed
        self._rng = None                      # created on first use (only the start-point draws need it)
        self._prior_rng = prior_rng
        self.var_threshold = 10 ** -5          # base_model.py:86
        self.hypers = np.empty((0,))
        self.is_trained = False
        self.mean_y_, self.std_y_ = 0.0, 1.0
        self.n_train_ = 0
        self.fit_info_ = None
        self._pack = None
        self._slot = 0
        self._p0_override = None     # facades share one set of start points (same seed -> same draws)

    @property
    def rng(self):
        if self._rng is None:
            self._rng = np.random.RandomState(self.seed)
        return self._rng

    @property
    def prior_rng(self):
        if self._prior_rng is None:
            self._prior_rng = np.random.RandomState(self.seed)
        return self._prior_rng

    def fresh_copy(self):
        """An untrained surrogate identical to a newly built one (default theta, same seed
        and shared start points) without re-running the factory: what a facade needs six
        times per iteration."""
        import copy
        m = copy.copy(self)
        m.kernel = copy.copy(self.kernel)
        m.kernel.theta = self._theta0.copy()
        m.hypers = np.empty((0,))
        m.is_trained = False
        m.fit_info_ = None
        m._pack, m._slot = None, 0
        m._rng = None
        return m

    # ---- pack placement ------------------------------------------------------------
    def bind(self, pack, slot):
        """Place this surrogate's fitted state in ``pack`` slot ``slot``."""
        self._pack, self._slot = pack, int(slot)
        return self

    def _view(self):
        from tlbo_b200.ops import pack_view
        return pack_view(self._pack, self._slot, 1)

    # ---- reference interface -------------------------------------------------------
    def train(self, X, Y, do_optimize=True):
        """base_model.py:93-137 checks, then gp.py:98-151 on the device."""
        _check_train_args(self, X, Y)
        fit_models([self], [X], [Y], do_optimize=do_optimize)
        return self

    def predict(self, X):
        """base_model.py:156-200 -> gp.py:250-305: ``(mean[N,1], var[N,1])``."""
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X.shape[1]))
        if not self.is_trained:
            raise Exception('Model has to be trained first!')
        from tlbo_b200 import ops
        mu, var = ops.gp_predict(self._view(), self.types, _impute_inactive(X))
        mean = ops.d2h(mu[0]).reshape((-1, 1))
        var = ops.d2h(var[0]).reshape((-1, 1))
        if len(mean.shape) == 1:
            mean = mean.reshape((-1, 1))
        return mean, var

    def predict_marginalized_over_instances(self, X):
        """base_model.py:202-249 without instance features: predict + the 1e-5 floor."""
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X.shape[1]))
        mean, var = self.predict(X)
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var

    # ---- restart start points (gp.py:211-239) ---------------------------------------
    def _start_points(self):
        """Default theta + ``n_opt_restarts`` samples: lognormal prior on the amplitude,
        uniform in the log bounds for length scales (model rng), horseshoe on the noise
        (both priors draw from the rng shared at construction)."""
        if self._p0_override is not None:
            return self._p0_override
        p0 = [self.kernel.theta.copy()]
        R = self.n_opt_restarts
        if R > 0:
            cols = []
            H = self.kernel.n_dims
            for dim in range(H):
                lo, hi = self.kernel.bounds[dim]
                if dim == 0:
                    s = np.log(self.prior_rng.lognormal(mean=0.0, sigma=1.0, size=R))
                elif dim == H - 1:
                    lamda = np.abs(self.prior_rng.standard_cauchy(size=R))
                    s = np.log(np.abs(self.prior_rng.randn() * lamda * HORSESHOE_SCALE))
                else:
                    s = self.rng.uniform(low=lo, high=hi, size=(R,))
                if np.any(~np.isfinite(s)):
                    raise ValueError('Sample %s from prior contains infinite values!' % s)
                cols.append(np.asarray(s).flatten())
            p0 += list(np.vstack(cols).transpose())
        return np.array(p0, dtype=np.float64)


def _check_train_args(model, X, Y):
    if len(X.shape) != 2:
        raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
    if X.shape[1] != len(model.types):
        raise ValueError('Feature mismatch: X should have %d features, but has %d' % (len(model.types), X.shape[1]))
    if X.shape[0] != Y.shape[0]:
        raise ValueError('X.shape[0] (%s) != y.shape[0] (%s)' % (X.shape[0], Y.shape[0]))


class _LazyFitInfo(object):
    """Per-restart results of a batched fit, read back from the device on first access only."""

    def __init__(self, handle):
        self._dev = {'restart_theta': handle.rt, 'restart_f': handle.rf, 'restart_info': handle.ri}
        self._host = {}

    def host(self, key):
        if key not in self._host:
            from tlbo_b200 import ops
            self._host[key] = ops.d2h(self._dev[key])
        return self._host[key]


class _FitInfoView(object):
    """``fit_info_`` of one surrogate: ``['restart_theta' | 'restart_f' | 'restart_info']``."""

    def __init__(self, lazy, b):
        self._lazy, self._b = lazy, b

    def __getitem__(self, key):
        return self._lazy.host(key)[self._b]


def fit_models(models, Xs, Ys, do_optimize=True):
    """Train several ``GaussianProcess`` objects in one batched device launch (see
    ``fit_models_async``)."""
    return fit_models_async(models, Xs, Ys, do_optimize)()


def fit_prepared_async(models, prep):
    """``fit_models_async`` for a batch whose host half was done by ``ops.prepare_jobs`` (one native call
    instead of a few dozen numpy calls per surrogate): ``models`` are fresh surrogates bound to consecutive
    slots of one pack and sharing their start points; returns ``finish()``."""
    from tlbo_b200 import ops
    m0 = models[0]
    p0 = m0._start_points()
    for b, mdl in enumerate(models):
        assert mdl._pack is m0._pack and mdl._slot == m0._slot + b and mdl.normalize_y
        assert mdl._start_points() is p0 or np.array_equal(mdl._start_points(), p0)
        mdl.mean_y_, mdl.std_y_ = float(prep.ymean[b]), float(prep.ystd[b])
        mdl.n_train_ = int(prep.n[b])          # known now: the kernels' shared-memory plans are sized by it before finish()
    lb = m0.kernel.bounds
    h = ops.gp_fit_batched(None, None, m0.types, p0, lb[:, 0], lb[:, 1], None, None, m0._pack, slot0=m0._slot,
                           sync=False, prepared=prep)

    snap = {}

    def snapshot_async():
        """Enqueue pinned copies of the fit's status and fitted theta NOW (behind the fit on the current stream):
        ``finish`` then reads those instead of the pack -- it may run after the pack slots have been refitted and
        never waits for work enqueued later."""
        import torch
        st_h = torch.empty(h.status.shape, dtype=h.status.dtype, pin_memory=True)
        th_dev = m0._pack.theta[m0._slot:m0._slot + len(models)]
        th_h = torch.empty(th_dev.shape, dtype=th_dev.dtype, pin_memory=True)
        st_h.copy_(h.status, non_blocking=True)
        th_h.copy_(th_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        snap.update(st=st_h, th=th_h, ev=ev)
        ops.STATS.d2h += st_h.numel() * st_h.element_size() + th_h.numel() * 8

    def finish():
        if snap:
            snap['ev'].synchronize()
            st, thetas = snap['st'].numpy(), snap['th'].numpy()
        else:
            st = h.finish()
            thetas = ops.d2h(m0._pack.theta[m0._slot:m0._slot + len(models)])
        lazy = _LazyFitInfo(h)
        for b, mdl in enumerate(models):
            if st[b] == 1:
                raise np.linalg.LinAlgError('Cholesky of the kernel matrix failed at the selected hyper-parameters')
            if st[b] != 0:
                raise RuntimeError('GP fit failed: no finite optimum (status %d)' % st[b])
            mdl.hypers = thetas[b].copy()
            mdl.kernel.theta = thetas[b].copy()
            mdl.is_trained = True
            mdl.n_train_ = int(prep.n[b])
            mdl.fit_info_ = _FitInfoView(lazy, b)
        return models
    finish.snapshot_async = snapshot_async
    return finish


def fit_models_async(models, Xs, Ys, do_optimize=True):
    """Enqueue the batched fit and return ``finish()``; host work done between the two
    overlaps the device fit.

    Models bound to consecutive slots of one pack are fitted in place; unbound models get a
    private one-slot pack.  Models whose restart start points differ are launched in
    separate groups (the C ABI shares ``p0`` across a batch)."""
    from tlbo_b200 import ops
    jobs = []
    for mdl, X, Y in zip(models, Xs, Ys):
        _check_train_args(mdl, X, Y)
        X = _impute_inactive(X)
        y = np.asarray(Y, dtype=np.float64)
        if mdl.normalize_y:
            # base_gp.py:48-64
            mdl.mean_y_ = float(np.mean(y))
            mdl.std_y_ = float(np.std(y))
            if mdl.std_y_ == 0:
                mdl.std_y_ = 1
            y = (y - mdl.mean_y_) / mdl.std_y_
        else:
            mdl.mean_y_, mdl.std_y_ = 0.0, 1.0
        y = y.flatten()
        p0 = mdl._start_points() if do_optimize else mdl.kernel.theta.copy()[None, :]
        if mdl._pack is None or mdl._pack.ncap < X.shape[0]:
            mdl.bind(ops.SurrogatePack(1, X.shape[0], X.shape[1]), 0)
        jobs.append((mdl, X, y, p0))

    # surrogates beyond the one-CTA panel kernels' reach (SCoT's stacked GP): the device-wide large-n path
    from tlbo_b200.model import gp_large
    large = [j for j in jobs if j[1].shape[0] > gp_large.LARGE_N]
    jobs = [j for j in jobs if j[1].shape[0] <= gp_large.LARGE_N]

    # group: same pack, consecutive slots, identical start points, same types
    groups = []
    for job in jobs:
        mdl, X, y, p0 = job
        if groups:
            g = groups[-1]
            last = g[-1][0]
            if (last._pack is mdl._pack and mdl._slot == last._slot + 1 and p0.shape == g[0][3].shape
                    and (p0 is g[0][3] or np.array_equal(p0, g[0][3])) and np.array_equal(mdl.types, g[0][0].types)):
                g.append(job)
                continue
        groups.append([job])

    handles = []
    for g in groups:
        m0 = g[0][0]
        lb = m0.kernel.bounds
        handles.append(ops.gp_fit_batched(
            [j[1] for j in g], [j[2] for j in g], m0.types, g[0][3], lb[:, 0], lb[:, 1],
            [j[0].mean_y_ for j in g], [j[0].std_y_ for j in g], m0._pack, slot0=m0._slot,
            do_optimize=do_optimize, sync=False))

    def finish():
        for mdl, X, y, p0 in large:
            theta, results, bumps = gp_large.fit_large(mdl, X, y, p0, do_optimize=do_optimize)
            mdl.hypers = np.array(theta, dtype=np.float64)
            mdl.kernel.theta = np.array(theta, dtype=np.float64)
            mdl.is_trained = True
            mdl.n_train_ = X.shape[0]
            mdl.fit_info_ = {'restart_theta': np.array([r[0] for r in results]),
                             'restart_f': np.array([r[1] for r in results]),
                             'restart_info': np.array([[r[2], 0, 0, bumps] for r in results], dtype=np.int32),
                             'path': 'large-n (cuSOLVER / cuBLAS through torch, scipy L-BFGS-B)'}
        for g, h in zip(groups, handles):
            m0 = g[0][0]
            st = h.finish()
            thetas = ops.d2h(m0._pack.theta[m0._slot:m0._slot + len(g)])
            lazy = _LazyFitInfo(h)
            for b, (mdl, X, y, p0) in enumerate(g):
                if st[b] == 1:
                    # the reference's final GaussianProcessRegressor.fit raises (gp.py:146)
                    raise np.linalg.LinAlgError(
                        'Cholesky of the kernel matrix failed at the selected hyper-parameters')
                if st[b] != 0:
                    raise RuntimeError('GP fit failed: no finite optimum (status %d)' % st[b])
                mdl.hypers = thetas[b].copy()
                mdl.kernel.theta = thetas[b].copy()
                mdl.is_trained = True
                mdl.n_train_ = X.shape[0]
                mdl.fit_info_ = _FitInfoView(lazy, b)
        return models
    return finish
