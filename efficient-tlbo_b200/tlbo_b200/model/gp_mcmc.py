"""MCMC-marginalised Gaussian-process surrogate with the reference's interface
(reference tlbo/model/gp_mcmc.py ``GaussianProcessMCMC``; built by the ``'gp_mcmc'`` branch of
tlbo/model/model_builder.py:74-89), backed by the persistent device sampler
``tlbo_gp_mcmc_chain`` and the marginalisation kernel ``tlbo_mcmc_marginalize``.

The reference samples the kernel hyper-parameters with ``emcee.EnsembleSampler`` (stretch
move).  Every random number of that chain -- the split permutation of each step, the stretch
uniforms, the partner indices and the acceptance uniforms -- is independent of the chain state,
so ``stretch_move_tables`` draws them up front from a copy of the model's legacy ``RandomState``
in emcee's order; the 250 + 250 steps then run in ONE kernel launch for any number of
surrogates, and the W walker GPs are factorised into W slots of a device pack.
Facades fit their target + cross-validation surrogates in one batched launch through
``fit_mcmc_models_async``.
"""
import numpy as np

from tlbo_b200.model.gp import (C_BOUNDS, HORSESHOE_SCALE, LS_BOUNDS, NOISE_BOUNDS, GaussianProcess, KernelInfo,
                                _check_train_args, _impute_inactive)


def stretch_move_tables(random, W, H, nsteps, a=2.0):
    """Random tables of ``nsteps`` emcee 3.x steps (``RedBlueMove.propose`` with
    ``nsplits=2, randomize_split=True`` + ``StretchMove.get_proposal``), drawn from
    ``random`` (a ``RandomState``, advanced in place) in emcee's order.  Arrays of shape
    ``[2 * nsteps, W // 2]``: active walker ids, partner walker ids, ``zz``,
    ``(H - 1) * log(zz)``, ``log(u)``."""
    Ns = W // 2
    sidx = np.empty((2 * nsteps, Ns), dtype=np.int32)
    cidx = np.empty((2 * nsteps, Ns), dtype=np.int32)
    zz = np.empty((2 * nsteps, Ns), dtype=np.float64)
    fac = np.empty((2 * nsteps, Ns), dtype=np.float64)
    logu = np.empty((2 * nsteps, Ns), dtype=np.float64)
    all_inds = np.arange(W)
    for step in range(nsteps):
        inds = all_inds % 2
        random.shuffle(inds)
        sets = (np.nonzero(inds == 0)[0], np.nonzero(inds == 1)[0])
        for split in (0, 1):
            h = 2 * step + split
            s_ids, c_ids = sets[split], sets[1 - split]
            z = ((a - 1.0) * random.rand(Ns) + 1) ** 2.0 / a
            rint = random.randint(len(c_ids), size=(Ns,))
            u = random.rand(Ns)
            sidx[h], cidx[h], zz[h] = s_ids, c_ids[rint], z
            fac[h] = (H - 1.0) * np.log(z)
            logu[h] = np.log(u)
    return sidx, cidx, zz, fac, logu


def walkers_independent(coords):
    """emcee's initial-state check (``ensemble.walkers_independent``)."""
    if not np.all(np.isfinite(coords)):
        return False
    C = coords - np.mean(coords, axis=0)[None, :]
    C_colmax = np.amax(np.abs(C), axis=0)
    if np.any(C_colmax == 0):
        return False
    C /= C_colmax
    C_colsum = np.sqrt(np.sum(C ** 2, axis=0))
    C /= C_colsum
    return np.linalg.cond(C.astype(float)) <= 1e8


class GaussianProcessMCMC(object):
    """Drop-in for the reference's ``GaussianProcessMCMC`` (emcee sampler) on the B200 path.

    ``bind(pack, slot)`` places the W walker GPs in pack slots ``[slot * W, (slot + 1) * W)``.
    """

    def __init__(self, configspace, types, bounds, kernel=None, n_mcmc_walkers=20, chain_length=50, burnin_steps=50,
                 normalize_y=True, mcmc_sampler='emcee', average_samples=False, seed=42, prior_rng=None):
        if mcmc_sampler != 'emcee':
            raise ValueError(mcmc_sampler)       # 'nuts' needs github.com/mfeurer/NUTS: outside this path
        if n_mcmc_walkers < 2 or n_mcmc_walkers % 2:
            raise ValueError('n_mcmc_walkers must be even (emcee splits the ensemble in two halves)')
        self.configspace = configspace
        self.types = np.asanyarray(types)        # keeps TypesArray.has_conditions
        self.bounds = bounds
        self.kernel = kernel if kernel is not None else KernelInfo(self.types)
        self._theta0 = self.kernel.theta.copy()
        self.n_mcmc_walkers = int(n_mcmc_walkers)
        self.chain_length = int(chain_length)
        self.burnin_steps = int(burnin_steps)
        self.normalize_y = normalize_y
        self.mcmc_sampler = mcmc_sampler
        self.average_samples = average_samples
        self.seed = seed
        self._rng = None
        self._prior_rng = prior_rng
        # state of the shared prior rng right after construction: what a freshly built model sees
        self._prior_state0 = None if prior_rng is None else prior_rng.get_state()
        self._prior_key = None if prior_rng is None else (self._prior_state0[1].tobytes(), self._prior_state0[2])
        self.burned = False
        self.p0 = None
        self.hypers = np.empty((0,))
        self.is_trained = False
        self.var_threshold = 10 ** -5
        self.mean_y_, self.std_y_ = 0.0, 1.0
        self.n_train_ = 0
        self._n_ll_evals = 0
        self.n_accepted_ = None
        self.last_log_prob_ = None
        self._pack, self._slot = None, 0
        self._shared = None          # facade-level cache of (p0, tables) for identically seeded models

    group_size = property(lambda self: self.n_mcmc_walkers)

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
        """An untrained surrogate identical to a newly built one (same seeds, hence the same
        initial ensemble and random tables, served from the shared cache)."""
        import copy
        m = copy.copy(self)
        m.kernel = copy.copy(self.kernel)
        m.kernel.theta = self._theta0.copy()
        m.hypers = np.empty((0,))
        m.is_trained, m.burned, m.p0 = False, False, None
        m._pack, m._slot = None, 0
        m._rng = None
        if self._prior_state0 is not None:
            m._prior_rng = np.random.RandomState()
            m._prior_rng.set_state(self._prior_state0)
        m._n_ll_evals = 0
        return m

    def bind(self, pack, slot):
        self._pack, self._slot = pack, int(slot)
        return self

    def _view(self):
        from tlbo_b200.ops import pack_view
        W = self.n_mcmc_walkers
        return pack_view(self._pack, self._slot * W, W)

    @property
    def models(self):
        """The per-walker GPs (reference ``self.models``) as ``GaussianProcess`` views of the
        walker slots."""
        out = []
        if not self.is_trained:
            return out
        W = self.n_mcmc_walkers
        for i in range(W):
            g = GaussianProcess(self.configspace, self.types, self.bounds, kernel=KernelInfo(self.types),
                                normalize_y=False, seed=0)
            g.kernel.theta = np.array(self.hypers[i])
            g.hypers = np.array(self.hypers[i])
            g.is_trained = True
            g.mean_y_, g.std_y_ = self.mean_y_, self.std_y_
            g.n_train_ = self.n_train_
            g.bind(self._pack, self._slot * W + i)
            out.append(g)
        return out

    # ---- prior stack ------------------------------------------------------------------
    def prior_bounds_raw(self):
        """Tophat bounds per theta dimension, original scale (base_gp.py:113-131)."""
        n_ls = self.kernel.n_dims - 2
        return np.array([C_BOUNDS] + [LS_BOUNDS] * n_ls + [NOISE_BOUNDS], dtype=np.float64)

    def _sample_p0(self):
        """gp_mcmc.py:153-171: the initial ensemble, one ``sample_from_prior`` of the first prior
        per dimension (lognormal / horseshoe draw from the rng shared with the kernel priors,
        the Tophat bound priors from the model's own rng)."""
        W, H = self.n_mcmc_walkers, self.kernel.n_dims
        b = self.prior_bounds_raw()
        cols = []
        for dim in range(H):
            if dim == 0:
                raw = self.prior_rng.lognormal(mean=0.0, sigma=1.0, size=W)
            elif dim == H - 1:
                lamda = np.abs(self.prior_rng.standard_cauchy(size=W))
                raw = np.abs(self.prior_rng.randn() * lamda * HORSESHOE_SCALE)
            else:
                raw = np.exp(self.rng.uniform(low=np.log(b[dim][0]), high=np.log(b[dim][1]), size=(W,)))
            sample = np.log(raw)
            if np.any(~np.isfinite(sample)):
                raise ValueError('Sample %s from prior contains infinite values!' % sample)
            cols.append(sample.flatten())
        return np.vstack(cols).transpose()

    def _plan_chain(self):
        """Host half of one ``_train``: the initial ensemble and the random tables of this call's
        steps, plus the model-rng bookkeeping of gp_mcmc.py:151,241.  Returns
        ``(p0 [W, H], nsteps, tables, cache_key)``."""
        W, H = self.n_mcmc_walkers, self.kernel.n_dims
        fresh = not self.burned
        key = None
        if fresh and self._shared is not None and self._rng is None:
            # identically seeded fresh models (every model a facade builds): same ensemble, same tables
            key = ('fresh', self.seed, W, H, self.burnin_steps, self.chain_length, self._prior_key)
            hit = self._shared.get(key)
            if hit is not None:
                self._rng = np.random.RandomState()
                self._rng.set_state(hit['rng_after'])
                if self._prior_rng is not None and hit.get('prior_after') is not None:
                    self._prior_rng.set_state(hit['prior_after'])
                self.burned = True
                return hit['p0'], hit['nsteps'], hit['tables'], key
        sampler_random = np.random.RandomState()
        sampler_random.set_state(self.rng.get_state())          # gp_mcmc.py:151
        if fresh:
            p0 = self._sample_p0()
            nsteps = self.burnin_steps + self.chain_length
            self.burned = True
        else:
            p0 = self.p0
            nsteps = self.chain_length
        if not walkers_independent(p0):
            raise ValueError("Initial state has a large condition number. "
                             "Make sure that your walkers are linearly independent for the best performance")
        tables = stretch_move_tables(sampler_random, W, H, nsteps)
        for _ in range(W):
            self.rng.randint(low=0, high=10000)                   # gp_mcmc.py:241 (per-walker model seeds)
        if key is not None:
            self._shared[key] = {'p0': p0, 'nsteps': nsteps, 'tables': tables, 'rng_after': self.rng.get_state(),
                                 'prior_after': None if self._prior_rng is None else self._prior_rng.get_state()}
        return p0, nsteps, tables, key

    # ---- reference interface ----------------------------------------------------------
    def train(self, X, Y, do_optimize=True):
        _check_train_args(self, X, Y)
        fit_mcmc_models_async([self], [X], [Y], do_optimize=do_optimize)()
        return self

    def _check_predict_args(self, X):
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X.shape[1]))
        if not self.is_trained:
            raise Exception('Model has to be trained first!')

    def predict(self, X):
        """base_model.py:156-200 -> gp_mcmc.py:384-440: ``(mean[N,1], var[N,1])``."""
        self._check_predict_args(X)
        from tlbo_b200 import ops
        mu, var = ops.gp_predict(self._view(), self.types, _impute_inactive(X))
        gmu, gvar = ops.mcmc_marginalize(mu, var, 1, mu.shape[0])
        return ops.d2h(gmu[0]).reshape((-1, 1)), ops.d2h(gvar[0]).reshape((-1, 1))

    def predict_marginalized_over_instances(self, X):
        mean, var = self.predict(X)
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var


def fit_mcmc_models_async(models, Xs, Ys, do_optimize=True):
    """Enqueue the chains of several ``GaussianProcessMCMC`` surrogates (ONE persistent launch
    per group of co-resident surrogates) and the factorisation of their walker GPs into the
    pack; returns ``finish()``, which synchronises once and publishes the results.

    Models bound to a shared pack keep their slots; unbound ones get a private pack."""
    from tlbo_b200 import ops
    if not models:
        return lambda: models
    W = models[0].n_mcmc_walkers
    H = models[0].kernel.n_dims
    jobs, plans = [], []
    for mdl, X, Y in zip(models, Xs, Ys):
        _check_train_args(mdl, X, Y)
        if mdl.n_mcmc_walkers != W or mdl.kernel.n_dims != H:
            raise ValueError('one batched MCMC fit needs surrogates of one shape')
        X = _impute_inactive(X)
        y = np.asarray(Y, dtype=np.float64)
        if mdl.normalize_y:
            mdl.mean_y_ = float(np.mean(y))
            mdl.std_y_ = float(np.std(y))
            if mdl.std_y_ == 0:
                mdl.std_y_ = 1
            y = (y - mdl.mean_y_) / mdl.std_y_
        else:
            mdl.mean_y_, mdl.std_y_ = 0.0, 1.0
        y = y.flatten()
        if mdl._pack is None or mdl._pack.ncap < X.shape[0]:
            mdl.bind(ops.SurrogatePack(W, X.shape[0], X.shape[1]), 0)
        jobs.append((mdl, X, y))
        if do_optimize:
            plans.append(mdl._plan_chain())
    m0 = models[0]
    B = len(jobs)
    batch = ops.upload_batch([j[1] for j in jobs], [j[2] for j in jobs])
    if do_optimize:
        # distinct tables are uploaded once each (device-cached by identity for shared ones)
        tab_ids, uniq = [], []
        for p in plans:
            for i, q in enumerate(uniq):
                if q[2] is p[2]:
                    tab_ids.append(i)
                    break
            else:
                tab_ids.append(len(uniq))
                uniq.append(p)
        nsteps = plans[0][1]
        if any(p[1] != nsteps for p in plans):
            raise ValueError('one batched MCMC fit needs chains of one length')
        tables = ops.McmcTables([p[2] for p in uniq])
        pos0 = np.stack([p[0] for p in plans])
        handle = ops.gp_mcmc_chain(batch, m0.types, W, nsteps, pos0, tables, tab_ids, m0.prior_bounds_raw(),
                                   m0._theta0)
        theta_dev = handle.pos.view(B * W, H)
    else:
        handle = None
        theta_dev = ops.h2d(np.repeat(np.stack([j[0].kernel.theta for j in jobs]), W, axis=0))
    # W walker GPs per surrogate (gp_mcmc.py:226-248), factorised at the sampled theta straight
    # into the pack: one launch when the surrogates sit in consecutive slots of one pack
    ymean = [j[0].mean_y_ for j in jobs]
    ystd = [j[0].std_y_ for j in jobs]
    consecutive = all(jobs[b][0]._pack is m0._pack and jobs[b][0]._slot == m0._slot + b for b in range(B))
    if consecutive:
        fac_all = ops.gp_factorize_walkers(batch, m0.types, theta_dev, ymean, ystd, m0._pack, m0._slot * W, W)
        fac = [fac_all[b * W:(b + 1) * W] for b in range(B)]
    else:
        fac = []
        for b, (mdl, X, y) in enumerate(jobs):
            sub = (batch[0][b:b + 1], batch[1][b:b + 1], batch[2][b:b + 1], batch[3])
            fac.append(ops.gp_factorize_walkers(sub, mdl.types, theta_dev[b * W:(b + 1) * W], ymean[b:b + 1],
                                                ystd[b:b + 1], mdl._pack, mdl._slot * W, W))

    def finish():
        if handle is not None:
            st, pos, lp, nacc = handle.finish()
        for b, (mdl, X, y) in enumerate(jobs):
            if handle is not None:
                if st[b] != 0:
                    raise np.linalg.LinAlgError('Cholesky of the kernel matrix failed at the default '
                                                'hyper-parameters (GaussianProcessRegressor.fit, gp_mcmc.py:141)')
                mdl.p0 = pos[b].copy()
                mdl.hypers = np.clip(pos[b], -50, 50)
                mdl.last_log_prob_ = lp[b].copy()
                mdl.n_accepted_ = nacc[b].copy()
                runs = 1 if plans[b][1] == mdl.chain_length else 2
                mdl._n_ll_evals += W * (plans[b][1] + runs)
            else:
                mdl.hypers = np.repeat(mdl.kernel.theta[None, :], W, axis=0)
            fst = ops.d2h(fac[b])
            if np.any(fst != 0):
                _refactorize_failed(mdl, X, y, fst)
            mdl.is_trained = True
            mdl.n_train_ = X.shape[0]
        return models
    return finish


def _refactorize_failed(mdl, X, y, fst):
    """gp.py:127-140 for walker GPs whose Cholesky failed at the sampled theta: add 1 to the noise
    level and retry, up to 10 tries (only reachable for walkers that never left an initial
    position of log-probability -inf)."""
    from tlbo_b200 import ops
    W = mdl.n_mcmc_walkers
    for i in np.nonzero(fst)[0]:
        theta = np.array(mdl.hypers[i], dtype=np.float64)
        for _ in range(10):
            e = np.exp(theta)
            e[-1] += 1
            theta = np.log(e)
            st = ops.gp_factorize_batched([X], [y], mdl.types, theta[None, :], [mdl.mean_y_], [mdl.std_y_], mdl._pack,
                                          slot0=mdl._slot * W + int(i))
            if st[0] == 0:
                break
        else:
            raise np.linalg.LinAlgError('Fail to fit GP after 10 tries!')
        mdl.hypers[i] = theta
