"""Oracle: the reference's MCMC-marginalised GP surrogate (BASELINE config 5), restated in
numpy/scipy.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations are relative to
``/root/reference``.

Follows ``tlbo/model/gp_mcmc.py`` (``GaussianProcessMCMC``), the ``'gp_mcmc'`` branch of
``tlbo/model/model_builder.py:74-89`` (walkers = 3 * H rounded up to even, 250 burn-in + 250
chain steps) and the prior stack of ``tlbo/model/base_gp.py:92-132`` with
``add_bound_priors=True`` (``TophatPrior``, ``tlbo/model/gp_base_prior.py:155-240``).

The sampler itself lives in an UN-VENDORED third-party dependency: ``emcee`` (unpinned in
``requirements.txt:16``; the reference uses the 3.x API -- ``get_chain``, ``run_mcmc`` returning
a ``State``).  emcee is absent from this image and from ``/root/reference``, so its published
algorithm is restated here from the emcee 3.x sources as recalled (``ensemble.py``
``EnsembleSampler.sample``, ``moves/red_blue.py`` ``RedBlueMove.propose``,
``moves/stretch.py`` ``StretchMove.get_proposal``, ``ensemble.py`` ``walkers_independent``):
the affine-invariant stretch move of Goodman & Weare (2010) with a = 2, two randomly split
halves per step, every random number drawn from ``sampler._random`` (a legacy
``RandomState``).  PARITY UNPINNED for the sampler: no golden vector of emcee exists here;
``tests/test_oracle_mcmc.py`` pins it against first principles (detailed balance on a known
Gaussian target, stream-consumption invariants).  Everything AROUND the sampler is pinned against
the reference's own ``GaussianProcessMCMC`` code, run unmodified in the build container over an
``emcee.EnsembleSampler`` stand-in that wraps ``emcee_run`` below
(``tests/golden/make_reference_golden.py`` -> ``reference_v1.npz`` ``mcmc_*``: ``_ll`` with its Tophat
bound priors, the prior-sampled initial ensemble, burn-in / continuation flow, walker models,
marginalised prediction; checked by ``tests/test_reference_golden.py::test_gp_mcmc_matches_reference``).
"""
import numpy as np
import scipy.linalg

from oracle.gp import (KernelSpec, OracleGP, kernel, _C_BOUNDS, _LS_BOUNDS, _NOISE_BOUNDS, horseshoe_lnprob, impute_inactive,
                       log_marginal_likelihood, lognormal_lnprob)


def prior_bounds_raw(spec):
    """``hps[0].bounds`` per theta dimension (original scale): base_gp.py:113."""
    return np.array([_C_BOUNDS] + [_LS_BOUNDS] * (spec.d_cont + spec.d_cat) + [_NOISE_BOUNDS], dtype=np.float64)


def tophat_lnprob(theta_log, lo, hi):
    """TophatPrior: gp_base_prior.py:196-211 via Prior.lnprob :35-51."""
    v = np.exp(theta_log)
    if v < lo or v > hi:
        return -np.inf
    return 0


def ll(theta, spec, X, y):
    """``GaussianProcessMCMC._ll`` (gp_mcmc.py:294-332); mutates ``theta`` like the reference."""
    if (theta < -50).any():
        theta[theta < -50] = -50
    if (theta > 50).any():
        theta[theta > 50] = 50
    try:
        lml = log_marginal_likelihood(theta, spec, X, y, eval_gradient=False)
    except ValueError:
        return -np.inf
    b = prior_bounds_raw(spec)
    H = spec.H
    for dim in range(H):
        # base_gp.py:116-131: the kernel's own prior first, then the Tophat bound prior
        if dim == 0:
            lml += lognormal_lnprob(theta[dim])
        elif dim == H - 1:
            lml += horseshoe_lnprob(theta[dim])
        lml += tophat_lnprob(theta[dim], b[dim][0], b[dim][1])
    if not np.isfinite(lml):
        return -np.inf
    return lml


def walkers_independent(coords):
    """emcee 3.x ``ensemble.walkers_independent`` (initial-state check of ``sample``)."""
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


def emcee_run(log_prob_fn, p0, nsteps, random, a=2.0, check_independent=True):
    """``EnsembleSampler.run_mcmc(p0, nsteps)`` with the default ``StretchMove``: returns
    ``(coords, log_prob, n_accepted)``.  ``random`` is the sampler's ``RandomState``
    (advanced in place).  ``log_prob_fn`` receives row VIEWS like emcee's map over ``p[i]``."""
    coords = np.array(p0, dtype=np.float64)
    nwalkers, ndim = coords.shape
    if check_independent and not walkers_independent(coords):
        raise ValueError("Initial state has a large condition number. "
                         "Make sure that your walkers are linearly independent for the best performance")
    log_prob = np.array([log_prob_fn(coords[i]) for i in range(nwalkers)], dtype=np.float64)
    if np.any(np.isnan(log_prob)):
        raise ValueError("The initial log_prob was NaN")
    naccept = np.zeros(nwalkers, dtype=np.int64)
    nsplits = 2
    for _ in range(nsteps):
        all_inds = np.arange(nwalkers)
        inds = all_inds % nsplits
        random.shuffle(inds)
        for split in range(nsplits):
            S1 = inds == split
            sets = [coords[inds == j] for j in range(nsplits)]
            s = sets[split]
            c = np.concatenate(sets[:split] + sets[split + 1:], axis=0)
            Ns, Nc = len(s), len(c)
            zz = ((a - 1.0) * random.rand(Ns) + 1) ** 2.0 / a
            factors = (ndim - 1.0) * np.log(zz)
            rint = random.randint(Nc, size=(Ns,))
            q = c[rint] - (c[rint] - s) * zz[:, None]
            new_log_probs = np.array([log_prob_fn(q[i]) for i in range(Ns)], dtype=np.float64)
            if np.any(np.isnan(new_log_probs)):
                raise ValueError("Probability function returned NaN")
            with np.errstate(invalid='ignore'):
                lnpdiff = factors + new_log_probs - log_prob[all_inds[S1]]
            accepted = np.log(random.rand(len(lnpdiff))) < lnpdiff
            m1 = all_inds[S1][accepted]
            coords[m1] = q[accepted]
            log_prob[m1] = new_log_probs[accepted]
            naccept[m1] += 1
    return coords, log_prob, naccept


class OracleGPMCMC(object):
    """``GaussianProcessMCMC`` as built by ``build_model('gp_mcmc', cs,
    RandomState(facade_seed))`` (model_builder.py:74-89)."""

    def __init__(self, types, facade_seed, n_mcmc_walkers=None, chain_length=250, burnin_steps=250,
                 normalize_y=True):
        self.types = np.asarray(types)
        self.spec = KernelSpec(self.types)
        self.prior_rng = np.random.RandomState(facade_seed)          # shared with the kernel priors
        self.rng = np.random.RandomState(self.prior_rng.randint(low=0, high=10000))   # model_builder.py:88
        if n_mcmc_walkers is None:
            n_mcmc_walkers = 3 * self.spec.H
            if n_mcmc_walkers % 2 == 1:
                n_mcmc_walkers += 1
        self.n_mcmc_walkers = n_mcmc_walkers
        self.chain_length = chain_length
        self.burnin_steps = burnin_steps
        self.normalize_y = normalize_y
        self.burned = False
        self.models = []
        self.is_trained = False
        self.var_threshold = 10 ** -5
        self.n_ll_evals = 0
        self.n_accepted = None

    def _sample_p0(self):
        """gp_mcmc.py:153-171: per dimension, ``sample_from_prior`` of the FIRST prior."""
        W, H = self.n_mcmc_walkers, self.spec.H
        b = prior_bounds_raw(self.spec)
        cols = []
        for dim in range(H):
            if dim == 0:       # LognormalPrior (shared rng): gp_base_prior.py:408-426
                raw = self.prior_rng.lognormal(mean=0.0, sigma=1.0, size=W)
            elif dim == H - 1:  # HorseshoePrior (shared rng): gp_base_prior.py:312-331
                lamda = np.abs(self.prior_rng.standard_cauchy(size=W))
                raw = np.abs(self.prior_rng.randn() * lamda * 0.1)
            else:              # TophatPrior (model rng): gp_base_prior.py:213-233
                raw = np.exp(self.rng.uniform(low=np.log(b[dim][0]), high=np.log(b[dim][1]), size=(W,)))
            sample = np.log(raw)                       # Prior.sample_from_prior :72-97
            if np.any(~np.isfinite(sample)):
                raise ValueError('Sample %s from prior contains infinite values!' % sample)
            cols.append(sample.flatten())
        return np.vstack(cols).transpose()

    def train(self, X, Y, do_optimize=True):
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Feature mismatch: X should have %d features, but has %d' % (len(self.types), X.shape[1]))
        if X.shape[0] != Y.shape[0]:
            raise ValueError('X.shape[0] (%s) != y.shape[0] (%s)' % (X.shape[0], Y.shape[0]))
        X = impute_inactive(X)
        y = np.asarray(Y, dtype=np.float64)
        if self.normalize_y:
            self.mean_y = np.mean(y)
            self.std_y = np.std(y)
            if self.std_y == 0:
                self.std_y = 1
            y = (y - self.mean_y) / self.std_y
        y = y.flatten()
        self.X, self.y = X, y
        theta0 = self.spec.default_theta()
        if do_optimize:
            # gp_mcmc.py:141: self.gp.fit(X, y) at the default theta (LinAlgError propagates)
            scipy.linalg.cholesky(kernel(theta0, self.spec, X), lower=True)

            def fn(theta):
                self.n_ll_evals += 1
                return ll(theta, self.spec, X, y)
            sampler_random = np.random.RandomState()
            sampler_random.set_state(self.rng.get_state())           # gp_mcmc.py:151
            nacc = np.zeros(self.n_mcmc_walkers, dtype=np.int64)
            if not self.burned:
                self.p0 = self._sample_p0()
                self.p0, _, na = emcee_run(fn, self.p0, self.burnin_steps, sampler_random)
                nacc += na
                self.burned = True
            self.p0, self.last_log_prob, na = emcee_run(fn, self.p0, self.chain_length, sampler_random)
            nacc += na
            self.n_accepted = nacc
            self.hypers = self.p0.copy()                             # sampler.get_chain()[-1]
        else:
            self.hypers = [theta0]
        self.models = []
        for sample in self.hypers:
            sample = np.array(sample, dtype=np.float64)
            sample[sample < -50] = -50
            sample[sample > 50] = 50
            model = OracleGP(self.types, 0, normalize_y=False)
            self.rng.randint(low=0, high=10000)                      # gp_mcmc.py:241
            model._train(X, y, do_optimize=False, theta=sample)
            self.models.append(model)
        if self.normalize_y:
            for model in self.models:                                # gp_mcmc.py:264-270
                model.normalize_y = True
                model.mean_y = self.mean_y
                model.std_y = self.std_y
        self.is_trained = True
        return self

    def predict(self, X_test):
        """base_model.py:156-200 -> gp_mcmc.py:384-440."""
        if len(X_test.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X_test.shape))
        if X_test.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X_test.shape[1]))
        if not self.is_trained:
            raise Exception('Model has to be trained first!')
        X_test = impute_inactive(X_test)
        mu = np.zeros([len(self.models), X_test.shape[0]])
        var = np.zeros([len(self.models), X_test.shape[0]])
        for i, model in enumerate(self.models):
            mu_tmp, var_tmp = model.predict(X_test)
            mu[i] = mu_tmp.flatten()
            var[i] = var_tmp.flatten()
        m = mu.mean(axis=0)
        v = np.var(mu, axis=0) + np.mean(var, axis=0)
        v = np.clip(v, np.finfo(v.dtype).eps, np.inf)
        return m.reshape((-1, 1)), v.reshape((-1, 1))

    def predict_marginalized_over_instances(self, X):
        mean, var = self.predict(X)
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var
