"""Mints tests/golden/reference_v1.npz by running the REFERENCE'S OWN modules.

Runs in the build container only (it needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_reference_golden.py

The reference is imported unmodified from /root/reference.  Its absent third-party imports
(scikit-optimize, ConfigSpace, pyrfr, emcee) are satisfied by the stand-ins under
tests/golden/refshims/ (see the README there: scikit-learn's kernels / regressor under
scikit-optimize's names with its K_inv_ variance form, value classes for ConfigSpace), and the
three numpy aliases the reference still uses and numpy 2 removed (np.int, np.float, np.mat) are
restored.  Everything recorded below is computed by reference code:

  kern_*   tlbo/model/gp_kernels.py kernels built by model_builder.create_gp_model: K, dK/dtheta,
           K(X*, X), diag -- RF-shaped space (2 categorical + 7 numeric columns) and 5 continuous
  nll_*    GaussianProcess._nll (value, gradient; priors included) at several theta
  fit_*    GaussianProcess._optimize: the 11 L-BFGS-B start points, per-restart optimum, theta*
  pred_*   GaussianProcess.predict / predict_marginalized_over_instances at theta*
  ei_*     acquisition.EI over those predictions
  slsqp_*  utils/scipy_solver.scipy_solve inputs / weights
  mcmc_*   model/gp_mcmc.GaussianProcessMCMC: _ll (Tophat bound priors, -inf outside the box), the initial
           ensemble drawn from the priors, the walkers' positions after short chains (two consecutive
           train calls: with and without burn-in), marginalised prediction.  The stretch-move sampler
           itself comes from the emcee stand-in (= the oracle's restatement): sampler NOT pinned
  online_* framework/online/smbo_online.SMBO_ONLINE + optimizer/ei_optimization.InterleavedLocalAndRandomSearch
           (LocalSearch, RandomSearch, ChallengerList) over RGPE / OBTLV: picked configurations, perfs, weights,
           fitted theta, np.random states.  ConfigSpace's sampling / neighbourhood streams come from the
           stand-in (= the oracle's restatement): those two streams NOT pinned
  loop_*   framework/smbo_offline.SMBO_OFFLINE driving facade/rgpe.RGPE, facade/topo.TransBO_RGPE,
           facade/topo_variant1.OBTLV, facade/topo_variant3.TOPO_V3, facade/obtl.OBTL, facade/tst.TST, facade/pogpe.POGPE and facade/notl.NoTL: selected rows, weights,
           dilution flags, the bootstrap ranking-loss tables (captured from the np.argmin calls of
           rgpe.py:108 / topo.py), every fitted theta in fit order, and the state of the global
           np.random stream after every iteration
"""
import collections
import contextlib
import io
import os
import sys

for _v in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
    os.environ[_v] = '1'          # as the reference's own driver does (tools/offline_benchmark.py:3-8)

import numpy as np  # noqa: E402
import scipy.optimize            # noqa: F401  (imported before the aliases are restored)
import scipy.stats               # noqa: F401
import sklearn.gaussian_process  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = '/root/reference'
sys.dont_write_bytecode = True   # /root/reference is read-only
np.int, np.float, np.mat = int, float, np.asmatrix
sys.path.insert(0, os.path.join(HERE, 'refshims'))
sys.path.insert(0, REFERENCE)

from tlbo.acquisition_function.acquisition import EI                       # noqa: E402
from tlbo.config_space import (CategoricalHyperparameter, Configuration, ConfigurationSpace, Constant,  # noqa: E402
                               UniformFloatHyperparameter, UniformIntegerHyperparameter)
from tlbo.facade.notl import NoTL                                           # noqa: E402
from tlbo.facade.obtl import OBTL                                           # noqa: E402
from tlbo.facade.pogpe import POGPE                                         # noqa: E402
from tlbo.facade.rgpe import RGPE                                           # noqa: E402
from tlbo.facade.topo import TransBO_RGPE                                   # noqa: E402
from tlbo.facade.topo_variant1 import OBTLV                                 # noqa: E402
from tlbo.facade.topo_variant3 import TOPO_V3                               # noqa: E402
from tlbo.facade.tst import TST                                             # noqa: E402
from tlbo.framework.online.smbo_online import SMBO_ONLINE                   # noqa: E402
from tlbo.framework.smbo_offline import SMBO_OFFLINE                        # noqa: E402
import tlbo.model.gp as ref_gp                                              # noqa: E402
from tlbo.model.model_builder import build_model                            # noqa: E402
from tlbo.utils.scipy_solver import scipy_solve                             # noqa: E402

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])


def rf_space():
    """The random_forest space of tools/../space_instance.py:11-35 by shape: 2 categorical (2 choices),
    3 numeric, 4 constants -> types [2, 2, 0 x 7]."""
    cs = ConfigurationSpace()
    cs.add_hyperparameters([
        CategoricalHyperparameter('criterion', ['gini', 'entropy']),
        CategoricalHyperparameter('bootstrap', ['True', 'False']),
        Constant('max_depth', 0), UniformFloatHyperparameter('max_features', 0, 1),
        Constant('min_weight_fraction_leaf', 0), Constant('max_leaf_nodes', 0),
        UniformIntegerHyperparameter('min_samples_split', 2, 20),
        UniformIntegerHyperparameter('min_samples_leaf', 1, 20), Constant('min_impurity_decrease', 0)])
    cs.get_default_configuration = lambda: Configuration(cs, vector=np.full(9, -7.0))
    return cs


def cont_space(d):
    cs = ConfigurationSpace()
    cs.add_hyperparameters([UniformFloatHyperparameter('x%d' % i, 0, 1) for i in range(d)])
    cs.get_default_configuration = lambda: Configuration(cs, vector=np.full(d, -7.0))
    return cs


def data(rng, n, types, task=0):
    d = len(types)
    X = rng.rand(n, d)
    for j in range(d):
        if types[j] > 0:
            X[:, j] = rng.randint(0, types[j], size=n)
    if d == 9:
        X[:, [2, 4, 5, 8]] = 0.0
    y = (np.sin(3 * X[:, 3 % d] + 0.3 * task) + 0.5 * X[:, 6 % d] ** 2 + 0.3 * X[:, 0] * (1 + 0.2 * task)
         + 0.05 * rng.randn(n))
    return X, y


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def gp_level(out):
    rng = np.random.RandomState(20261020)
    for tag, cs, types, n in (('rf', rf_space(), RF_TYPES, 24), ('c5', cont_space(5), np.zeros(5, int), 40)):
        X, y = data(rng, n, types)
        Xq, _ = data(rng, 64, types)
        Xq[:5] = X[:5]
        model = build_model('gp', cs, np.random.RandomState(7))
        H = len(model.kernel.theta)
        thetas = np.array([np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.5, 0.05, H - 2),
                                           [rng.uniform(-9, -4)]]) for _ in range(4)])
        # ---- kernels (gp_kernels.py) at fixed theta
        k = model.kernel.clone_with_theta(thetas[0])
        K, dK = k(X, eval_gradient=True)
        out.update({'kern_%s_types' % tag: types, 'kern_%s_X' % tag: X, 'kern_%s_y' % tag: y,
                    'kern_%s_Xq' % tag: Xq, 'kern_%s_thetas' % tag: thetas, 'kern_%s_K' % tag: K,
                    'kern_%s_dK' % tag: dK, 'kern_%s_Ks' % tag: k(Xq, X), 'kern_%s_diag' % tag: k.diag(Xq)})
        # ---- _nll (gp.py:164-197) with the priors of base_gp.py:92-132
        quiet(model._train, X, y.copy(), do_optimize=False)
        model._all_priors = model._get_all_priors(add_bound_priors=False)
        vals = [model._nll(th.copy()) for th in thetas]
        out['nll_%s_f' % tag] = np.array([v[0] for v in vals])
        out['nll_%s_g' % tag] = np.array([v[1] for v in vals])
        out['nll_%s_ynorm' % tag] = model.gp.y_train_.copy()
        # ---- predict at a fixed theta (gp.py:250-305 over the regressor's K_inv_ form)
        model.gp.kernel.theta = thetas[1]
        model.gp.fit(X, model.gp.y_train_.copy())
        model.is_trained = True
        mu, var = model.predict(Xq)
        out['pred_%s_fixed_mu' % tag], out['pred_%s_fixed_var' % tag] = mu, var
        out['pred_%s_fixed_L' % tag] = model.gp.L_.copy()
        out['pred_%s_fixed_alpha' % tag] = model.gp.alpha_.copy()
        out['pred_%s_fixed_Kinv' % tag] = model.gp.K_inv_.copy()
        # ---- full train (11 restarts of scipy L-BFGS-B through _nll)
        model = build_model('gp', cs, np.random.RandomState(7))
        rec = []
        orig = ref_gp.optimize.fmin_l_bfgs_b

        def spy(func, x0, **kw):
            r = orig(func, x0, **kw)
            rec.append((np.array(x0, dtype=np.float64), np.array(r[0]), float(r[1]), int(r[2]['funcalls'])))
            return r
        ref_gp.optimize.fmin_l_bfgs_b = spy
        try:
            quiet(model.train, X, y.copy())
        finally:
            ref_gp.optimize.fmin_l_bfgs_b = orig
        out['fit_%s_p0' % tag] = np.array([r[0] for r in rec])
        out['fit_%s_theta' % tag] = np.array([r[1] for r in rec])
        out['fit_%s_f' % tag] = np.array([r[2] for r in rec])
        out['fit_%s_funcalls' % tag] = np.array([r[3] for r in rec])
        out['fit_%s_theta_star' % tag] = np.array(model.hypers)
        mu, var = model.predict(Xq)
        out['pred_%s_mu' % tag], out['pred_%s_var' % tag] = mu, var
        mu2, var2 = model.predict_marginalized_over_instances(Xq)
        out['pred_%s_marg_mu' % tag], out['pred_%s_marg_var' % tag] = mu2, var2
        # ---- EI (acquisition.py:130-177) over the model (plain-SMBO floor 1e-5, base_model.py:244-249)
        ei = EI(model)
        eta = float(np.min(y))
        ei.update(model=model, eta=eta, num_data=n)
        out['ei_%s_eta' % tag] = eta
        out['ei_%s_acq' % tag] = ei._compute(Xq)
        # __call__ path: Configuration objects -> array -> _compute, NaN -> -max
        out['ei_%s_acq_call' % tag] = ei([Configuration(cs, vector=x) for x in Xq])


def mcmc_level(out):
    rng = np.random.RandomState(20261021)
    for tag, cs, types, n in (('rf', rf_space(), RF_TYPES, 20), ('c4', cont_space(4), np.zeros(4, int), 16)):
        X, y = data(rng, n, types)
        Xq, _ = data(rng, 48, types)
        Xq[:4] = X[:4]
        model = build_model('gp_mcmc', cs, np.random.RandomState(11))
        model.burnin_steps, model.chain_length = 12, 12
        H = len(model.kernel.theta)
        out.update({'mcmc_%s_types' % tag: types, 'mcmc_%s_X' % tag: X, 'mcmc_%s_y' % tag: y, 'mcmc_%s_Xq' % tag: Xq,
                    'mcmc_%s_walkers' % tag: model.n_mcmc_walkers})
        quiet(model.train, X, y.copy())
        out['mcmc_%s_hypers1' % tag] = np.array(model.hypers)
        mu, var = model.predict(Xq)
        out['mcmc_%s_mu1' % tag], out['mcmc_%s_var1' % tag] = mu, var
        # _ll at in-box, out-of-box and clipped points (the model is trained: self.gp / priors exist)
        thetas = np.array([np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.5, 0.05, H - 2),
                                           [rng.uniform(-9, -4)]]) for _ in range(4)])
        thetas[2, 1] = 0.5          # length scale above its bound -> -inf
        thetas[3, -1] = -60.0       # clipped to -50, still below the noise bound -> -inf
        out['mcmc_%s_ll_thetas' % tag] = thetas
        out['mcmc_%s_ll' % tag] = np.array([model._ll(th.copy()) for th in thetas], dtype=np.float64)
        # second train on more data: no burn-in, the ensemble continues from p0
        X2, y2 = data(rng, n + 3, types)
        quiet(model.train, X2, y2.copy())
        out['mcmc_%s_X2' % tag], out['mcmc_%s_y2' % tag] = X2, y2
        out['mcmc_%s_hypers2' % tag] = np.array(model.hypers)
        mu, var = model.predict_marginalized_over_instances(Xq)
        out['mcmc_%s_mu2' % tag], out['mcmc_%s_var2' % tag] = mu, var


def slsqp_level(out):
    rng = np.random.RandomState(11)
    for tag, n, m in (('a', 12, 4), ('b', 30, 6)):
        A = rng.randn(n, m)
        y = A @ np.abs(rng.randn(m)) + 0.3 * rng.randn(n)
        w, status = scipy_solve(np.mat(A), np.mat(y.reshape(-1, 1)), 3, debug=False)
        out['slsqp_%s_A' % tag], out['slsqp_%s_y' % tag] = A, y
        out['slsqp_%s_w' % tag], out['slsqp_%s_status' % tag] = np.array(w, dtype=np.float64), bool(status)


def loop_level(out):
    """Whole offline BO loops through the reference's SMBO_OFFLINE + facade + EI + OfflineSearch."""
    types = RF_TYPES
    N, K, n_src, seed = 300, 3, 50, 4465
    rng = np.random.RandomState(99)
    tables = [data(rng, N, types, task=t) for t in range(K + 1)]
    for t, (X, y) in enumerate(tables):
        out['loop_task%d_X' % t], out['loop_task%d_y' % t] = X, y
    out['loop_types'], out['loop_seed'], out['loop_K'] = types, seed, K

    for method, cls, iters in (('rgpe', RGPE, 12), ('trans_rgpe', TransBO_RGPE, 12), ('topo', OBTLV, 10), ('topo_v3', TOPO_V3, 10), ('obtl', OBTL, 10),
                               ('tst', TST, 8), ('pogpe', POGPE, 8), ('notl', NoTL, 9)):
        cs = rf_space()
        as_dict = lambda X, y: collections.OrderedDict((Configuration(cs, vector=x), float(v)) for x, v in zip(X, y))
        sources = [as_dict(*tables[t]) for t in range(1, K + 1)]
        target = as_dict(*tables[0])
        thetas = []
        orig_train = ref_gp.GaussianProcess._train

        def spy_train(self, X, y, do_optimize=True):
            r = orig_train(self, X, y, do_optimize)
            thetas.append(np.array(self.hypers, dtype=np.float64))
            return r
        losses, orig_argmin = [], np.argmin

        def spy_argmin(a, *args, **kw):
            if isinstance(a, list) and len(a) == K + 1 and not args and not kw:
                losses.append([float(v) for v in a])
            return orig_argmin(a, *args, **kw)
        ref_gp.GaussianProcess._train = spy_train
        np.argmin = spy_argmin
        try:
            kw = dict(surrogate_type='gp', num_src_hpo_trial=n_src)
            fac = quiet(cls, cs, sources, list(target.keys()), seed, **kw)
            n_source_fits = len(thetas)
            smbo = quiet(SMBO_OFFLINE, target, cs, fac, random_seed=seed, max_runs=iters, source_hpo_data=sources,
                         num_src_hpo_trial=n_src, surrogate_type='gp', initial_runs=3,
                         logging_dir='/tmp/tlbo_ref_logs')
            keys = list(target.keys())
            rows, states, fits_per_iter, w_hist, flags = [], [], [], [], []
            for _ in range(iters):
                before = len(thetas)
                cfg, _, _, _ = quiet(smbo.iterate)
                rows.append(keys.index(cfg))
                fits_per_iter.append(len(thetas) - before)
                st = np.random.get_state()
                states.append(np.concatenate([st[1].astype(np.int64), [st[2], st[3]], [np.float64(st[4]).view(np.int64)]]))
                w_hist.append(np.array(fac.w, dtype=np.float64).ravel() if fac.w is not None else np.zeros(K + 1))
                flags.append(np.array(getattr(fac, 'ignored_flag', [False] * K), dtype=bool))
        finally:
            ref_gp.GaussianProcess._train = orig_train
            np.argmin = orig_argmin
        p = 'loop_%s_' % method
        out[p + 'rows'] = np.array(rows)
        out[p + 'perfs'] = np.array(smbo.perfs, dtype=np.float64)
        out[p + 'thetas'] = np.array(thetas)
        out[p + 'n_source_fits'] = n_source_fits
        out[p + 'fits_per_iter'] = np.array(fits_per_iter)
        out[p + 'np_random_state'] = np.array(states)
        out[p + 'w'] = np.array(w_hist)
        out[p + 'ignored'] = np.array(flags)
        out[p + 'target_weight'] = np.array(fac.target_weight, dtype=np.float64)
        if hasattr(fac, 'hist_ws'):
            out[p + 'hist_ws'] = np.array([np.asarray(w, dtype=np.float64) for w in fac.hist_ws])
        out[p + 'losses'] = np.array(losses, dtype=np.float64 if method == 'obtl' else np.int64).reshape(-1, K + 1)
        out[p + 'eta_list'] = np.array(fac.eta_list, dtype=np.float64)
        print(method, 'rows', rows, 'fits', n_source_fits, fits_per_iter, 'loss rows', len(losses))


def online_objective(x):
    return float(np.sin(3 * x[3]) + 0.5 * x[6] ** 2 + 0.3 * x[0] + 0.2 * np.cos(5 * x[7]) * x[1])


def online_level(out):
    """SMBO_ONLINE + InterleavedLocalAndRandomSearch (LocalSearch one neighbour per acquisition call,
    sorted RandomSearch, ChallengerList) over RGPE and OBTLV.  The objective returns ``(perf, test_perf)``
    because ``SMBO_ONLINE.evaluate`` hands its value to an ``iterate`` that unpacks two
    (smbo_online.py:28-39 / smbo_offline.py:162).  num_points is the reference's hard-coded 5000."""
    types = RF_TYPES
    K, n_src, seed, iters = 3, 50, 3822, 8
    tables = [(out['loop_task%d_X' % t], out['loop_task%d_y' % t]) for t in range(1, K + 1)]
    first = np.array([1., 0., 0., 0.25, 0., 0., 0.75, 0.5, 0.])
    out['online_first'], out['online_seed'] = first, seed
    for method, cls in (('rgpe', RGPE), ('topo', OBTLV)):
        cs = rf_space()
        sources = [collections.OrderedDict((Configuration(cs, vector=x), float(v)) for x, v in zip(X, y))
                   for X, y in tables]
        c0 = Configuration(cs, vector=first)
        target = collections.OrderedDict([(c0, (online_objective(first), -1))])
        thetas = []
        orig_train = ref_gp.GaussianProcess._train

        def spy_train(self, X, y, do_optimize=True):
            r = orig_train(self, X, y, do_optimize)
            thetas.append(np.array(self.hypers, dtype=np.float64))
            return r
        ref_gp.GaussianProcess._train = spy_train
        try:
            fac = quiet(cls, cs, sources, list(target.keys()), seed, surrogate_type='gp', num_src_hpo_trial=n_src)
            n_source_fits = len(thetas)
            smbo = quiet(SMBO_ONLINE, lambda c: (online_objective(c.get_array()), -1), target, cs, fac,
                         random_seed=seed, max_runs=iters, source_hpo_data=sources, num_src_hpo_trial=n_src,
                         surrogate_type='gp', initial_runs=3, logging_dir='/tmp/tlbo_ref_logs')
            picks, states, fits_per_iter, w_hist = [], [], [], []
            for _ in range(iters):
                before = len(thetas)
                cfg, _, _, _ = quiet(smbo.iterate)
                picks.append(np.array(cfg.get_array(), dtype=np.float64))
                fits_per_iter.append(len(thetas) - before)
                st = np.random.get_state()
                states.append(np.concatenate([st[1].astype(np.int64), [st[2], st[3]], [np.float64(st[4]).view(np.int64)]]))
                w_hist.append(np.array(fac.w, dtype=np.float64).ravel())
        finally:
            ref_gp.GaussianProcess._train = orig_train
        p = 'online_%s_' % method
        out[p + 'picks'] = np.array(picks)
        out[p + 'perfs'] = np.array(smbo.perfs, dtype=np.float64)
        out[p + 'thetas'] = np.array(thetas)
        out[p + 'n_source_fits'] = n_source_fits
        out[p + 'fits_per_iter'] = np.array(fits_per_iter)
        out[p + 'np_random_state'] = np.array(states)
        out[p + 'w'] = np.array(w_hist)
        print('online', method, 'fits', fits_per_iter, 'perfs', np.round(smbo.perfs, 4))


def main():
    out = {}
    gp_level(out)
    mcmc_level(out)
    slsqp_level(out)
    loop_level(out)
    online_level(out)
    path = os.path.join(HERE, 'reference_v1.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes,', len(out), 'arrays')


if __name__ == '__main__':
    main()
