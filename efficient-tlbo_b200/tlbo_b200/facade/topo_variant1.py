"""OBTLV -- what ``--methods topo`` runs (reference tlbo/facade/topo_variant1.py:9-106,
tools/offline_benchmark.py:310-311): two-phase simplex-constrained minimisation of a
pairwise logistic ranking loss (sources, then sources + cross-validated target)."""
import numpy as np

from tlbo_b200.facade.base_facade import BaseFacade
from tlbo_b200.utils.scipy_solver import scipy_solve

_scale_method = 'standardize'


def _kfold_val_slices(n, n_splits=5):
    """Validation ranges of sklearn ``KFold(n_splits)`` without shuffling: the first
    ``n % n_splits`` folds hold one extra row."""
    sizes = np.full(n_splits, n // n_splits, dtype=int)
    sizes[:n % n_splits] += 1
    out, cur = [], 0
    for s in sizes:
        out.append((cur, cur + int(s)))
        cur += int(s)
    return out


class OBTLV(BaseFacade):
    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, fusion_method='idp_lc', only_source=False, **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        self.method_id = 'topo_v1'
        self.fusion_method = fusion_method
        self.build_source_surrogates(normalize=_scale_method)
        self.only_source = only_source
        self.w = np.array([1. / self.K] * self.K + [0.])
        self.min_num_y = 5
        self.hist_ws = list()
        self.iteration_id = 0
        self.target_y_range = None
        self._side = None             # CUDA stream of the phase-1 weight solve

    def batch_predict(self, X):
        """topo_variant1.py:27-35: [n, K] source means at X."""
        mu, _ = self._source_predictions(X)
        return np.ascontiguousarray(mu[:self.K].T)

    def train(self, X, y):
        K = self.K
        instance_num = X.shape[0]
        do_cv = (not self.only_source) and instance_num >= self.min_num_y
        # host halves of every build_single_surrogate of this call, then ONE batched fit
        folds = _kfold_val_slices(instance_num, 5) if do_cv else []
        if self._native_prep_ok(X, y):
            # one native call for the host half of all 1 + 5 training sets (see BaseFacade._prepare_cv_jobs)
            fit_models, launch_fit = self._prepare_cv_jobs(X, y, [(0, 0)] + list(folds), [0] * (1 + len(folds)),
                                                           _scale_method)
        else:
            jobs = [self._prepare_single_surrogate(X, y, _scale_method, K)]
            for i, (lo, hi) in enumerate(folds):
                train_idx = np.concatenate([np.arange(0, lo), np.arange(hi, instance_num)])
                jobs.append(self._prepare_single_surrogate(X[train_idx, :], y[train_idx], _scale_method, K + 1 + i))
            for j, job in enumerate(jobs):
                job[0].bind(self._pack, K + j)
            fit_models = [j[0] for j in jobs]
            launch_fit = lambda: self._fit_async(fit_models, [j[1] for j in jobs], [j[2] for j in jobs])
        import torch
        from tlbo_b200 import ops
        X_dev = self._to_device(X)                       # before the fit is enqueued
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(main)                     # X_dev and the source slots are ready
        finish_fit = launch_fit()
        self.target_surrogate = fit_models[0]
        self.last_fit_models = fit_models
        self.target_y_range = 0.5 * (np.max(y) - np.min(y))
        # Phase 1 (source weights) depends only on the SOURCE surrogates: it runs on a side
        # stream -- source predictions at X plus the SLSQP callbacks -- while the batched fit of
        # the target + CV surrogates occupies the main stream.
        if self._overlap_hook is not None:
            self._overlap_hook()                         # tie keys etc., overlapped with the fit
        with torch.cuda.stream(self._side):
            mu_src_dev, _ = self._predict_slots(0, K, X_dev)
            pred_y = np.ascontiguousarray(ops.d2h(mu_src_dev).T)
            self.learn_source_weights(pred_y, y)
        main.wait_stream(self._side)
        mu = None
        if len(folds):
            mu_cv_dev, _ = self._predict_slots(K + 1, len(folds), X_dev)
        finish_fit()
        if len(folds):
            mu = ops.d2h(mu_cv_dev)
        w_source = self.w[:K]               # a view: the in-place scaling below is the reference's
        w_target = 0.
        if do_cv:
            _t_pred_y = np.empty(instance_num)
            for i, (lo, hi) in enumerate(folds):
                _t_pred_y[lo:hi] = mu[i, lo:hi]
            _pred_y = np.c_[pred_y, _t_pred_y.reshape((-1, 1))]
            w_target = self.compute_target_weight(_pred_y, y)
            if instance_num >= 2 * self.min_num_y:
                w_target = np.max([w_target, self.w[-1]])
            w_source *= (1 - w_target)
        w_new = np.asarray(list(w_source) + [w_target])
        rho = 0.6
        self.w = rho * w_new + (1 - rho) * self.w
        w = self.w.copy()
        self.target_weight.append(w[-1])
        self.hist_ws.append(w)
        self.iteration_id += 1

    def learn_source_weights(self, pred_y, true_y):
        x, status = scipy_solve(pred_y, true_y, 3)
        if status:
            x[x < 1e-3] = 0.
            self.w[:self.K] = x

    def compute_target_weight(self, pred_y, true_y):
        x, status = scipy_solve(pred_y, true_y, 3)
        if status:
            x[x < 1e-3] = 0.
            return x[-1]
        return self.w[-1]

    def predict(self, X):
        return self.combine_predictions(X, self.fusion_method)

    def get_weights(self):
        return self.w


class TOPO_V3(OBTLV):
    """``--methods topo_v3`` (reference tlbo/facade/topo_variant3.py:9-109, tools/offline_benchmark.py:312-313):
    ONE simplex-constrained solve over the K source columns plus the cross-validated target column (the
    sources alone while there are fewer than 5 observations), the target weight floored at 0.3, no smoothing
    (rho = 1) and no ``target_weight`` log.  Same kernels as ``OBTLV``."""

    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, fusion_method='idp_lc', **kw):
        super().__init__(config_space, source_hpo_data, target_hp_configs, seed, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, fusion_method=fusion_method, **kw)
        self.method_id = 'topo_v3'

    def train(self, X, y):
        from tlbo_b200 import ops
        K = self.K
        instance_num = X.shape[0]
        do_cv = instance_num >= self.min_num_y
        folds = _kfold_val_slices(instance_num, 5) if do_cv else []
        if self._native_prep_ok(X, y):
            fit_models, launch_fit = self._prepare_cv_jobs(X, y, [(0, 0)] + list(folds), [0] * (1 + len(folds)),
                                                           _scale_method)
        else:
            jobs = [self._prepare_single_surrogate(X, y, _scale_method, K)]
            for i, (lo, hi) in enumerate(folds):
                train_idx = np.concatenate([np.arange(0, lo), np.arange(hi, instance_num)])
                jobs.append(self._prepare_single_surrogate(X[train_idx, :], y[train_idx], _scale_method, K + 1 + i))
            for j, job in enumerate(jobs):
                job[0].bind(self._pack, K + j)
            fit_models = [j[0] for j in jobs]
            launch_fit = lambda: self._fit_async(fit_models, [j[1] for j in jobs], [j[2] for j in jobs])
        X_dev = self._to_device(X)                       # before the fit is enqueued
        finish_fit = launch_fit()
        self.target_surrogate = fit_models[0]
        self.last_fit_models = fit_models
        self.target_y_range = 0.5 * (np.max(y) - np.min(y))
        if self._overlap_hook is not None:
            self._overlap_hook()
        mu_dev, _ = self._predict_slots(0, K + 1 + len(folds), X_dev)     # sources, target, CV folds: one launch
        finish_fit()
        mu = ops.d2h(mu_dev)
        pred_y = np.ascontiguousarray(mu[:K].T)
        if not do_cv:
            x, status = scipy_solve(pred_y, y, 3)         # learn_source_weights
            if status:
                x[x < 1e-3] = 0.
                self.w[:K] = x
        else:
            _t_pred_y = np.empty(instance_num)
            for i, (lo, hi) in enumerate(folds):
                _t_pred_y[lo:hi] = mu[K + 1 + i, lo:hi]
            x, status = scipy_solve(np.c_[pred_y, _t_pred_y.reshape((-1, 1))], y, 3)      # learn_weights
            if status:
                x[x < 1e-3] = 0.
            else:
                x = self.w.copy()
            w_source, w_target = x[:-1], x[-1]
            if status:
                w_target = np.max([w_target, 0.3])
                if instance_num >= 2 * self.min_num_y:
                    w_target = np.max([w_target, self.w[-1]])
                    w_source *= (1 - w_target)
            w_new = np.asarray(list(w_source) + [w_target])
            rho = 1.0
            self.w = rho * w_new + (1 - rho) * self.w
        self.hist_ws.append(self.w.copy())
        self.iteration_id += 1


class OBTL(OBTLV):
    """``--methods obtl`` (reference tlbo/facade/obtl.py:9-150, tools/offline_benchmark.py:292-293): source weights
    by the simplex-constrained pairwise-logistic solve (as ``OBTLV`` phase 1), target weight by SAMPLING -- 50
    bootstrap draws ``normal(mu, var)`` of every source surrogate and of the cross-validated target surrogate,
    each scored by the soft ranking loss ``sum_{i<j} log(1 + exp(-10 z_ij)) / n^2``; the target weight is the
    share of samples in which the target wins the arg-min.  The reference's pure-Python pair loop
    (``calculate_ranking_loss``) is the ``pair_penalty_kernel``; the draws come from the GLOBAL legacy
    ``np.random`` stream in the reference's order (sample-major, surrogate 0..K, n normals each)."""
    ensemble_size = 50

    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, fusion_method='idp_lc', **kw):
        super().__init__(config_space, source_hpo_data, target_hp_configs, seed, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, fusion_method=fusion_method, **kw)
        self.method_id = 'obtl'
        self.last_sample_losses = None

    def _source_weights_for_update(self):
        """obtl.py:46: ``w_source = self.w[:self.K]`` -- a VIEW: the in-place ``w_source *= (1 - w_target)`` below also
        scales the OLD weights that enter the exponential smoothing (kept as the reference has it)."""
        return self.w[:self.K]

    def train(self, X, y):
        import torch
        from tlbo_b200 import ops
        K = self.K
        instance_num = X.shape[0]
        do_cv = instance_num >= self.min_num_y
        folds = _kfold_val_slices(instance_num, 5) if do_cv else []
        if self._native_prep_ok(X, y):
            fit_models, launch_fit = self._prepare_cv_jobs(X, y, [(0, 0)] + list(folds), [0] * (1 + len(folds)),
                                                           _scale_method)
        else:
            jobs = [self._prepare_single_surrogate(X, y, _scale_method, K)]
            for i, (lo, hi) in enumerate(folds):
                train_idx = np.concatenate([np.arange(0, lo), np.arange(hi, instance_num)])
                jobs.append(self._prepare_single_surrogate(X[train_idx, :], y[train_idx], _scale_method, K + 1 + i))
            for j, job in enumerate(jobs):
                job[0].bind(self._pack, K + j)
            fit_models = [j[0] for j in jobs]
            launch_fit = lambda: self._fit_async(fit_models, [j[1] for j in jobs], [j[2] for j in jobs])
        X_dev = self._to_device(X)                       # before the fit is enqueued
        y_dev = ops.h2d(np.ascontiguousarray(y, dtype=np.float64))
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(main)
        finish_fit = launch_fit()
        self.target_surrogate = fit_models[0]
        self.last_fit_models = fit_models
        self.target_y_range = 0.5 * (np.max(y) - np.min(y))
        if self._overlap_hook is not None:
            self._overlap_hook()
        z_dev = None
        if do_cv:
            # obtl.py:111-117: ensemble_size x (K + 1) x n normals from the global stream (drawn while the
            # fit runs: standard normals do not depend on the posteriors)
            z_dev = self._stage.upload('z_obtl', np.random.standard_normal((self.ensemble_size, K + 1, instance_num)))
        with torch.cuda.stream(self._side):              # source side: overlaps the fit
            mu_src_dev, var_src_dev = self._predict_slots(0, K, X_dev)
            pred_y = np.ascontiguousarray(ops.d2h(mu_src_dev).T)
            self.learn_source_weights(pred_y, y)
        main.wait_stream(self._side)
        if do_cv:
            mu_cv_dev, var_cv_dev = self._predict_slots(K + 1, len(folds), X_dev)
        finish_fit()
        w_source = self._source_weights_for_update()
        w_target = 0.
        if do_cv:
            # cross-validated target posterior: fold i's model at its own validation rows (obtl.py:73-93)
            loc = torch.empty(K + 1, instance_num, dtype=torch.float64, device=X_dev.device)
            scale = torch.empty_like(loc)
            loc[:K] = mu_src_dev
            scale[:K] = var_src_dev
            for i, (lo, hi) in enumerate(folds):
                loc[K, lo:hi] = mu_cv_dev[i, lo:hi]
                scale[K, lo:hi] = var_cv_dev[i, lo:hi]
            losses = ops.pair_penalty_loss_normal(y_dev, z_dev, loc, scale)
            self.last_sample_losses = losses
            _w = np.bincount(np.argmin(losses, axis=1), minlength=K + 1).astype(np.float64)
            _w = _w / np.sum(_w)
            w_target = _w[-1]
            if instance_num >= 2 * self.min_num_y:
                w_target = np.max([w_target, self.w[-1]])
            w_source *= (1 - w_target)
        w_new = np.asarray(list(w_source) + [w_target])
        rho = 0.6
        self.w = rho * w_new + (1 - rho) * self.w
        self.hist_ws.append(self.w.copy())
        self.iteration_id += 1


class ES(OBTL):
    """``--methods es`` (reference tlbo/facade/obtl_es.py:9-214, tools/offline_benchmark.py:290-291): OBTL with the
    source weights chosen by greedy ENSEMBLE SELECTION instead of the SLSQP solve -- 50 rounds, each adding the source
    surrogate whose inclusion gives the mean prediction ``mean(ensemble + [mu_i])`` the smallest soft ranking loss
    (``calculate_ranking_loss``, the ``pair_penalty_kernel`` with zero spread: one launch scores all K candidates of
    a round); weights = selection counts / 50.  Target weight by sampling, smoothing (rho = 0.6) as OBTL -- except
    that ``w_source`` is a fresh array here, so the old weights enter the smoothing unscaled (obtl_es.py:82-97)."""

    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, fusion_method='idp_lc', **kw):
        super().__init__(config_space, source_hpo_data, target_hp_configs, seed, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, fusion_method=fusion_method, **kw)
        self.method_id = 'es'
        self._es_w_source = None
        self.last_selection = None

    def learn_source_weights(self, pred_y, true_y):
        """obtl_es.py:61-87.  ``pred_y [n, K]``: the source surrogates' means at the target points."""
        import torch
        from tlbo_b200 import ops
        base = np.ascontiguousarray(np.asarray(pred_y, dtype=np.float64).T)          # [K, n]
        K, n = base.shape
        y = np.ascontiguousarray(np.asarray(true_y, dtype=np.float64).reshape(-1))
        y_dev = ops.h2d(y)
        base_dev = ops.h2d(base)
        zero = torch.zeros(1, K, n, dtype=torch.float64, device=base_dev.device)
        run = torch.zeros(n, dtype=torch.float64, device=base_dev.device)             # running sum of the ensemble
        picks = []
        for m in range(self.ensemble_size):
            # np.mean(ensemble + [mu_i], axis=0): the rows added in order, then one division by the count
            cand = base_dev if m == 0 else (run[None, :] + base_dev) / float(m + 1)
            losses = ops.pair_penalty_loss_normal(y_dev, zero, cand.contiguous(), zero[0])
            idx = int(np.argmin(losses[0]))
            picks.append(idx)
            run = run + base_dev[idx] if m else base_dev[idx].clone()
        w_source = np.zeros(K)
        for idx in picks:
            w_source[idx] += 1
        self._es_w_source = w_source / np.sum(w_source)
        self.last_selection = picks

    def _source_weights_for_update(self):
        return self._es_w_source

    def _ensemble_args(self):
        # obtl_es.py:190-207: the target first, then the sources with a POSITIVE weight in order
        w = np.asarray(self.w, dtype=np.float64)
        skip = np.array([0 if w[i] > 0 else 1 for i in range(self.K)] + [0], dtype=np.int32)
        order = np.array([self.K] + list(range(self.K)), dtype=np.int32)
        return w, skip, order

    def predict(self, X):
        return self._predict_rows(X, 0.0)
