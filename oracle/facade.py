"""Oracle: ensemble facades (RGPE / TransBO_RGPE / OBTLV "topo" / TST / NoTL).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations are relative
to ``/root/reference``.

Source histories are passed as ``[(X_k float64[T_k, d], y_k float64[T_k]), ...]``
(the array form of the reference's ``OrderedDict{Configuration -> perf}`` after
``convert_configurations_to_array``, base_facade.py:66-71).
"""
import itertools

import numpy as np
from scipy.optimize import minimize

from oracle.gp import OracleGP, VERY_SMALL_NUMBER


def zero_one_normalization(X, lower=None, upper=None):
    """tlbo/utils/normalization.py:4-13."""
    if lower is None:
        lower = np.min(X, axis=0)
    if upper is None:
        upper = np.max(X, axis=0)
    return np.true_divide((X - lower), (upper - lower)), lower, upper


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    """tlbo/utils/normalization.py:20-28."""
    if mean is None:
        mean = np.mean(X, axis=0)
    if std is None:
        std = np.std(X, axis=0)
    return (X - mean) / std, mean, std


# ---- pair counting ----------------------------------------------------------

def rank_loss_literal(y, sampled_y, row_lo, row_hi):
    """The reference's pure-Python double loop (rgpe.py:78-82, 98-102)."""
    n = len(y)
    rank_loss = 0
    for i in range(row_lo, row_hi):
        for j in range(n):
            if (y[i] < y[j]) ^ (sampled_y[i] < sampled_y[j]):
                rank_loss += 1
    return rank_loss


def rank_loss_vectorised(y, sampled_y, row_lo, row_hi):
    """Same count, numpy broadcasting (best-effort CPU variant)."""
    y = np.asarray(y).reshape(-1)
    s = np.asarray(sampled_y).reshape(-1)
    a = y[row_lo:row_hi, None] < y[None, :]
    b = s[row_lo:row_hi, None] < s[None, :]
    return int(np.count_nonzero(a ^ b))


class OracleFacade(object):
    """``BaseFacade`` (tlbo/facade/base_facade.py:15-183).

    ``surrogate_type`` / ``mcmc_steps`` are class attributes (subclass to change them):
    ``'gp'`` builds ``OracleGP``, ``'gp_mcmc'`` the MCMC-marginalised ``OracleGPMCMC`` with
    ``mcmc_steps = (burn-in, chain)`` (reference: 250, 250 -- model_builder.py:74-89)."""
    surrogate_type = 'gp'
    mcmc_steps = None

    def _new_model(self):
        """``build_model(surrogate_type, config_space, RandomState(seed))``: base_facade.py:64-65,95."""
        if self.surrogate_type == 'gp_mcmc':
            from oracle.gp_mcmc import OracleGPMCMC
            burn, chain = self.mcmc_steps if self.mcmc_steps is not None else (250, 250)
            return OracleGPMCMC(self.types, self.random_seed, chain_length=chain, burnin_steps=burn)
        return OracleGP(self.types, self.random_seed)

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, literal_loops=False):
        self.method_id = None
        self.types = np.asarray(types)
        self.random_seed = seed
        self.K = len(source_hpo_data)
        self.num_src_hpo_trial = num_src_hpo_trial
        self.source_hpo_data = source_hpo_data
        self.source_surrogates = None
        self.target_surrogate = None
        self.var_threshold = VERY_SMALL_NUMBER
        self.w = None
        self.eta_list = list()
        self.target_weight = []
        self.literal_loops = literal_loops
        self._rank_loss = rank_loss_literal if literal_loops else rank_loss_vectorised

    def build_source_surrogates(self, normalize):
        """base_facade.py:58-91."""
        self.source_surrogates = list()
        for X, y in self.source_hpo_data:
            model = self._new_model()
            X = np.asarray(X, dtype=np.float64)[:self.num_src_hpo_trial]
            y = np.array(y, dtype=np.float64)[:self.num_src_hpo_trial]
            if normalize == 'standardize':
                if (y == y[0]).all():
                    y[0] += 1e-4
                y, _, _ = zero_mean_unit_var_normalization(y)
            elif normalize == 'scale':
                if (y == y[0]).all():
                    y[0] += 1e-4
                y, _, _ = zero_one_normalization(y)
                y = 2 * y - 1.
            else:
                raise ValueError('Invalid parameter in norm.')
            self.eta_list.append(np.min(y))
            model.train(X, y)
            self.source_surrogates.append(model)

    def build_single_surrogate(self, X, y, normalize):
        """base_facade.py:93-108 (mutates the caller's ``y`` in the all-equal case)."""
        assert normalize in ['standardize', 'scale', 'none']
        model = self._new_model()
        if normalize == 'standardize':
            if (y == y[0]).all():
                y[0] += 1e-4
            y, _, _ = zero_mean_unit_var_normalization(y)
        elif normalize == 'scale':
            if (y == y[0]).all():
                y[0] += 1e-4
            y, _, _ = zero_one_normalization(y)
        model.train(X, y)
        return model

    def predict_marginalized_over_instances(self, X):
        """base_facade.py:110-143."""
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X.shape[1]))
        mean, var = self.predict(X)
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var

    def combine_predictions(self, X, combination_method='idp_lc'):
        """base_facade.py:145-183, 'idp_lc' branch."""
        n = X.shape[0]
        mu, var = np.zeros((n, 1)), np.zeros((n, 1))
        w = self.w
        for i in range(0, self.K + 1):
            if i == self.K:
                if self.target_surrogate is None:
                    raise ValueError('Target surrogate is none.')
                _mu, _var = self.target_surrogate.predict(X)
            else:
                _mu, _var = self.source_surrogates[i].predict(X)
            mu += w[i] * _mu
            var += w[i] * w[i] * _var
        if combination_method != 'idp_lc':
            raise ValueError('Invalid combination method %s.' % combination_method)
        return mu, var


class OracleNoTL(OracleFacade):
    """tlbo/facade/notl.py."""

    def __init__(self, types, source_hpo_data, seed, **kw):
        super().__init__(types, source_hpo_data, seed, **kw)
        self.method_id = 'notl'

    def train(self, X, y):
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='standardize')

    def predict(self, X):
        return self.target_surrogate.predict(X)


class OracleRGPE(OracleFacade):
    """tlbo/facade/rgpe.py (``smooth=None``) and tlbo/facade/topo.py
    ``TransBO_RGPE`` (``smooth=0.7``): identical bootstrap ranking loss, the
    latter with the extra weight smoothing of topo.py:111-133."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, only_source=False,
                 smooth=None, literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, literal_loops)
        self.method_id = 'rgpe' if smooth is None else 'trans_rgpe'
        self.only_source = only_source
        self.smooth = smooth
        self.build_source_surrogates(normalize='standardize')
        if smooth is None:
            self.w = [1. / self.K] * self.K + [0.]
        else:
            self.w = np.array([1. / self.K] * self.K + [0.])
        self.num_sample = 30
        self.ignored_flag = [False] * self.K
        self.hist_ws = list()
        self.iteration_id = 0
        self.last_losses = None

    def train(self, X, y):
        """rgpe.py:24-139 / topo.py:24-145.  Normal draws come from the GLOBAL
        legacy ``np.random`` stream, scale = the predictive VARIANCE (rgpe.py:77)."""
        mu_list, var_list = list(), list()
        for id in range(self.K):
            mu, var = self.source_surrogates[id].predict(X)
            mu_list.append(mu)
            var_list.append(var)

        self.target_surrogate = self.build_single_surrogate(X, y, normalize='standardize')

        k_fold_num = 5
        cached_mu_list, cached_var_list = list(), list()
        instance_num = len(y)
        skip_target_surrogate = False if instance_num >= k_fold_num else True

        if not skip_target_surrogate:
            # (the instance_num < k_fold_num branch of rgpe.py:45-54 is dead code)
            fold_num = instance_num // k_fold_num
            for i in range(k_fold_num):
                row_indexs = list(range(instance_num))
                bound = (instance_num - i * fold_num) if i == (k_fold_num - 1) else fold_num
                for index in range(bound):
                    del row_indexs[i * fold_num]
                if (y[row_indexs] == y[row_indexs[0]]).all():
                    y[row_indexs[0]] += 1e-4
                model = self.build_single_surrogate(X[row_indexs, :], y[row_indexs], normalize='standardize')
                mu, var = model.predict(X)
                cached_mu_list.append(mu)
                cached_var_list.append(var)

        argmin_list = [0] * (self.K + 1)
        ranking_loss_caches = list()
        for _ in range(self.num_sample):
            ranking_loss_list = list()
            for id in range(self.K):
                sampled_y = np.random.normal(mu_list[id], var_list[id])
                ranking_loss_list.append(self._rank_loss(y, sampled_y, 0, len(y)))

            rank_loss = 0
            if not skip_target_surrogate:
                fold_num = instance_num // k_fold_num
                for fold in range(k_fold_num):
                    sampled_y = np.random.normal(cached_mu_list[fold], cached_var_list[fold])
                    bound = instance_num if fold == (k_fold_num - 1) else (fold + 1) * fold_num
                    rank_loss += self._rank_loss(y, sampled_y, fold_num * fold, bound)
            else:
                rank_loss = instance_num * instance_num
            ranking_loss_list.append(rank_loss)
            ranking_loss_caches.append(ranking_loss_list)
            argmin_task = np.argmin(ranking_loss_list)
            argmin_list[argmin_task] += 1

        ranking_loss_caches = np.array(ranking_loss_caches)
        self.last_losses = ranking_loss_caches
        if self.smooth is None:
            for id in range(self.K + 1):
                self.w[id] = argmin_list[id] / self.num_sample
        else:
            new_weights = np.zeros(self.K + 1)
            for id in range(self.K + 1):
                new_weights[id] = argmin_list[id] / self.num_sample

        threshold = sorted(ranking_loss_caches[:, -1])[int(self.num_sample * 0.95)]
        for id in range(self.K):
            median = sorted(ranking_loss_caches[:, id])[int(self.num_sample * 0.5)]
            self.ignored_flag[id] = median > threshold

        if self.smooth is None:
            if self.only_source:
                self.w[-1] = 0.
                if np.sum(self.w) == 0:
                    self.w = [1. / self.K] * self.K + [0.]
                else:
                    self.w[:-1] = np.array(self.w[:-1]) / np.sum(self.w[:-1])
        else:
            min_num_y = 5
            w_target = new_weights[-1]
            w_source = np.array(new_weights[:-1])
            if instance_num >= min_num_y:
                if instance_num >= 2 * min_num_y:
                    w_target = np.max([w_target, self.w[-1]])
                if np.sum(w_source) > 0:
                    w_source = w_source / np.sum(w_source) * (1 - w_target)
            w_new = np.array(list(w_source) + [w_target])
            rho = self.smooth
            self.w = rho * w_new + (1 - rho) * self.w

        w = self.w.copy()
        for id in range(self.K):
            if self.ignored_flag[id]:
                w[id] = 0.
        self.target_weight.append(w[-1])
        self.hist_ws.append(w)
        self.iteration_id += 1

    def predict(self, X):
        """rgpe.py:141-153."""
        mu, var = self.target_surrogate.predict(X)
        mu *= self.w[-1]
        var *= (self.w[-1] * self.w[-1])
        for i in range(0, self.K):
            if not self.ignored_flag[i]:
                mu_t, var_t = self.source_surrogates[i].predict(X)
                mu += self.w[i] * mu_t
                var += self.w[i] * self.w[i] * var_t
        return mu, var


# ---- OBTLV ("--methods topo"): pairwise logistic loss + SLSQP ----------------

def ranking_pairs(true_y):
    """Pairs (i, j) with true_y[i] > true_y[j] in the reference's enumeration
    order (utils/scipy_solver.py:15-22)."""
    pairs = list()
    for (i, j) in itertools.combinations(range(true_y.shape[0]), 2):
        if true_y[i] > true_y[j]:
            pairs.append((i, j))
        elif true_y[i] < true_y[j]:
            pairs.append((j, i))
    return pairs


def loss_func3(true_y, pred_y, literal=False):
    """``Loss_func(..., func_id=3)``: utils/scipy_solver.py:6-40."""
    true_y = np.asarray(true_y).reshape(-1)
    pred_y = np.asarray(pred_y).reshape(-1)
    pairs = ranking_pairs(true_y)
    if len(pairs) == 0:
        return 0.
    if literal:
        loss = 0.
        for (i, j) in pairs:
            loss += np.log(1 + np.exp(pred_y[j] - pred_y[i]))
        return loss / len(pairs)
    ii = np.array([p[0] for p in pairs])
    jj = np.array([p[1] for p in pairs])
    return float(np.sum(np.log(1 + np.exp(pred_y[jj] - pred_y[ii]))) / len(pairs))


def loss_der3(true_y, A, x, literal=False):
    """``Loss_der(..., func_id=3)``: utils/scipy_solver.py:43-78."""
    A = np.asarray(A)
    true_y = np.asarray(true_y).reshape(-1)
    pred_y = A.dot(np.asarray(x).reshape(-1))
    pairs = ranking_pairs(true_y)
    grad = np.zeros(A.shape[1])
    if len(pairs) == 0:
        return grad
    if literal:
        for (i, j) in pairs:
            e_z = np.exp(pred_y[j] - pred_y[i])
            grad += e_z / (1 + e_z) * (A[j] - A[i])
        return grad / len(pairs)
    ii = np.array([p[0] for p in pairs])
    jj = np.array([p[1] for p in pairs])
    e_z = np.exp(pred_y[jj] - pred_y[ii])
    s = e_z / (1 + e_z)
    return (s[:, None] * (A[jj] - A[ii])).sum(0) / len(pairs)


def scipy_solve(A, b, loss_fn=loss_func3, der_fn=loss_der3):
    """utils/scipy_solver.py:81-115 with ``loss_type=3``.  ``loss_fn`` /
    ``der_fn`` are injectable so the product path can run the same SLSQP
    driver over its device loss (the product has its own copy of this driver;
    it never imports the oracle)."""
    A = np.asarray(A)
    n, m = A.shape
    ineq_cons = {'type': 'ineq', 'fun': lambda x: np.array(x), 'jac': lambda x: np.eye(len(x))}
    eq_cons = {'type': 'eq', 'fun': lambda x: np.array([sum(x) - 1]), 'jac': lambda x: np.array([1.] * len(x))}
    x0 = np.array([1. / m] * m)

    def f(x):
        return loss_fn(b, A.dot(x))

    def f_der(x):
        return der_fn(b, A, x)

    res = minimize(f, x0, method='SLSQP', jac=f_der, constraints=[eq_cons, ineq_cons],
                   options={'ftol': 1e-8, 'disp': False})
    status = False if np.isnan(res.x).any() else True
    if not res.success and status:
        res.x[res.x < 0.] = 0.
        res.x[res.x > 1.] = 1.
        if sum(res.x) > 1.5:
            status = False
    return res.x, status


def kfold_indices(n, n_splits=5):
    """sklearn ``KFold(n_splits, shuffle=False).split``: the first ``n % k``
    folds get one extra row (used by topo_variant1.py:42-44)."""
    fold_sizes = np.full(n_splits, n // n_splits, dtype=int)
    fold_sizes[:n % n_splits] += 1
    idx = np.arange(n)
    out, cur = [], 0
    for fs in fold_sizes:
        val = idx[cur:cur + fs]
        train = np.concatenate([idx[:cur], idx[cur + fs:]])
        out.append((train, val))
        cur += fs
    return out


class OracleOBTLV(OracleFacade):
    """tlbo/facade/topo_variant1.py -- what ``--methods topo`` runs
    (tools/offline_benchmark.py:310-311)."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, fusion_method='idp_lc',
                 only_source=False, literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, literal_loops)
        self.method_id = 'topo_v1'
        self.fusion_method = fusion_method
        self.build_source_surrogates(normalize='standardize')
        self.only_source = only_source
        self.w = np.array([1. / self.K] * self.K + [0.])
        self.min_num_y = 5
        self.hist_ws = list()
        self.iteration_id = 0
        self.target_y_range = None

    def _solve(self, A, y):
        lit = self.literal_loops
        return scipy_solve(A, y, lambda b, p: loss_func3(b, p, lit), lambda b, A_, x: loss_der3(b, A_, x, lit))

    def batch_predict(self, X):
        """topo_variant1.py:27-35."""
        pred_y = None
        for i in range(0, self.K):
            mu, _ = self.source_surrogates[i].predict(X)
            pred_y = mu if pred_y is None else np.c_[pred_y, mu]
        return pred_y

    def predict_target_surrogate_cv(self, X, y):
        """topo_variant1.py:37-53."""
        _mu, _var = list(), list()
        for train_idx, val_idx in kfold_indices(X.shape[0], 5):
            X_train, X_val, y_train = X[train_idx, :], X[val_idx, :], y[train_idx]
            model = self.build_single_surrogate(X_train, y_train, normalize='standardize')
            mu, var = model.predict(X_val)
            _mu.extend(list(mu.flatten()))
            _var.extend(list(var.flatten()))
        return np.asarray(_mu), np.asarray(_var)

    def train(self, X, y):
        """topo_variant1.py:55-86."""
        instance_num = X.shape[0]
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='standardize')
        self.target_y_range = 0.5 * (np.max(y) - np.min(y))
        pred_y = self.batch_predict(X)

        x, status = self._solve(pred_y, y)            # learn_source_weights :88-92
        if status:
            x[x < 1e-3] = 0.
            self.w[:self.K] = x
        w_source = self.w[:self.K]
        w_target = 0.
        if not self.only_source:
            if instance_num >= self.min_num_y:
                _t_pred_y, _ = self.predict_target_surrogate_cv(X, y)
                _pred_y = np.c_[pred_y, _t_pred_y.reshape((-1, 1))]
                x, status = self._solve(_pred_y, y)   # compute_target_weight :94-100
                if status:
                    x[x < 1e-3] = 0.
                    w_target = x[-1]
                else:
                    w_target = self.w[-1]
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

    def predict(self, X):
        return self.combine_predictions(X, self.fusion_method)


class OracleTOPOV3(OracleOBTLV):
    """tlbo/facade/topo_variant3.py (``--methods topo_v3``, tools/offline_benchmark.py:312-313): ONE
    simplex-constrained solve over the K source columns plus the cross-validated target column, target
    weight floored at 0.3, no smoothing (rho = 1), no ``target_weight`` log."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, fusion_method='idp_lc', literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, fusion_method, False, literal_loops)
        self.method_id = 'topo_v3'

    def train(self, X, y):
        """topo_variant3.py:53-89."""
        instance_num = X.shape[0]
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='standardize')
        self.target_y_range = 0.5 * (np.max(y) - np.min(y))
        pred_y = self.batch_predict(X)
        if instance_num < self.min_num_y:
            x, status = self._solve(pred_y, y)            # learn_source_weights :91-95
            if status:
                x[x < 1e-3] = 0.
                self.w[:self.K] = x
        else:
            _pred_y, _ = self.predict_target_surrogate_cv(X, y)
            _pred_y = np.c_[pred_y, _pred_y.reshape((-1, 1))]
            x, status = self._solve(_pred_y, y)           # learn_weights :97-103
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


def obtl_ranking_loss(y_pred, y, literal=False):
    """``OBTL.calculate_ranking_loss`` / ``penalty_func`` (tlbo/facade/obtl.py:124-139):
    sum_{i<j} log(1 + exp(-10 z)) / n^2 with z = y_pred_j - y_pred_i if y_i < y_j else y_pred_i - y_pred_j."""
    n = y_pred.shape[0]
    if literal:
        ranking_loss = 0.
        for i in range(n):
            for j in range(i + 1, n):
                z = (y_pred[j] - y_pred[i]) if y[i] < y[j] else (y_pred[i] - y_pred[j])
                z *= 10.
                ranking_loss += np.log(1 + np.exp(-z))
        return ranking_loss / (n * n)
    iu, ju = np.triu_indices(n, 1)
    z = np.where(y[iu] < y[ju], y_pred[ju] - y_pred[iu], y_pred[iu] - y_pred[ju]) * 10.
    with np.errstate(over='ignore'):
        return float(np.sum(np.log(1 + np.exp(-z)))) / (n * n)


class OracleOBTL(OracleOBTLV):
    """tlbo/facade/obtl.py (``--methods obtl``): source weights by the SLSQP solve, target weight by sampling
    (``calculate_weight_by_sampling`` :95-122), rho = 0.6, no ``target_weight`` log."""
    ensemble_size = 50

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, fusion_method='idp_lc', literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, fusion_method, False, literal_loops)
        self.method_id = 'obtl'
        self.last_sample_losses = None

    def calculate_weight_by_sampling(self, X, y):
        preds = []
        for idx in range(self.K):
            _mu, _var = self.source_surrogates[idx].predict(X)
            preds.append((_mu.flatten(), _var.flatten()))
        preds.append(self.predict_target_surrogate_cv(X, y))
        _K = len(preds)
        losses = np.zeros((self.ensemble_size, _K))
        for s in range(self.ensemble_size):
            for i in range(_K):
                _mu, _var = preds[i]
                y_pred = np.random.normal(_mu, _var)
                losses[s, i] = obtl_ranking_loss(y_pred, y, self.literal_loops)
        self.last_sample_losses = losses
        _w = np.zeros(_K)
        for idx in np.argmin(losses, axis=1):
            _w[idx] += 1
        _w = _w / np.sum(_w)
        return _w[-1]

    def train(self, X, y):
        """obtl.py:37-63."""
        instance_num = X.shape[0]
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='standardize')
        self.target_y_range = 0.5 * (np.max(y) - np.min(y))
        pred_y = self.batch_predict(X)
        x, status = self._solve(pred_y, y)            # learn_source_weights :65-69
        if status:
            x[x < 1e-3] = 0.
            self.w[:self.K] = x
        w_source = self.w[:self.K]
        w_target = 0.
        if instance_num >= self.min_num_y:
            w_target = self.calculate_weight_by_sampling(X, y)
            if instance_num >= 2 * self.min_num_y:
                w_target = np.max([w_target, self.w[-1]])
            w_source *= (1 - w_target)
        w_new = np.asarray(list(w_source) + [w_target])
        rho = 0.6
        self.w = rho * w_new + (1 - rho) * self.w
        self.hist_ws.append(self.w.copy())
        self.iteration_id += 1


class OracleTST(OracleFacade):
    """tlbo/facade/tst.py (``use_metafeatures=False``)."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, only_source=False, literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, literal_loops)
        self.method_id = 'tst'
        self.only_source = only_source
        self.build_source_surrogates(normalize='scale')
        self.w = [0.75] * (self.K + 1)
        self.bandwidth = 0.45
        self.hist_ws = list()
        self.iteration_id = 0

    def train(self, X, y):
        """tst.py:30-62: discordant pairs over i<j."""
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='scale')
        n_sample = X.shape[0]
        for _id in range(self.K):
            mu, _ = self.source_surrogates[_id].predict(X)
            mu = mu.reshape(-1)
            if self.literal_loops:
                discordant, total = 0, 0
                for i in range(n_sample):
                    for j in range(i + 1, n_sample):
                        if (y[i] < y[j]) ^ (mu[i] < mu[j]):
                            discordant += 1
                        total += 1
            else:
                iu = np.triu_indices(n_sample, 1)
                a = (y[:, None] < y[None, :])[iu]
                b = (mu[:, None] < mu[None, :])[iu]
                discordant, total = int(np.count_nonzero(a ^ b)), len(iu[0])
            tmp = discordant / total / self.bandwidth
            self.w[_id] = 0.75 * (1 - tmp * tmp) if tmp <= 1 else 0
        if self.only_source:
            self.w[-1] = 0.
            self.w = np.array(self.w) / np.sum(self.w)
        w = self.w.copy()
        self.target_weight.append(w[-1] / sum(w))
        self.hist_ws.append(w)
        self.iteration_id += 1

    def predict(self, X):
        """tst.py:64-73."""
        mu, var = self.target_surrogate.predict(X)
        denominator = 0.75
        for i in range(0, self.K):
            weight = self.w[i]
            mu_t, _ = self.source_surrogates[i].predict(X)
            mu += weight * mu_t
            denominator += weight
        mu /= denominator
        return mu, var


class OraclePOGPE(OracleFacade):
    """tlbo/facade/pogpe.py (the "gogpe" baseline)."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, only_source=False, literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, literal_loops)
        self.method_id = 'pogpe'
        self.only_source = only_source
        self.build_source_surrogates(normalize='scale')
        self.w = np.array([0.5 / self.K] * self.K + [0.5])
        self.hist_ws = list()
        self.iteration_id = 0

    def train(self, X, y):
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='scale')
        if self.only_source:
            self.w[-1] = 0.
            self.w = self.w / np.sum(self.w)
        self.hist_ws.append(self.w.copy())
        self.iteration_id += 1

    def predict(self, X):
        """pogpe.py:36-60."""
        w = self.w.copy()
        n, m = X.shape[0], len(self.w)
        var_buf = np.zeros((n, m))
        mu_buf = np.zeros((n, m))
        for i in range(0, self.K + 1):
            if i == self.K:
                if self.target_surrogate is not None:
                    _mu, _var = self.target_surrogate.predict(X)
                else:
                    _mu, _var = 0., 0.
            else:
                _mu, _var = self.source_surrogates[i].predict(X)
            _mu, _var = _mu.flatten(), _var.flatten()
            if (_var != 0).all():
                var_buf[:, i] = (1. / _var * w[i])
                mu_buf[:, i] = (1. / _var * _mu * w[i])
        tmp = np.sum(var_buf, axis=1)
        tmp[tmp == 0.] = 1e-5
        var = 1. / tmp
        mu = np.sum(mu_buf, axis=1) * var
        return mu.reshape(-1, 1), var.reshape(-1, 1)


# ---- meta-feature baselines ------------------------------------------------------
def scale_fit_meta_features(meta_features):
    """base_facade.py:193-202: mean imputation, then min-max scaling fitted on the SOURCE tasks."""
    from sklearn.impute import SimpleImputer
    from sklearn.preprocessing import MinMaxScaler
    meta_features = np.array(meta_features)
    imputer = SimpleImputer(missing_values=np.nan, strategy='mean')
    imputer.fit(meta_features)
    meta_features = imputer.transform(meta_features)
    scaler = MinMaxScaler()
    scaler.fit(meta_features)
    return scaler.transform(meta_features), imputer, scaler


def scale_transform_meta_features(meta_feature, imputer, scaler):
    """base_facade.py:204-208 (the target's vector, clipped into [0, 1])."""
    v = scaler.transform(imputer.transform(np.array([meta_feature])))
    return np.clip(v, 0, 1)


def transform_pairwise(X, y):
    """tlbo/utils/rank_svm.py:7-43, literally: all within-group pairs with different targets, labels balanced by
    the parity of the COMBINATION index k (skipped combinations count)."""
    X_new, y_new = [], []
    y = np.asarray(y)
    if y.ndim == 1:
        y = np.c_[y, np.ones(y.shape[0])]
    for k, (i, j) in enumerate(itertools.combinations(range(X.shape[0]), 2)):
        if y[i, 0] == y[j, 0] or y[i, 1] != y[j, 1]:
            continue
        X_new.append(X[i] - X[j])
        y_new.append(np.sign(y[i, 0] - y[j, 0]))
        if y_new[-1] != (-1) ** k:
            y_new[-1] = -y_new[-1]
            X_new[-1] = -X_new[-1]
    return np.asarray(X_new), np.asarray(y_new).ravel()


class OracleRankSVM(object):
    """tlbo/utils/rank_svm.py:46-81 (sklearn ``SVC(kernel='linear', C=.1)`` as the reference)."""

    def __init__(self):
        from sklearn import svm
        self.clf = svm.SVC(kernel='linear', C=.1)
        self.coef = None

    def fit(self, X, y):
        X_trans, y_trans = transform_pairwise(X, y)
        self.clf.fit(X_trans, y_trans)
        self.coef = self.clf.coef_.ravel() / np.linalg.norm(self.clf.coef_)

    def predict(self, X):
        return np.dot(X, self.coef)


class OracleSCoT(OracleFacade):
    """tlbo/facade/scot.py: RankSVM scores over the stacked source + target rows (meta-feature columns appended),
    then ONE surrogate on all K x 50 + n rows; ``predict`` appends the target's meta-features."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, metafeatures=None, **kw):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, **kw)
        self.method_id = 'scot'
        if metafeatures is None:
            raise ValueError('SCoT needs meta-features about the datasets.')
        assert len(metafeatures) == self.K + 1
        self.metafeatures = np.zeros(shape=(len(metafeatures), len(metafeatures[0])))
        self.metafeatures[:-1], imp, sc = scale_fit_meta_features(metafeatures[:-1])
        self.metafeatures[-1] = scale_transform_meta_features(metafeatures[-1], imp, sc)
        self.space_types = self.types
        self.types = np.concatenate([self.types, np.zeros(self.metafeatures.shape[1], dtype=self.types.dtype)])
        self.X, self.y = None, None
        for i, (Xs, ys) in enumerate(self.source_hpo_data):
            X = np.asarray(Xs, dtype=np.float64)[:self.num_src_hpo_trial]
            y = np.array(ys, dtype=np.float64)[:self.num_src_hpo_trial]
            X = np.c_[X, np.array([list(self.metafeatures[i]) for _ in range(len(y))])]
            y = np.c_[y, np.array([i] * X.shape[0])]
            if self.X is not None:
                self.X, self.y = np.r_[self.X, X], np.r_[self.y, y]
            else:
                self.X, self.y = X, y
        self.svm_coefs = []

    def train(self, X, y):
        num_sample = y.shape[0]
        meta_vec = np.array([list(self.metafeatures[self.K]) for _ in range(num_sample)])
        _X = np.r_[self.X, np.c_[X, meta_vec]]
        _y = np.r_[self.y, np.c_[y, np.array([self.K] * num_sample)]]
        rank_svm = OracleRankSVM()
        rank_svm.fit(_X, _y)
        self.svm_coefs.append(rank_svm.coef.copy())
        pred_y = rank_svm.predict(_X)
        self.target_surrogate = self.build_single_surrogate(_X, pred_y, normalize='none')

    def predict(self, X):
        meta_vec = np.array([list(self.metafeatures[self.K]) for _ in range(X.shape[0])])
        return self.target_surrogate.predict(np.c_[X, meta_vec])

    def predict_marginalized_over_instances(self, X):
        if X.shape[1] != len(self.space_types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.space_types), X.shape[1]))
        mean, var = self.predict(X)
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var


class OracleTSTM(OracleFacade):
    """tlbo/facade/tstm.py: TST with Epanechnikov weights from the meta-feature distance (bandwidth 1.8);
    the reference never appends to ``target_weight`` here."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, metafeatures=None, **kw):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, **kw)
        self.method_id = 'tst'
        self.build_source_surrogates(normalize='scale')
        self.w = [0.75] * (self.K + 1)
        self.metafeatures = np.zeros(shape=(len(metafeatures), len(metafeatures[0])))
        self.metafeatures[:-1], imp, sc = scale_fit_meta_features(metafeatures[:-1])
        self.metafeatures[-1] = scale_transform_meta_features(metafeatures[-1], imp, sc)
        self.bandwidth = 1.8
        assert len(metafeatures) == self.K + 1
        self.meta_dist = [float(np.sqrt(np.sum(np.square(self.metafeatures[self.K] - self.metafeatures[_id]))))
                          for _id in range(self.K)]
        self.hist_ws = list()
        self.iteration_id = 0

    def train(self, X, y):
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='scale')
        for _id in range(self.K):
            tmp = self.meta_dist[_id] / self.bandwidth
            self.w[_id] = 0.75 * (1 - tmp * tmp) if tmp <= 1 else 0
        self.hist_ws.append(self.w.copy())
        self.iteration_id += 1

    def predict(self, X):
        """tstm.py:49-58."""
        mu, var = self.target_surrogate.predict(X)
        denominator = 0.75
        for i in range(0, self.K):
            weight = self.w[i]
            mu_t, _ = self.source_surrogates[i].predict(X)
            mu += weight * mu_t
            denominator += weight
        mu /= denominator
        return mu, var


class OracleSGPR(OracleFacade):
    """tlbo/facade/stacking_gpr.py:7-128 ``SGPR``: a stack of residual GPs.  For every source history in turn (and
    finally for the target observations) ONE freshly built model is trained twice -- on ``y``, then on ``y`` minus
    the prior mean cached at those configurations (the second fit starts from the first fit's optimum and the
    model's continuing random streams) -- and its posterior over the union of all configurations updates the cached
    prior: ``mu += mu_top``, ``sigma = sigma_top ** beta * sigma ** (1 - beta)`` with
    ``beta = alpha n / (alpha n + prior_size)`` (``sigma_top`` is the VARIANCE ``model.predict`` returns).
    ``predict`` looks configurations up in that cache.

    ``target_hp_configs``: float64[N, d], the target's candidate table."""

    def __init__(self, types, source_hpo_data, seed, target_hp_configs, num_src_hpo_trial=50, literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, literal_loops)
        self.method_id = 'sgpr'
        self.target_hp_configs = np.asarray(target_hp_configs, dtype=np.float64)
        self.alpha = 0.95
        self.iteration_id = 0
        self.get_regressor()

    @staticmethod
    def _key(row):
        return str(list(row))                                   # stacking_gpr.py:41-42,77

    def get_regressor(self):
        """stacking_gpr.py:25-63."""
        configs, self.index_mapper = [], dict()
        rows = [np.asarray(X, dtype=np.float64)[:self.num_src_hpo_trial] for X, _ in self.source_hpo_data]
        for X in rows + [self.target_hp_configs]:
            for x in X:
                k = self._key(x)
                if k not in self.index_mapper:                  # `_config not in configs_list`
                    self.index_mapper[k] = len(configs)
                    configs.append(np.array(x))
        self.configs_X = np.array(configs)
        n = len(configs)
        self.cached_prior_mu, self.cached_prior_sigma = np.zeros(n), np.ones(n)
        self.cached_stacking_mu, self.cached_stacking_sigma = np.zeros(n), np.ones(n)
        self.prior_size = 0
        for X, y in self.source_hpo_data:
            X = np.asarray(X, dtype=np.float64)[:self.num_src_hpo_trial]
            y = np.array(y, dtype=np.float64)[:self.num_src_hpo_trial]
            self.train_regressor(X, y)
            self.prior_size = len(y)

    def train_regressor(self, X, y, is_top=False):
        """stacking_gpr.py:65-99."""
        model = self._new_model()
        model.train(X, y)
        idxs = [self.index_mapper[self._key(x)] for x in X]
        prior_mu = np.array(self.cached_prior_mu[idxs])
        model.train(X, y - prior_mu)
        top_size = len(y)
        beta = self.alpha * top_size / (self.alpha * top_size + self.prior_size)
        mu_top, sigma_top = model.predict(self.configs_X)
        mu_top, sigma_top = mu_top.flatten(), sigma_top.flatten()
        if is_top:
            self.cached_stacking_mu = self.cached_prior_mu + mu_top
            self.cached_stacking_sigma = np.power(sigma_top, beta) * np.power(self.cached_prior_sigma, 1 - beta)
        else:
            self.cached_prior_mu += mu_top.flatten()
            self.cached_prior_sigma = np.power(sigma_top, beta) * np.power(self.cached_prior_sigma, 1 - beta)
        self.last_model = model

    def train(self, X, y):
        """stacking_gpr.py:101-114."""
        if any(self._key(x) not in self.index_mapper for x in X):
            self.get_regressor()
        self.train_regressor(X, y, is_top=True)
        self.iteration_id += 1

    def predict(self, X):
        """stacking_gpr.py:116-124."""
        idx = [self.index_mapper[self._key(x)] for x in X]
        return (np.array(self.cached_stacking_mu[idx]).reshape(-1, 1),
                np.array(self.cached_stacking_sigma[idx]).reshape(-1, 1))


class OracleRandomSearch(OracleFacade):
    """tlbo/facade/random_surrogate.py:5-19: ``predict`` is ``rng.rand(n, 1)`` with variance 1e-5."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, literal_loops)
        self.method_id = 'rs'
        self.rng = np.random.RandomState(seed)

    def train(self, X, y):
        pass

    def predict(self, X):
        n = X.shape[0]
        return self.rng.rand(n, 1), np.array([1e-5] * n).reshape(-1, 1)


class OracleES(OracleOBTL):
    """tlbo/facade/obtl_es.py (``--methods es``): source weights by greedy ensemble selection over the soft ranking
    loss (:61-87), target weight by sampling, rho = 0.6; ``predict`` accumulates the target first and then the
    sources with a positive weight (:190-207)."""

    def __init__(self, types, source_hpo_data, seed, num_src_hpo_trial=50, fusion_method='idp_lc', literal_loops=False):
        super().__init__(types, source_hpo_data, seed, num_src_hpo_trial, fusion_method, literal_loops)
        self.method_id = 'es'
        self.last_selection = None

    def train(self, X, y):
        """obtl_es.py:54-106."""
        instance_num = X.shape[0]
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='standardize')
        self.target_y_range = 0.5 * (np.max(y) - np.min(y))
        base_predictions, surrogate_ensemble, surrogate_idx = [], [], []
        for i in range(self.K):
            mean, _ = self.source_surrogates[i].predict(X)
            base_predictions.append(mean.flatten())
        for _ in range(self.ensemble_size):
            loss_list = []
            for i in range(self.K):
                if len(surrogate_ensemble) == 0:
                    _predictions = base_predictions[i]
                else:
                    _predictions = np.mean(surrogate_ensemble + [base_predictions[i]], axis=0)
                loss_list.append(obtl_ranking_loss(_predictions, y, self.literal_loops))
            argmin_idx = np.argmin(loss_list)
            surrogate_ensemble.append(base_predictions[argmin_idx])
            surrogate_idx.append(int(argmin_idx))
        self.last_selection = surrogate_idx
        w_source = np.zeros(self.K)
        for idx in surrogate_idx:
            w_source[idx] += 1
        w_source = w_source / np.sum(w_source)
        w_target = 0.
        if instance_num >= self.min_num_y:
            w_target = self.calculate_weight_by_sampling(X, y)
            if instance_num >= 2 * self.min_num_y:
                w_target = np.max([w_target, self.w[-1]])
            w_source *= (1 - w_target)
        w_new = np.asarray(list(w_source) + [w_target])
        rho = 0.6
        self.w = rho * w_new + (1 - rho) * self.w
        self.hist_ws.append(self.w.copy())
        self.iteration_id += 1

    def predict(self, X):
        w = self.w.copy()
        if self.target_surrogate is not None:
            mu, var = self.target_surrogate.predict(X)
        else:
            mu, var = np.zeros((X.shape[0], 1)), np.zeros((X.shape[0], 1))
        mu *= w[-1]
        var *= (w[-1] * w[-1])
        for i in range(0, self.K):
            if w[i] > 0:
                mu_t, var_t = self.source_surrogates[i].predict(X)
                mu += (w[i] * mu_t)
                var += (w[i] * w[i] * var_t)
        return mu, var
