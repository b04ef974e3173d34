"""Stacking-GP baseline on the device path (reference tlbo/facade/stacking_gpr.py:7-128 ``SGPR``).

A stack of residual GPs: for every source history in turn, and finally for the target observations, ONE freshly
built surrogate is trained twice -- on ``y``, then on ``y`` minus the prior mean cached at those configurations
(stacking_gpr.py:66-81: the second fit starts from the first fit's optimum and continues the model's random
streams) -- and its posterior over the union of all configurations updates the cached prior
(``mu += mu_top``; ``sigma = sigma_top ** beta * sigma ** (1 - beta)``, ``beta = alpha n / (alpha n + prior_size)``,
``sigma_top`` being the VARIANCE ``model.predict`` returns, :83-99).  ``predict`` is a lookup in that cache.

Device mapping: the union of configurations (K x 50 source rows + the candidate table, de-duplicated) lives in HBM
next to the four cached vectors; a regressor is two launches of the batched fit kernel (one CTA per restart), one
launch of the pool-posterior kernel over the union and three element-wise updates; the acquisition call is the fused
kernel in all-resident mode (combine + EI + ordering + arg-max over the cached rows of the candidate table) -- no
per-row Python lookups (the reference builds ``str(list(x))`` keys for every candidate on every call)."""
import numpy as np

from tlbo_b200.facade.base_facade import BaseFacade, _history_to_arrays


class SGPR(BaseFacade):
    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        if self._G > 1:
            raise ValueError("SGPR on the device path takes surrogate_type='gp'")
        if getattr(self.types, 'has_conditions', False):
            raise ValueError('SGPR over a space with conditions is not supported on the device path')
        self.method_id = 'sgpr'
        self.alpha = 0.95
        self.prior_size = 0
        self.iteration_id = 0
        self.hist_ws = list()
        self._sources = [_history_to_arrays(h) for h in source_hpo_data]
        from tlbo_b200.config_space import convert_configurations_to_array
        self._target_X = np.asarray(convert_configurations_to_array(target_hp_configs), dtype=np.float64)
        self._pool_rows = {}          # (data_ptr, N) of a registered candidate table -> device row indices
        self._one = None              # one placeholder slot for the fused kernel's all-resident mode
        self._host_cache = None       # host copies of the stacked mean / sigma (predict() lookups)
        self.get_regressor()

    # ---- the configuration set ---------------------------------------------------------------------------
    @staticmethod
    def _key(row):
        return np.ascontiguousarray(row, dtype=np.float64).tobytes()

    def get_regressor(self):
        """stacking_gpr.py:25-63: the de-duplicated union of the first ``num_src_hpo_trial`` configurations of every
        source and the candidate table (first occurrence keeps the index), then one regressor per source."""
        import torch
        from tlbo_b200 import ops
        configs, self.index_mapper = [], dict()
        for X in [s[0][:self.num_src_hpo_trial] for s in self._sources] + [self._target_X]:
            for x in X:
                k = self._key(x)
                if k not in self.index_mapper:
                    self.index_mapper[k] = len(configs)
                    configs.append(x)
        self.configs_X = np.array(configs, dtype=np.float64).reshape(len(configs), len(self.types))
        n = len(configs)
        Xd = self.configs_X.copy()
        Xd[~np.isfinite(Xd)] = -1
        self._configs_dev = ops.h2d(Xd)
        dev = self._configs_dev.device
        self._prior_mu = torch.zeros(n, dtype=torch.float64, device=dev)
        self._prior_sigma = torch.ones(n, dtype=torch.float64, device=dev)
        self._stack_mu = torch.zeros(n, dtype=torch.float64, device=dev)
        self._stack_sigma = torch.ones(n, dtype=torch.float64, device=dev)
        self._target_idx = np.array([self.index_mapper[self._key(x)] for x in self._target_X], dtype=np.int64)
        self._target_idx_dev = ops.h2d(self._target_idx)
        self._pool_rows = {k: self._target_idx_dev for k in self._pool_rows}
        self._host_cache = None
        self.prior_size = 0
        for X, y in self._sources:
            X = np.asarray(X, dtype=np.float64)[:self.num_src_hpo_trial]
            y = np.array(y, dtype=np.float64)[:self.num_src_hpo_trial]
            self.train_regressor(X, y)
            self.prior_size = len(y)

    def _indices(self, X):
        try:
            return np.array([self.index_mapper[self._key(x)] for x in X], dtype=np.int64)
        except KeyError:
            raise KeyError('configuration outside the stacking set (stacking_gpr.py:120 raises the same)')

    # ---- one residual regressor -----------------------------------------------------------------------------
    def train_regressor(self, X, y, is_top=False):
        """stacking_gpr.py:65-99."""
        import torch
        from tlbo_b200 import ops
        model = self._new_model()
        pack = self._ensure_pack(X.shape[0])
        model.bind(pack, self.K)
        model.train(X, y)
        idxs = self._indices(X)
        prior_mu = ops.d2h(self._prior_mu[ops.h2d(idxs)])
        model.train(X, y - prior_mu)
        top_size = len(y)
        beta = self.alpha * top_size / (self.alpha * top_size + self.prior_size)
        mu_top, sigma_top = ops.predict_pool_posteriors(model._view(), self.types, self._configs_dev,
                                                        nmax=model.n_train_)
        mu_top, sigma_top = mu_top[0], sigma_top[0]
        sigma = torch.pow(sigma_top, beta) * torch.pow(self._prior_sigma, 1 - beta)
        if is_top:
            self._stack_mu = self._prior_mu + mu_top
            self._stack_sigma = sigma
        else:
            self._prior_mu = self._prior_mu + mu_top
            self._prior_sigma = sigma
        self._host_cache = None
        self.target_surrogate = model
        return model

    def train(self, X, y):
        """stacking_gpr.py:101-114."""
        X = np.asarray(X, dtype=np.float64)
        if any(self._key(x) not in self.index_mapper for x in X):
            self.get_regressor()
        self.train_regressor(X, np.asarray(y, dtype=np.float64), is_top=True)
        self.iteration_id += 1

    # ---- reference-visible state ----------------------------------------------------------------------------
    def _host(self):
        if self._host_cache is None:
            from tlbo_b200 import ops
            self._host_cache = tuple(ops.d2h(t) for t in (self._prior_mu, self._prior_sigma, self._stack_mu,
                                                           self._stack_sigma))
        return self._host_cache

    cached_prior_mu = property(lambda self: self._host()[0])
    cached_prior_sigma = property(lambda self: self._host()[1])
    cached_stacking_mu = property(lambda self: self._host()[2])
    cached_stacking_sigma = property(lambda self: self._host()[3])

    def predict(self, X):
        """stacking_gpr.py:116-124: lookup of the cached stacked mean / sigma."""
        idx = self._indices(np.asarray(X, dtype=np.float64))
        _, _, mu, sigma = self._host()
        return mu[idx].reshape(-1, 1), sigma[idx].reshape(-1, 1)

    def predict_marginalized_over_instances(self, X):
        """base_facade.py:110-143: ``predict`` with the 1e-10 variance floor."""
        mean, var = self.predict(X)
        mean, var = mean.copy(), var.copy()
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var

    # ---- acquisition over the candidate table ------------------------------------------------------------------
    def register_static_pool(self, X_dev):
        """``X_dev`` holds the candidate table the facade was constructed with (``OfflineSearch`` uploads
        ``target_hp_configs`` once): its rows map to fixed positions of the stacking set."""
        if int(X_dev.shape[0]) > len(self._target_idx):
            raise ValueError('candidate pool has more rows than target_hp_configs')
        self._pool_rows[(X_dev.data_ptr(), int(X_dev.shape[0]))] = self._target_idx_dev
        self._pool_keepalive = X_dev

    def _fused(self, X_dev, eta=0.0, par=0.0, var_floor=0.0, keys=None, excluded=None, idx_offset=0,
               want_rows=False, ws=None, sync=True):
        import torch
        from tlbo_b200 import ops
        N = int(X_dev.shape[0])
        rows = self._pool_rows.get((X_dev.data_ptr(), N))
        if rows is not None:
            rows = rows[int(idx_offset):int(idx_offset) + N]           # a row-sharded pool passes its offset
        else:
            rows = ops.h2d(self._indices(ops.d2h(X_dev)))
        mu = self._stack_mu[rows].reshape(1, N)
        var = self._stack_sigma[rows].reshape(1, N)
        if self._one is None:
            self._one = ops.SurrogatePack(1, 8, len(self.types))
            self._one.n.fill_(1)
            self._w_one = torch.ones(1, dtype=torch.float64, device=X_dev.device)
        return ops.predict_ei_fused(self._one, self.types, self._w_one, None, X_dev, eta, par=par, var_floor=var_floor,
                                    keys=keys, excluded=excluded, idx_offset=idx_offset, mode=0, want_rows=want_rows,
                                    ws=ws, sync=sync, nmax=8, cache=(mu, var))

    def acq_small(self, X, eta, par, var_floor):
        return None

    def climb(self, *a, **k):
        return None

    def _ensemble_args(self):
        return np.ones(1), None, None
