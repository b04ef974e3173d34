"""Random-search baseline behind the facade interface (reference tlbo/facade/random_surrogate.py:5-19,
``--methods rs`` of tools/offline_benchmark.py): ``predict`` returns ``rng.rand(n, 1)`` as the mean and 1e-5 as
the variance, so EI + ordering pick a (nearly) random unevaluated row.

Device path: the facade's legacy MT19937 state lives on the device (``ops.DeviceRandomState``, bit-identical to
numpy's ``random_sample``), the "mean" is generated there and the fused kernel runs in all-resident mode over it --
the candidate table never produces a host array."""
import numpy as np

from tlbo_b200.facade.base_facade import BaseFacade


class RandomSearch(BaseFacade):
    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, fusion_method='idp', **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        self.method_id = 'rs'
        self.rng = np.random.RandomState(seed)
        self.hist_ws = list()
        self._dev_rng = None
        self._one = None

    def train(self, X, y):
        pass

    def _rand_dev(self, n):
        """``self.rng.rand(n)`` generated on the device; the host generator is kept in step."""
        import torch
        from tlbo_b200 import ops
        if self._dev_rng is None:
            self._dev_rng = ops.DeviceRandomState(self.rng)
        else:
            self._dev_rng.from_host(self.rng)
        out = torch.empty(int(n), dtype=torch.float64, device='cuda')
        self._dev_rng.rand_into(out)
        self._dev_rng.to_host(self.rng)
        return out

    def predict(self, X):
        n = X.shape[0]
        return self.rng.rand(n, 1), np.array([1e-5] * n).reshape(-1, 1)

    def predict_marginalized_over_instances(self, X):
        mean, var = self.predict(X)
        var[var < self.var_threshold] = self.var_threshold
        var[np.isnan(var)] = self.var_threshold
        return mean, var

    def register_static_pool(self, X_dev):
        pass

    def _fused(self, X_dev, eta=0.0, par=0.0, var_floor=0.0, keys=None, excluded=None, idx_offset=0,
               want_rows=False, ws=None, sync=True):
        import torch
        from tlbo_b200 import ops
        N = int(X_dev.shape[0])
        mu = self._rand_dev(N).reshape(1, N)
        var = torch.full((1, N), 1e-5, dtype=torch.float64, device=X_dev.device)
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
