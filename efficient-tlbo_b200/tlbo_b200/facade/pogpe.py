"""POGPE -- the "gogpe" baseline of BASELINE configs[4] (reference tlbo/facade/pogpe.py:5-60):
a generalised product of GP experts with fixed weights (0.5 / K per source, 0.5 for the target);
the precision-weighted combination runs in the fused predict kernel (mode 2)."""
import numpy as np

from tlbo_b200.facade.base_facade import BaseFacade


class POGPE(BaseFacade):
    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, only_source=False, **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        self.method_id = 'pogpe'
        self.only_source = only_source
        self.build_source_surrogates(normalize='scale')
        self.w = np.array([0.5 / self.K] * self.K + [0.5])
        self.hist_ws = list()
        self.iteration_id = 0
        self._fused_mode = 2

    def train(self, X, y):
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='scale')
        if self.only_source:
            self.w[-1] = 0.
            self.w = self.w / np.sum(self.w)
        w = self.w.copy()
        self.hist_ws.append(w)
        self.iteration_id += 1

    def predict(self, X):
        # pogpe.py:36-60; surrogate variances are >= 1e-10 * std_y^2 > 0 (gp.py:297-300), so the
        # reference's per-surrogate `(_var != 0).all()` guard always passes
        return self._predict_rows(X, 0.0)
