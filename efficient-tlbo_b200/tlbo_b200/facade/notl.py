"""No-transfer baseline: the target GP alone (reference tlbo/facade/notl.py)."""
import numpy as np

from tlbo_b200.facade.base_facade import BaseFacade


class NoTL(BaseFacade):
    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        self.method_id = 'notl'
        self.hist_ws = list()

    def train(self, X, y):
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='standardize')

    def _ensemble_args(self):
        w = np.zeros(self.K + 1)
        w[-1] = 1.0
        skip = np.array([1] * self.K + [0], dtype=np.int32)
        return w, skip, None

    def predict(self, X):
        return self._predict_rows(X, 0.0)
