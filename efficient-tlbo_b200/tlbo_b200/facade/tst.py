"""TST-R baseline (reference tlbo/facade/tst.py, ``use_metafeatures=False``): Epanechnikov
kernel weights from the fraction of discordant pairs between the target observations and
each source surrogate's predictions -- the pair count is the integer kernel, bit-exact."""
import numpy as np

from tlbo_b200.facade.base_facade import BaseFacade


class TST(BaseFacade):
    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, use_metafeatures=False, metafeatures=None, only_source=False, **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        if use_metafeatures:
            raise ValueError('use TSTM for the meta-feature variant (reference tlbo/facade/tstm.py)')
        self.method_id = 'tst'
        self.only_source = only_source
        self.build_source_surrogates(normalize='scale')
        self.w = [0.75] * (self.K + 1)
        self.bandwidth = 0.45
        self.hist_ws = list()
        self.iteration_id = 0
        self._fused_mode = 1

    def train(self, X, y):
        from tlbo_b200 import ops
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='scale')
        n_sample = X.shape[0]
        mu, _ = self._source_predictions(X)
        counts = ops.rank_loss_counts(np.asarray(y, dtype=np.float64), mu[None, :self.K, :], [0] * self.K,
                                      [n_sample] * self.K, upper_only=True)[0]
        total_pairs = n_sample * (n_sample - 1) // 2
        for _id in range(self.K):
            tmp = int(counts[_id]) / total_pairs / self.bandwidth
            self.w[_id] = 0.75 * (1 - tmp * tmp) if tmp <= 1 else 0
        if self.only_source:
            self.w[-1] = 0.
            self.w = np.array(self.w) / np.sum(self.w)
        w = self.w.copy()
        self.target_weight.append(w[-1] / sum(w))
        self.hist_ws.append(w)
        self.iteration_id += 1

    def _ensemble_args(self):
        # tst.py:64-73: target first (its weight is the fixed 0.75 in the denominator)
        order = np.array([self.K] + list(range(self.K)), dtype=np.int32)
        return np.asarray(self.w, dtype=np.float64), None, order

    def predict(self, X):
        return self._predict_rows(X, 0.0)


class TSTM(BaseFacade):
    """TST-M (reference tlbo/facade/tstm.py:5-58): the Epanechnikov weights come from the Euclidean distance
    between the scaled data-set meta-features of the target and each source task (bandwidth 1.8) -- no pair
    counts; the reference never appends to ``target_weight`` here.  Posterior: TST's (fused mode 1)."""

    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, metafeatures=None, **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        self.method_id = 'tst'
        self.build_source_surrogates(normalize='scale')
        # Weights for base surrogates and the target surrogate.
        self.w = [0.75] * (self.K + 1)
        self.metafeatures = np.zeros(shape=(len(metafeatures), len(metafeatures[0])))
        self.metafeatures[:-1] = self.scale_fit_meta_features(metafeatures[:-1])
        self.metafeatures[-1] = self.scale_transform_meta_features(metafeatures[-1])
        self.bandwidth = 1.8
        assert len(metafeatures) == self.K + 1
        self.meta_dist = [0.] * self.K
        for _id in range(self.K):
            _dist = np.sqrt(np.sum(np.square(self.metafeatures[self.K] - self.metafeatures[_id])))
            self.meta_dist[_id] = _dist
        self.hist_ws = list()
        self.iteration_id = 0
        self._fused_mode = 1

    def train(self, X, y):
        self.target_surrogate = self.build_single_surrogate(X, y, normalize='scale')
        for _id in range(self.K):
            tmp = self.meta_dist[_id] / self.bandwidth
            self.w[_id] = 0.75 * (1 - tmp * tmp) if tmp <= 1 else 0
        self.hist_ws.append(self.w.copy())
        self.iteration_id += 1

    def _ensemble_args(self):
        order = np.array([self.K] + list(range(self.K)), dtype=np.int32)
        return np.asarray(self.w, dtype=np.float64), None, order

    def predict(self, X):
        return self._predict_rows(X, 0.0)
