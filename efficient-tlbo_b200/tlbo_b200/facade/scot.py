"""SCoT baseline (reference tlbo/facade/scot.py:11-102; Bardenet et al. 2013): a RankSVM over all within-task
pairs of the stacked source + target observations turns the K + 1 response scales into ONE ranking score, and a
single surrogate is fitted on all ``K * num_src_hpo_trial + n`` rows, the scaled data-set meta-features
appended as extra (continuous) columns; prediction appends the target task's meta-features.

The GP on the stacked rows (n ~ 1.5 k at 29 sources) runs on the same device fit / posterior kernels as the
other facades through an inner single-surrogate facade over the EXTENDED space (configuration columns +
meta-feature columns): the candidate pool is extended once on the device (the target's meta-feature
columns are constants), so the per-iteration launch is the usual fused posterior + EI + arg-max."""
import numpy as np

from tlbo_b200.config_space import TableSpace
from tlbo_b200.facade.base_facade import BaseFacade, _history_to_arrays
from tlbo_b200.facade.notl import NoTL
from tlbo_b200.utils.rank_svm import RankSVM


class SCoT(BaseFacade):
    def __init__(self, config_space, source_hpo_data, target_hp_configs, seed, surrogate_type='gp',
                 num_src_hpo_trial=50, metafeatures=None, **kw):
        super().__init__(config_space, source_hpo_data, seed, target_hp_configs, surrogate_type=surrogate_type,
                         num_src_hpo_trial=num_src_hpo_trial, **kw)
        self.method_id = 'scot'
        if metafeatures is None:
            raise ValueError('SCoT needs meta-features about the datasets.')
        # Should include the meta-feature for target problem.
        assert len(metafeatures) == (self.K + 1)
        self.metafeatures = np.zeros(shape=(len(metafeatures), len(metafeatures[0])))
        self.metafeatures[:-1] = self.scale_fit_meta_features(metafeatures[:-1])
        self.metafeatures[-1] = self.scale_transform_meta_features(metafeatures[-1])

        self.X, self.y = None, None
        for i, hist in enumerate(self.source_hpo_data):
            X, y = _history_to_arrays(hist)
            X = X[:self.num_src_hpo_trial]
            y = y[:self.num_src_hpo_trial]
            meta_vec = np.array([list(self.metafeatures[i]) for _ in range(len(y))])
            X = np.c_[X, meta_vec]
            y = np.c_[y, np.array([i] * X.shape[0])]
            if self.X is not None:
                self.X = np.r_[self.X, X]
                self.y = np.r_[self.y, y]
            else:
                self.X, self.y = X, y
        self.iteration_id = 0
        self.svm_coefs = []
        # the stacked surrogate lives in an inner facade over the extended space (scot.py:75-88 builds the same
        # space: the configuration's hyper-parameters plus one UniformFloat in [0, 1] per meta-feature)
        m = self.metafeatures.shape[1]
        ext_types = np.concatenate([np.asarray(self.types), np.zeros(m, dtype=np.asarray(self.types).dtype)])
        rows = (0 if self.X is None else self.X.shape[0]) + 96
        self._inner = NoTL(TableSpace(ext_types), [], None, seed, surrogate_type=surrogate_type,
                           num_src_hpo_trial=num_src_hpo_trial, max_target_rows=rows,
                           **{k: v for k, v in kw.items() if k == 'mcmc_steps'})
        self._ext_pools = {}

    # ---- training ---------------------------------------------------------------------------
    def train(self, X, y):
        num_sample = y.shape[0]
        meta_vec = np.array([list(self.metafeatures[self.K]) for _ in range(num_sample)])
        _X = np.c_[X, meta_vec]
        _y = np.c_[y, np.array([self.K] * num_sample)]
        if self.X is not None:
            _X = np.r_[self.X, _X]
            _y = np.r_[self.y, _y]
        rank_svm = RankSVM()
        rank_svm.fit(_X, _y)
        self.svm_coefs.append(rank_svm.coef.copy())
        pred_y = rank_svm.predict(_X)
        self._inner._overlap_hook = self._overlap_hook
        try:
            self.target_surrogate = self._inner.build_single_surrogate(_X, pred_y, normalize='none')
        finally:
            self._inner._overlap_hook = None
        self._inner.target_surrogate = self.target_surrogate
        self.iteration_id += 1

    # ---- prediction -------------------------------------------------------------------------
    def _extend_host(self, X):
        X = np.asarray(X, dtype=np.float64)
        meta_vec = np.array([list(self.metafeatures[self.K]) for _ in range(X.shape[0])])
        return np.c_[X, meta_vec]

    def _extend_device(self, X_dev):
        """``X_dev [N, d]`` with the target's meta-feature columns appended; the extension of a registered
        (immutable) pool is built once and kept."""
        import torch
        key = (X_dev.data_ptr(), int(X_dev.shape[0]))
        hit = self._ext_pools.get(key)
        if hit is not None:
            return hit[1]
        meta = torch.from_numpy(np.ascontiguousarray(self.metafeatures[self.K])).to(X_dev.device)
        return torch.cat([X_dev, meta[None, :].expand(X_dev.shape[0], -1)], dim=1).contiguous()

    def register_static_pool(self, X_dev):
        key = (X_dev.data_ptr(), int(X_dev.shape[0]))
        if key not in self._ext_pools:
            self._ext_pools = {key: (X_dev, self._extend_device(X_dev))}      # holds X_dev: the address stays taken

    def _fused(self, X_dev, **kw):
        if self.target_surrogate is None:
            raise ValueError('Target surrogate is none.')
        return self._inner._fused(self._extend_device(X_dev), **kw)

    def acq_small(self, X, eta, par, var_floor):
        return self._inner.acq_small(self._extend_host(X), eta, par, var_floor)

    def predict(self, X):
        if X.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X.shape[1]))
        return self._inner.predict(self._extend_host(X))

    def predict_marginalized_over_instances(self, X):
        if len(X.shape) != 2:
            raise ValueError('Expected 2d array, got %dd array!' % len(X.shape))
        if X.shape[1] != len(self.types):
            raise ValueError('Rows in X should have %d entries but have %d!' % (len(self.types), X.shape[1]))
        return self._inner.predict_marginalized_over_instances(self._extend_host(X))
