"""Expected improvement with the reference's call interface
(reference tlbo/acquisition_function/acquisition.py:13-77 ``AbstractAcquisitionFunction``,
:112-177 ``EI``).  ``_compute`` is one launch of the fused ensemble predict + EI kernel;
``best_unevaluated`` additionally folds the candidate ordering of
tlbo/optimizer/ei_optimization.py:106-132 and the first-unevaluated scan of
tlbo/framework/smbo_offline.py:261-263 into the same launch."""
import abc
import logging

import numpy as np

from tlbo_b200.config_space import convert_configurations_to_array


class AbstractAcquisitionFunction(object, metaclass=abc.ABCMeta):
    def __str__(self):
        return type(self).__name__ + " (" + self.long_name + ")"

    def __init__(self, model, **kwargs):
        self.model = model
        self.logger = logging.getLogger(self.__module__ + "." + self.__class__.__name__)

    def update(self, **kwargs):
        for key in kwargs:
            setattr(self, key, kwargs[key])

    def __call__(self, configurations):
        X = convert_configurations_to_array(configurations)
        if len(X.shape) == 1:
            X = X[np.newaxis, :]
        acq = self._compute(X)
        if np.any(np.isnan(acq)):
            idx = np.where(np.isnan(acq))[0]
            acq[idx, :] = -np.finfo(np.float64).max
        return acq

    @abc.abstractmethod
    def _compute(self, X):
        raise NotImplementedError()


class EI(AbstractAcquisitionFunction):
    def __init__(self, model, par=0.0, **kwargs):
        super(EI, self).__init__(model)
        self.long_name = 'Expected Improvement'
        self.par = par
        self.eta = None

    def _var_floor(self):
        # facades floor at 1e-10 (base_facade.py:140), plain models at 1e-5 (base_model.py:247)
        return float(self.model.var_threshold)

    def _require_eta(self):
        if self.eta is None:
            raise ValueError('No current best specified. Call update('
                             'eta=<int>) to inform the acquisition function '
                             'about the current best value.')

    def _fused(self, X_dev, **kw):
        model = self.model
        if hasattr(model, '_fused'):
            return model._fused(X_dev, eta=self.eta, par=self.par, var_floor=self._var_floor(), **kw)
        # a single GaussianProcess: ensemble of one
        import torch
        from tlbo_b200 import ops
        if not model.is_trained:
            raise Exception('Model has to be trained first!')
        w = torch.ones(1, dtype=torch.float64, device=X_dev.device)
        return ops.predict_ei_fused(model._view(), model.types, w, None, X_dev, self.eta, par=self.par,
                                    var_floor=self._var_floor(), nmax=model.n_train_, **kw)

    def _to_device(self, X):
        import torch
        from tlbo_b200 import ops
        if torch.is_tensor(X):
            return X
        X = np.array(X, dtype=np.float64)
        X[~np.isfinite(X)] = -1
        return ops.h2d(X)

    def _compute(self, X, **kwargs):
        """EI(X) as ``float64[N, 1]`` (acquisition.py:130-177)."""
        if len(X.shape) == 1:
            X = X[:, np.newaxis]
        self._require_eta()
        if hasattr(self.model, 'acq_small') and not hasattr(X, 'is_cuda'):
            # small host query sets (online local search): mapped pinned buffers, one synchronisation
            fast = self.model.acq_small(X, self.eta, self.par, self._var_floor())
            if fast is not None:
                if fast[1] > 0:
                    raise ValueError("Expected Improvement is smaller than 0 for at least one sample.")
                return fast[0].reshape((-1, 1))
        res, mu, var, ei = self._fused(self._to_device(X), want_rows=True)
        if res[3] > 0:
            raise ValueError("Expected Improvement is smaller than 0 for at least one sample.")
        from tlbo_b200 import ops
        return ops.d2h(ei).reshape((-1, 1))

    def best_unevaluated(self, X_dev, keys=None, excluded=None, idx_offset=0, ws=None, sync=True):
        """(ei, key, row) of the first candidate in ``lexsort((keys, EI))[::-1]`` order that
        is not in ``excluded`` -- without materialising EI on the host."""
        self._require_eta()
        res = self._fused(X_dev, keys=keys, excluded=excluded, idx_offset=idx_offset, ws=ws, sync=sync)
        if sync and res[3] > 0:
            raise ValueError("Expected Improvement is smaller than 0 for at least one sample.")
        return res
