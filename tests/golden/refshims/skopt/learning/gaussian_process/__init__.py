"""Stand-in for scikit-optimize's GaussianProcessRegressor (see ../../../README.md)."""
import warnings

import numpy as np
from scipy.linalg import solve_triangular
from sklearn.gaussian_process import GaussianProcessRegressor as _SkGPR


class GaussianProcessRegressor(_SkGPR):
    # scikit-learn >= 1.2 validates constructor arguments; the reference passes
    # n_restarts_optimizer=-1 and scikit-optimize's extra ``noise`` argument
    _parameter_constraints = dict(_SkGPR._parameter_constraints)
    _parameter_constraints['n_restarts_optimizer'] = 'no_validation'
    _parameter_constraints['noise'] = 'no_validation'

    def __init__(self, kernel=None, alpha=1e-10, optimizer='fmin_l_bfgs_b', n_restarts_optimizer=0,
                 normalize_y=False, copy_X_train=True, random_state=None, noise=None):
        super().__init__(kernel=kernel, alpha=alpha, optimizer=optimizer,
                         n_restarts_optimizer=n_restarts_optimizer, normalize_y=normalize_y,
                         copy_X_train=copy_X_train, random_state=random_state)
        self.noise = noise

    def fit(self, X, y):
        if self.noise is not None:
            raise NotImplementedError('the reference always passes noise=None')
        super().fit(X, y)
        self.noise_ = None
        # scikit-optimize: pre-compute K^-1 = L^-T L^-1 for predict
        L_inv = solve_triangular(self.L_.T, np.eye(self.L_.shape[0]))
        self.K_inv_ = L_inv.dot(L_inv.T)
        return self

    def predict(self, X, return_std=False, return_cov=False, return_mean_grad=False, return_std_grad=False):
        if return_std and return_cov:
            raise RuntimeError('Not returning standard deviation of predictions when returning full covariance.')
        if return_cov or not return_std:
            return super().predict(X, return_std=False, return_cov=return_cov)
        X = np.asarray(X)
        K_trans = self.kernel_(X, self.X_train_)
        y_mean = K_trans.dot(self.alpha_)
        y_mean = self._y_train_std * y_mean + self._y_train_mean
        y_var = self.kernel_.diag(X)
        y_var -= np.einsum('ki,kj,ij->k', K_trans, K_trans, self.K_inv_)
        y_var_negative = y_var < 0
        if np.any(y_var_negative):
            warnings.warn('Predicted variances smaller than 0. Setting those variances to 0.')
            y_var[y_var_negative] = 0.0
        y_std = np.sqrt(y_var) * self._y_train_std
        return y_mean, y_std
