"""Stand-in for skopt.learning.gaussian_process.kernels (see ../../../README.md): scikit-learn's
kernel classes under scikit-optimize's names, plus HammingKernel's declaration."""
import numpy as np
from sklearn.gaussian_process.kernels import (ConstantKernel, DotProduct, Exponentiation, ExpSineSquared,  # noqa
                                              Hyperparameter, Kernel, Matern, NormalizedKernelMixin, Product,
                                              RationalQuadratic, RBF, StationaryKernelMixin, Sum, WhiteKernel)


class HammingKernel(StationaryKernelMixin, NormalizedKernelMixin, Kernel):
    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5)):
        self.length_scale = length_scale
        self.length_scale_bounds = length_scale_bounds

    @property
    def anisotropic(self):
        return np.iterable(self.length_scale) and len(self.length_scale) > 1

    @property
    def hyperparameter_length_scale(self):
        if self.anisotropic:
            return Hyperparameter('length_scale', 'numeric', self.length_scale_bounds, len(self.length_scale))
        return Hyperparameter('length_scale', 'numeric', self.length_scale_bounds)

    def __call__(self, X, Y=None, eval_gradient=False):
        raise NotImplementedError('the reference overrides _call / __call__')
