"""Surrogate factory with the reference's signature
(reference tlbo/model/model_builder.py:11-25 ``build_model``, :28-98 ``create_gp_model``)."""
import numpy as np

from tlbo_b200.config_space import get_types
from tlbo_b200.model.gp import GaussianProcess, KernelInfo


def build_model(model_type, config_space, rng):
    """``'gp'`` (and the alias ``'gp_b200'``) build the CUDA-backed GP.  ``'rf'`` (pyrfr),
    ``'gp_mcmc'`` and ``'gp_rbf'`` are outside this hot path and raise."""
    types, bounds = get_types(config_space)
    if model_type in ('gp', 'gp_b200'):
        return create_gp_model(model_type, config_space, types, bounds, rng)
    if model_type == 'rf' or 'gp' in model_type:
        raise ValueError("Surrogate type %r is not part of the B200 hot path (supported: 'gp')." % model_type)
    raise ValueError("Invalid model str %s!" % model_type)


def create_gp_model(model_type, config_space, types, bounds, rng):
    """Kernel tree ``cov_amp * (matern52[cont] * hamming[cat]) + white`` with a lognormal
    prior on the amplitude and a horseshoe(0.1) prior on the noise, both drawing from
    ``rng``; the model's own seed is the next ``rng.randint(0, 10000)``
    (model_builder.py:28-98)."""
    kernel = KernelInfo(types)
    return GaussianProcess(configspace=config_space, types=types, bounds=bounds, kernel=kernel, normalize_y=True,
                           seed=rng.randint(low=0, high=10000), prior_rng=rng)
