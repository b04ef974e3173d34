"""Surrogate factory with the reference's signature
(reference tlbo/model/model_builder.py:11-25 ``build_model``, :28-98 ``create_gp_model``)."""
import numpy as np

from tlbo_b200.config_space import get_types
from tlbo_b200.model.gp import GaussianProcess, KernelInfo
from tlbo_b200.model.gp_mcmc import GaussianProcessMCMC


def build_model(model_type, config_space, rng):
    """``'gp'`` (and the alias ``'gp_b200'``) build the CUDA-backed MAP GP, ``'gp_mcmc'`` the
    MCMC-marginalised one.  ``'rf'`` (pyrfr) and ``'gp_rbf'`` are outside this hot path and raise."""
    types, bounds = get_types(config_space)
    if model_type in ('gp', 'gp_b200', 'gp_mcmc'):
        return create_gp_model(model_type, config_space, types, bounds, rng)
    if model_type == 'rf' or 'gp' in model_type:
        raise ValueError("Surrogate type %r is not part of the B200 hot path (supported: 'gp', 'gp_mcmc')."
                         % model_type)
    raise ValueError("Invalid model str %s!" % model_type)


def create_gp_model(model_type, config_space, types, bounds, rng):
    """Kernel tree ``cov_amp * (matern52[cont] * hamming[cat]) + white`` with a lognormal
    prior on the amplitude and a horseshoe(0.1) prior on the noise, both drawing from
    ``rng``; the model's own seed is the next ``rng.randint(0, 10000)``
    (model_builder.py:28-98)."""
    kernel = KernelInfo(types)
    if model_type == 'gp_mcmc':
        # model_builder.py:74-89: 3 walkers per hyper-parameter, rounded up to even; 250 + 250 steps
        n_mcmc_walkers = 3 * kernel.n_dims
        if n_mcmc_walkers % 2 == 1:
            n_mcmc_walkers += 1
        return GaussianProcessMCMC(configspace=config_space, types=types, bounds=bounds, kernel=kernel,
                                   n_mcmc_walkers=n_mcmc_walkers, chain_length=250, burnin_steps=250,
                                   normalize_y=True, seed=rng.randint(low=0, high=10000), prior_rng=rng)
    return GaussianProcess(configspace=config_space, types=types, bounds=bounds, kernel=kernel, normalize_y=True,
                           seed=rng.randint(low=0, high=10000), prior_rng=rng)
