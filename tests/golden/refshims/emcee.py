"""Stand-in for emcee 3.x (see README.md): the EnsembleSampler surface the reference calls
(constructor, ``random_state`` setter, ``run_mcmc`` -> 3-tuple-like state, ``get_chain``), running the
stretch move as RESTATED in oracle/gp_mcmc.py ``emcee_run``.  emcee itself is not on this machine, so
fixtures made through this module pin the reference's own GaussianProcessMCMC code (``_ll``, prior
sampling of the initial ensemble, walker-model construction, marginalised prediction) but NOT the
sampler against emcee."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


class State(object):
    def __init__(self, coords, log_prob, random_state):
        self.coords, self.log_prob, self.random_state = coords, log_prob, random_state

    def __iter__(self):
        return iter((self.coords, self.log_prob, self.random_state))


class EnsembleSampler(object):
    def __init__(self, nwalkers, ndim, log_prob_fn, **kwargs):
        if kwargs:
            raise NotImplementedError('the reference passes no further arguments')
        self.nwalkers, self.ndim, self.log_prob_fn = nwalkers, ndim, log_prob_fn
        self._random = np.random.mtrand.RandomState()
        self._chain = []
        self.naccepted = np.zeros(nwalkers, dtype=np.int64)

    @property
    def random_state(self):
        return self._random.get_state()

    @random_state.setter
    def random_state(self, state):
        self._random.set_state(state)

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        from oracle.gp_mcmc import emcee_run
        coords, log_prob, nacc = emcee_run(self.log_prob_fn, np.array(initial_state, dtype=np.float64), nsteps,
                                           self._random)
        self.naccepted += nacc
        self._chain.append(np.array(coords))
        return State(coords, log_prob, self.random_state)

    def get_chain(self):
        # only the last stored step is ever read (gp_mcmc.py: ``sampler.get_chain()[-1]``)
        return np.array(self._chain)
