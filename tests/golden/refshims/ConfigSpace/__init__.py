"""Stand-in for ConfigSpace (see ../README.md): value classes with the attributes the reference reads.
``sample_configuration`` and ``util.get_one_exchange_neighbourhood`` run the sampling / neighbourhood
streams as RESTATED in oracle/ei_optimization.py ``OracleSpace`` (ConfigSpace itself is not on this
machine: fixtures made through them pin the reference's own maximisers, not ConfigSpace's streams)."""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from ConfigSpace.hyperparameters import (CategoricalHyperparameter, Constant, OrdinalHyperparameter,  # noqa
                                         UniformFloatHyperparameter, UniformIntegerHyperparameter)


class ConfigurationSpace(object):
    def __init__(self, seed=None):
        self._hps = []
        self.random = np.random.RandomState(seed)

    def seed(self, seed):
        self.random = np.random.RandomState(seed)

    def add_hyperparameter(self, hp):
        self._hps.append(hp)
        return hp

    def add_hyperparameters(self, hps):
        for hp in hps:
            self.add_hyperparameter(hp)

    def get_hyperparameters(self):
        return list(self._hps)

    def get_hyperparameter_names(self):
        return [hp.name for hp in self._hps]

    def get_idx_by_hyperparameter_name(self, name):
        return self.get_hyperparameter_names().index(name)

    def get_conditions(self):
        return []

    def _oracle_space(self):
        from oracle.ei_optimization import OracleSpace
        types = [len(hp.choices) if isinstance(hp, CategoricalHyperparameter) else 0 for hp in self._hps]
        const = [isinstance(hp, Constant) for hp in self._hps]
        sp = OracleSpace(types, const)
        sp.rng = self.random                      # one stream: the space's own RandomState
        return sp

    def sample_configuration(self, size=1):
        X = self._oracle_space().sample(size)
        if size == 1:
            return Configuration(self, vector=X[0])
        return [Configuration(self, vector=x) for x in X]

    def get_forbiddens(self):
        return []

    def add_condition(self, cond):
        raise NotImplementedError('stand-in spaces have no conditions')

    def add_forbidden_clause(self, clause):
        raise NotImplementedError('stand-in spaces have no forbidden clauses')


class Configuration(object):
    """Holds the internal (unit-hypercube / category-index) vector, as ConfigSpace does."""

    def __init__(self, configuration_space, vector=None, values=None):
        self.configuration_space = configuration_space
        self._vector = np.array(vector, dtype=np.float64)
        self.origin = None

    def get_array(self):
        return self._vector

    def __hash__(self):
        return hash(self._vector.tobytes())

    def __eq__(self, other):
        return isinstance(other, Configuration) and np.array_equal(self._vector, other._vector)

    def __repr__(self):
        return 'Configuration(%s)' % np.array2string(self._vector, precision=4)


class InCondition(object):
    def __init__(self, child, parent, values):
        self.child, self.parent, self.values = child, parent, values
