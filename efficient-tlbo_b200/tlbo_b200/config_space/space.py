"""ConfigSpace-free configuration spaces: the classes the reference builds its search spaces from
(tlbo/config_space/space_instance.py; ``from ConfigSpace import ...``) with the ARRAY ENCODING the hot path
consumes -- ``Configuration.get_array()``, ``convert_configurations_to_array`` / ``impute_default_values``
(tlbo/config_space/util.py:11-60) and ``get_types`` (tlbo/model/util_funcs.py:14-55).

ConfigSpace 0.4.12 itself is an un-vendored dependency absent from this machine; the encoding below restates its
published behaviour: floats / integers on the unit interval (integers through a float range widened by 0.49999 on
both sides; ``log=True`` in log space), categorical / ordinal values as their index, constants as 0, inactive
(conditional) hyper-parameters as NaN; hyper-parameters ordered parents-before-children and alphabetically within
a level.  **Parity unpinned** against ConfigSpace (stated in DESIGN.md); everything downstream of the arrays is
pinned against the reference.  ``load_history`` reads histories exported as JSON lines by
``tools/export_history.py`` (run where ConfigSpace and the reference's pickles live)."""
import collections
import json
import math

import numpy as np

from tlbo_b200.config_space import TypesArray


class _Hyperparameter(object):
    def __init__(self, name, default_value):
        self.name = name
        self.default_value = default_value

    @property
    def normalized_default_value(self):
        return self._transform(self.default_value)


class UniformFloatHyperparameter(_Hyperparameter):
    def __init__(self, name, lower, upper, default_value=None, q=None, log=False):
        self.lower, self.upper, self.q, self.log = float(lower), float(upper), q, bool(log)
        if default_value is None:
            default_value = math.exp((math.log(lower) + math.log(upper)) / 2.) if log else (lower + upper) / 2.
        super().__init__(name, float(default_value))
        self._lo = math.log(self.lower) if log else self.lower
        self._hi = math.log(self.upper) if log else self.upper

    def _transform(self, value):
        v = math.log(value) if self.log else float(value)
        return (v - self._lo) / (self._hi - self._lo)

    def _inverse_transform(self, x):
        v = x * (self._hi - self._lo) + self._lo
        return math.exp(v) if self.log else v


class UniformIntegerHyperparameter(_Hyperparameter):
    def __init__(self, name, lower, upper, default_value=None, q=None, log=False):
        self.lower, self.upper, self.q, self.log = int(lower), int(upper), q, bool(log)
        if default_value is None:
            default_value = int(round(math.exp((math.log(lower) + math.log(upper)) / 2.) if log
                                      else (lower + upper) / 2.))
        super().__init__(name, int(default_value))
        # ConfigSpace maps integers through a float hyper-parameter on [lower - 0.49999, upper + 0.49999]
        self._ufhp = UniformFloatHyperparameter(name, self.lower - 0.49999, self.upper + 0.49999, log=log,
                                                default_value=float(self.default_value))

    def _transform(self, value):
        return self._ufhp._transform(value)

    def _inverse_transform(self, x):
        return int(round(self._ufhp._inverse_transform(x)))


class CategoricalHyperparameter(_Hyperparameter):
    def __init__(self, name, choices, default_value=None, weights=None):
        self.choices = tuple(choices)
        super().__init__(name, self.choices[0] if default_value is None else default_value)

    def _transform(self, value):
        return float(self.choices.index(value))

    def _inverse_transform(self, x):
        return self.choices[int(round(x))]


class OrdinalHyperparameter(_Hyperparameter):
    def __init__(self, name, sequence, default_value=None):
        self.sequence = tuple(sequence)
        super().__init__(name, self.sequence[0] if default_value is None else default_value)

    def _transform(self, value):
        return float(self.sequence.index(value))

    def _inverse_transform(self, x):
        return self.sequence[int(round(x))]


class Constant(_Hyperparameter):
    def __init__(self, name, value):
        self.value = value
        super().__init__(name, value)

    def _transform(self, value):
        return 0.0

    def _inverse_transform(self, x):
        return self.value


class UnParametrizedHyperparameter(Constant):
    pass


class EqualsCondition(object):
    def __init__(self, child, parent, value):
        self.child, self.parent, self.value = child, parent, value

    def holds(self, values):
        return self.parent.name in values and values[self.parent.name] == self.value


class InCondition(object):
    def __init__(self, child, parent, values):
        self.child, self.parent, self.values = child, parent, tuple(values)

    def holds(self, values):
        return self.parent.name in values and values[self.parent.name] in self.values


class ConfigurationSpace(object):
    def __init__(self, name=None, seed=None):
        self.name = name
        self._hps = collections.OrderedDict()
        self._conditions = []
        self._order = None
        self.random = np.random.RandomState(seed)

    def seed(self, seed):
        self.random = np.random.RandomState(seed)

    def add_hyperparameter(self, hp):
        if hp.name in self._hps:
            raise ValueError("Hyperparameter '%s' is already in the configuration space." % hp.name)
        self._hps[hp.name] = hp
        self._order = None
        return hp

    def add_hyperparameters(self, hps):
        for hp in hps:
            self.add_hyperparameter(hp)
        return hps

    def add_condition(self, cond):
        self._conditions.append(cond)
        self._order = None
        return cond

    def add_conditions(self, conds):
        for c in conds:
            self.add_condition(c)
        return conds

    def get_conditions(self):
        return list(self._conditions)

    def get_forbiddens(self):
        return []

    def _sorted(self):
        if self._order is None:
            parents = collections.defaultdict(set)
            for c in self._conditions:
                parents[c.child.name].add(c.parent.name)
            level = {}

            def depth(n):
                if n not in level:
                    level[n] = 0 if not parents[n] else 1 + max(depth(p) for p in parents[n])
                return level[n]
            self._order = sorted(self._hps, key=lambda n: (depth(n), n))
        return self._order

    def get_hyperparameters(self):
        return [self._hps[n] for n in self._sorted()]

    def get_hyperparameter_names(self):
        return list(self._sorted())

    def get_hyperparameter(self, name):
        return self._hps[name]

    def get_idx_by_hyperparameter_name(self, name):
        return self._sorted().index(name)

    def is_active(self, name, values):
        conds = [c for c in self._conditions if c.child.name == name]
        return all(c.holds(values) for c in conds)

    def get_default_configuration(self):
        values = {}
        for hp in self.get_hyperparameters():
            if self.is_active(hp.name, values):
                values[hp.name] = hp.default_value
        return Configuration(self, values=values)

    def get_types(self):
        """tlbo/model/util_funcs.py:14-55 ``get_types``."""
        hps = self.get_hyperparameters()
        types = np.zeros(len(hps), dtype=np.uint)
        bounds = [(np.nan, np.nan)] * len(hps)
        for i, param in enumerate(hps):
            if isinstance(param, CategoricalHyperparameter):
                types[i] = len(param.choices)
                bounds[i] = (int(len(param.choices)), np.nan)
            elif isinstance(param, OrdinalHyperparameter):
                bounds[i] = (0, int(len(param.sequence)) - 1)
            elif isinstance(param, Constant):
                bounds[i] = (0, np.nan)
            elif isinstance(param, (UniformFloatHyperparameter, UniformIntegerHyperparameter)):
                bounds[i] = (0.0, 1.0)
            else:
                raise TypeError("Unknown hyperparameter type %s" % type(param))
        return TypesArray(types, len(self._conditions) > 0), np.array(bounds, dtype=object)

    def __len__(self):
        return len(self._hps)


class Configuration(object):
    """A point of a ``ConfigurationSpace``: ``values`` (dict name -> value, inactive hyper-parameters absent) or the
    internal ``vector``; hashable / comparable by its vector like ConfigSpace's."""

    def __init__(self, configuration_space, values=None, vector=None):
        self.configuration_space = self.space = configuration_space
        self.origin = None
        self.index = -1
        names = configuration_space.get_hyperparameter_names()
        if vector is not None:
            self._vector = np.array(vector, dtype=np.float64)
            self._values = None
        else:
            vec = np.full(len(names), np.nan)
            for i, n in enumerate(names):
                if n in values and values[n] is not None and configuration_space.is_active(n, values):
                    vec[i] = configuration_space.get_hyperparameter(n)._transform(values[n])
            self._vector = vec
            self._values = dict(values)

    def get_array(self):
        return self._vector

    def get_dictionary(self):
        if self._values is None:
            cs = self.configuration_space
            self._values = {n: cs.get_hyperparameter(n)._inverse_transform(float(v))
                            for n, v in zip(cs.get_hyperparameter_names(), self._vector) if np.isfinite(v)}
        return dict(self._values)

    def key(self):
        return self._vector.tobytes()

    def __hash__(self):
        return hash(self._vector.tobytes())

    def __eq__(self, other):
        return isinstance(other, Configuration) and np.array_equal(self._vector, other._vector, equal_nan=True)

    def __repr__(self):
        return 'Configuration(%r)' % (self.get_dictionary(),)


def impute_default_values(configuration_space, configs_array):
    """tlbo/config_space/util.py:33-60: inactive (non-finite) entries take the hyper-parameter's normalised
    default."""
    for hp in configuration_space.get_hyperparameters():
        default = hp.normalized_default_value
        idx = configuration_space.get_idx_by_hyperparameter_name(hp.name)
        nonfinite_mask = ~np.isfinite(configs_array[:, idx])
        configs_array[nonfinite_mask, idx] = default
    return configs_array


def convert_configurations_to_array(configs):
    """tlbo/config_space/util.py:11-30."""
    configs_array = np.array([config.get_array() for config in configs], dtype=np.float64)
    return impute_default_values(configs[0].configuration_space, configs_array)


def load_history(path, configuration_space):
    """A benchmark history exported by ``tools/export_history.py`` -- JSON lines ``{"config": {name: value},
    "perf": float}`` in the pickle's insertion order -- as the reference's
    ``OrderedDict{Configuration -> perf}`` (tools/offline_benchmark.py:154-176)."""
    hist = collections.OrderedDict()
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            rec = json.loads(line)
            hist[Configuration(configuration_space, values=rec['config'])] = rec['perf']
    return hist
