"""ConfigSpace-free stand-ins for the reference's configuration objects.

The reference encodes configurations with ConfigSpace 0.4.12 (absent here) and feeds
the hot path ``convert_configurations_to_array(configs) -> float64[N, d]``
(tlbo/config_space/util.py:11-30) plus ``get_types`` (tlbo/model/util_funcs.py:14-55).
The drop-in contract is therefore: a float64 table and a ``types`` vector.  These light
classes carry exactly that, with the attribute / method names the callers use."""
import numpy as np


class TableSpace(object):
    """Search-space descriptor: ``types[d]`` (0 = continuous / integer / constant,
    k > 0 = categorical with k choices) and per-column bounds as ``get_types`` returns."""

    def __init__(self, types, bounds=None, default=None):
        self.types = np.asarray(types, dtype=np.uint)
        d = len(self.types)
        if bounds is None:
            bounds = [(float(t), np.nan) if t > 0 else (0.0, 1.0) for t in self.types]
        self.bounds = np.array(bounds, dtype=object)
        self.default = None if default is None else np.asarray(default, dtype=np.float64).reshape(d)
        self._seed = None

    def seed(self, seed):
        self._seed = seed

    def get_types(self):
        return self.types.copy(), self.bounds.copy()

    def get_default_configuration(self):
        return None if self.default is None else Configuration(self, self.default, -1)

    def __len__(self):
        return len(self.types)


class Configuration(object):
    """One row of a candidate table.  Hashable / comparable by value like ConfigSpace's."""
    __slots__ = ('space', 'array', 'index', 'origin', '_key')

    def __init__(self, space, array, index=-1):
        self.space = space
        self.array = np.asarray(array, dtype=np.float64)
        self.index = int(index)
        self.origin = None
        self._key = self.array.tobytes()

    def get_array(self):
        return self.array

    def __eq__(self, other):
        return isinstance(other, Configuration) and self._key == other._key

    def __hash__(self):
        return hash(self._key)

    def __repr__(self):
        return 'Configuration(index=%d, %s)' % (self.index, np.array2string(self.array, precision=4))


def get_types(config_space, instance_features=None):
    """``get_types`` (reference tlbo/model/util_funcs.py:14-55) for a TableSpace or a bare
    types vector."""
    if isinstance(config_space, TableSpace):
        types, bounds = config_space.get_types()
    else:
        types = np.asarray(config_space, dtype=np.uint)
        bounds = TableSpace(types).bounds
    if instance_features is not None:
        types = np.hstack((types, np.zeros((instance_features.shape[1])))).astype(np.uint)
    return types, bounds


def convert_configurations_to_array(configs):
    """reference tlbo/config_space/util.py:11-30: stack ``get_array()`` rows (float64)."""
    if isinstance(configs, np.ndarray):
        return np.asarray(configs, dtype=np.float64)
    if len(configs) == 0:
        return np.zeros((0, 0), dtype=np.float64)
    return np.array([c.get_array() for c in configs], dtype=np.float64)


def table_to_configurations(space, X):
    return [Configuration(space, X[i], i) for i in range(X.shape[0])]
