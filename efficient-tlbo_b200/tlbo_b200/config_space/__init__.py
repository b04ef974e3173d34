"""ConfigSpace-free stand-ins for the reference's configuration objects.

The reference encodes configurations with ConfigSpace 0.4.12 (absent here) and feeds
the hot path ``convert_configurations_to_array(configs) -> float64[N, d]``
(tlbo/config_space/util.py:11-30) plus ``get_types`` (tlbo/model/util_funcs.py:14-55).
The drop-in contract is therefore: a float64 table and a ``types`` vector.  These light
classes carry exactly that, with the attribute / method names the callers use."""
import numpy as np


class TypesArray(np.ndarray):
    """``get_types``' vector with one extra fact riding along: ``has_conditions`` (the space has conditional
    hyper-parameters, reference tlbo/model/base_gp.py:134-146 ``_set_has_conditions``).  The device kernels then
    multiply every kernel value by the activity mask of tlbo/model/gp_kernels.py:21-29; the flag travels to the C
    ABI as bit 30 of ``types[0]`` (``tlbo_b200.ops._types_arr``)."""

    def __new__(cls, types, has_conditions=False):
        obj = np.asarray(types, dtype=np.uint).view(cls)
        obj.has_conditions = bool(has_conditions)
        return obj

    def __array_finalize__(self, obj):
        self.has_conditions = getattr(obj, 'has_conditions', False)


class TableSpace(object):
    """Search-space descriptor: ``types[d]`` (0 = continuous / integer / constant,
    k > 0 = categorical with k choices) and per-column bounds as ``get_types`` returns.
    ``has_conditions``: the space has conditional hyper-parameters (``len(cs.get_conditions()) > 0``)."""

    def __init__(self, types, bounds=None, default=None, const_mask=None, has_conditions=False):
        self.has_conditions = bool(has_conditions or getattr(types, 'has_conditions', False))
        self.types = TypesArray(types, self.has_conditions)
        d = len(self.types)
        # columns holding a Constant hyper-parameter (encoded 0, no neighbours, never sampled)
        self.const_mask = (np.zeros(d, dtype=bool) if const_mask is None
                           else np.asarray(const_mask, dtype=bool).reshape(d))
        self._rng = np.random.RandomState()
        if bounds is None:
            bounds = [(float(t), np.nan) if t > 0 else (0.0, 1.0) for t in self.types]
        self.bounds = np.array(bounds, dtype=object)
        self.default = None if default is None else np.asarray(default, dtype=np.float64).reshape(d)
        self._seed = None

    def seed(self, seed):
        """``ConfigurationSpace.seed``: re-seeds the sampler behind ``sample_configuration``."""
        self._seed = seed
        self._rng = np.random.RandomState(seed)

    def sample_array(self, size):
        """``size`` configurations in the reference's array encoding: unit-interval floats for
        numerical columns, choice indices for categorical ones, 0 for constants (column by
        column from the space's own RandomState, like ConfigSpace's vectorised sampler)."""
        d = len(self.types)
        X = np.zeros((int(size), d), dtype=np.float64)
        for j in range(d):
            if self.types[j] > 0:
                X[:, j] = self._rng.randint(0, int(self.types[j]), size=int(size))
            elif not self.const_mask[j]:
                X[:, j] = self._rng.uniform(0.0, 1.0, size=int(size))
        return X

    def sample_batch(self, size):
        """The same draws as ``sample_configuration(size)`` (size > 1) without building ``size`` objects: a
        ``ConfigurationBatch`` over the sampled array (the online random search scores ~5000 configurations per
        iteration and looks at one of them)."""
        return ConfigurationBatch(self, self.sample_array(size))

    def sample_configuration(self, size=1):
        """``ConfigurationSpace.sample_configuration``: one Configuration for ``size == 1``, a
        list otherwise."""
        X = self.sample_array(size)
        if size == 1:
            return Configuration(self, X[0], -1)
        return [Configuration(self, X[i], -1) for i in range(X.shape[0])]

    def get_types(self):
        return TypesArray(self.types.copy(), self.has_conditions), self.bounds.copy()

    def get_conditions(self):
        """Non-empty iff the space has conditions (only the count is read: base_gp.py:135)."""
        return [True] if self.has_conditions else []

    def get_default_configuration(self):
        return None if self.default is None else Configuration(self, self.default, -1)

    def __len__(self):
        return len(self.types)


class Configuration(object):
    """One row of a candidate table.  Hashable / comparable by value like ConfigSpace's."""
    __slots__ = ('space', 'array', 'index', 'origin', '_key')

    def __init__(self, space, array, index=-1):
        self.space = space
        self.array = array if (type(array) is np.ndarray and array.dtype == np.float64) else np.asarray(
            array, dtype=np.float64)
        self.index = int(index)
        self.origin = None
        self._key = None                 # value key, built on first comparison / hash

    def get_array(self):
        return self.array

    def key(self):
        if self._key is None:
            self._key = self.array.tobytes()
        return self._key

    def __eq__(self, other):
        return isinstance(other, Configuration) and self.key() == other.key()

    def __hash__(self):
        return hash(self.key())

    def __repr__(self):
        return 'Configuration(index=%d, %s)' % (self.index, np.array2string(self.array, precision=4))


class ConfigurationBatch(object):
    """A read-only list of Configurations held as the rows of one array; elements are materialised on access
    (``batch[i]``, iteration), slices and index arrays give sub-batches.  ``convert_configurations_to_array`` returns
    the array itself."""
    __slots__ = ('space', 'array', 'origin')

    def __init__(self, space, array, origin=None):
        self.space = space
        self.array = np.ascontiguousarray(array, dtype=np.float64)
        self.origin = origin

    def __len__(self):
        return self.array.shape[0]

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            c = Configuration(self.space, self.array[i], -1)
            c.origin = self.origin
            return c
        return ConfigurationBatch(self.space, self.array[i], self.origin)

    def __iter__(self):
        for i in range(self.array.shape[0]):
            yield self[i]


def get_types(config_space, instance_features=None):
    """``get_types`` (reference tlbo/model/util_funcs.py:14-55) for a TableSpace or a bare
    types vector."""
    if hasattr(config_space, 'get_types'):
        types, bounds = config_space.get_types()
    else:
        types = TypesArray(config_space, getattr(config_space, 'has_conditions', False))
        bounds = TableSpace(types).bounds
    if instance_features is not None:
        types = TypesArray(np.hstack((types, np.zeros((instance_features.shape[1])))).astype(np.uint),
                           getattr(types, 'has_conditions', False))
    return types, bounds


def convert_configurations_to_array(configs):
    """reference tlbo/config_space/util.py:11-30: stack ``get_array()`` rows (float64)."""
    if isinstance(configs, np.ndarray):
        return np.asarray(configs, dtype=np.float64)
    if isinstance(configs, ConfigurationBatch):
        return configs.array
    if len(configs) == 0:
        return np.zeros((0, 0), dtype=np.float64)
    X = np.array([c.get_array() for c in configs], dtype=np.float64)
    cs = getattr(configs[0], 'configuration_space', None)
    if cs is not None and hasattr(cs, 'get_hyperparameters'):
        # tlbo/config_space/util.py:33-60: inactive (non-finite) entries take the normalised default
        from tlbo_b200.config_space.space import impute_default_values
        X = impute_default_values(cs, X)
    return X


def table_to_configurations(space, X):
    return [Configuration(space, X[i], i) for i in range(X.shape[0])]


# The reference binds these two at tlbo/config_space/__init__.py:10
# (``partial(get_one_exchange_neighbourhood, stdev=0.05, num_neighbors=8)``) -- not ConfigSpace's own defaults.
NUM_NEIGHBORS, NEIGHBOUR_STDEV = 8, 0.05


def one_exchange_neighbourhood_arrays_py(config, seed, num_neighbors=NUM_NEIGHBORS, stdev=NEIGHBOUR_STDEV):
    """Readable Python statement of ``one_exchange_neighbourhood_arrays`` (the product calls the
    native ``tlbo_one_exchange_neighbourhood``; tests check the two draw for draw).

    The complete one-exchange neighbourhood of ``config`` as an array ``[m, d]``, in the order
    ConfigSpace 0.4.x's lazy generator ``util.get_one_exchange_neighbourhood(configuration,
    seed)`` would yield it (the reference calls it at tlbo/optimizer/ei_optimization.py:266-267;
    ConfigSpace itself is an un-vendored dependency, restated from its published algorithm):

    a private ``RandomState(seed)`` repeatedly picks a hyper-parameter index; a numerical column
    yields ONE neighbour per pick -- ``normal(value, stdev)`` redrawn until it falls in [0, 1] --
    up to ``num_neighbors`` picks; a categorical column's other choices are shuffled once on its
    first pick and popped one per pick; constants have no neighbours.  Because the generator's
    random stream is private, materialising the whole (finite) neighbourhood up front yields the
    same sequence the lazy consumer would have seen -- which is what lets the local search score
    a full neighbourhood in one device launch."""
    space = config.space
    array = config.get_array()
    d = len(array)
    random = np.random.RandomState(seed)
    types = space.types
    remaining = np.zeros(d, dtype=np.int64)
    for j in range(d):
        if types[j] > 0:
            remaining[j] = int(types[j]) - 1
        elif not space.const_mask[j]:
            remaining[j] = num_neighbors
    usable = int(np.sum(np.isfinite(array)))
    used = int(np.sum((remaining == 0) & np.isfinite(array)))
    stacks = {}
    out = []
    while used < usable:
        index = int(random.randint(d))
        if remaining[index] == 0 or not np.isfinite(array[index]):
            continue
        value = array[index]
        if types[index] > 0:
            if index not in stacks:
                neighbors = [float(c) for c in range(int(types[index])) if c != int(value)]
                random.shuffle(neighbors)
                stacks[index] = neighbors
            neighbor = stacks[index].pop()
        else:
            neighbor = None
            for _ in range(101):                       # "no infinite loops" guard of the generator
                cand = random.normal(value, stdev)
                if 0.0 <= cand <= 1.0:
                    neighbor = cand
                    break
            if neighbor is None:
                remaining[index] = 0
                used += 1
                continue
        new_array = array.copy()
        new_array[index] = neighbor
        out.append(new_array)
        remaining[index] -= 1
        if remaining[index] == 0:
            used += 1
    return np.array(out, dtype=np.float64).reshape(-1, d)


def one_exchange_neighbourhood_arrays(config, seed, num_neighbors=NUM_NEIGHBORS, stdev=NEIGHBOUR_STDEV):
    """The complete one-exchange neighbourhood of ``config`` as an array ``[m, d]`` (see
    ``one_exchange_neighbourhood_arrays_py`` for the algorithm), generated by the native helper
    ``tlbo_one_exchange_neighbourhood`` of the C ABI: the Python loop costs more than the device
    work of a hill-climb step."""
    import ctypes
    from tlbo_b200._cabi import check, lib, ptr
    space = config.space
    array = np.ascontiguousarray(config.get_array(), dtype=np.float64)
    d = len(array)
    # per-space constants of the call (this runs once per hill-climb step): converted once
    cache = getattr(space, '_nbh_cache', None)
    if cache is None or cache[0] != num_neighbors:
        types = np.ascontiguousarray(space.types, dtype=np.int32)
        cmask = np.ascontiguousarray(space.const_mask, dtype=np.uint8)
        max_rows = int(np.sum(np.where(types > 0, types - 1, num_neighbors))) + 1
        n_out = ctypes.c_int32(0)
        cache = (num_neighbors, types, cmask, max_rows, n_out, ptr(types), ptr(cmask),
                 ctypes.cast(ctypes.byref(n_out), ctypes.c_void_p))
        try:
            space._nbh_cache = cache
        except AttributeError:
            pass
    _, types, cmask, max_rows, n_out, p_types, p_cmask, p_n = cache
    out = np.empty((max_rows, d), dtype=np.float64)
    check(lib.tlbo_one_exchange_neighbourhood(ptr(array), d, p_types, p_cmask, int(seed) & 0xffffffff,
                                              int(num_neighbors), float(stdev), ptr(out), max_rows, p_n))
    return out[:n_out.value]


def get_one_exchange_neighbourhood(config, seed, num_neighbors=NUM_NEIGHBORS, stdev=NEIGHBOUR_STDEV):
    """Generator form (the reference's call signature, tlbo/config_space/__init__.py:10)."""
    for row in one_exchange_neighbourhood_arrays(config, seed, num_neighbors, stdev):
        yield Configuration(config.space, row, -1)
