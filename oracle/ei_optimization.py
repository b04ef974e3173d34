"""Oracle: the acquisition maximisers of the online scenario, literal and sequential.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations are relative to
``/root/reference``.

Follows ``tlbo/optimizer/ei_optimization.py``: ``LocalSearch`` (:135-294, one acquisition call
per neighbour, stop at the first improvement), ``RandomSearch`` (:297-357),
``InterleavedLocalAndRandomSearch.maximize`` (:406-467), ``ChallengerList`` (:478-524), and the
online loop ``tlbo/framework/online/smbo_online.py`` over ``smbo_offline.py:147-264``.

Configurations are plain ``float64[d]`` arrays in the reference's encoding.  The search space
(``OracleSpace``) restates the two ConfigSpace services the maximisers use -- ConfigSpace 0.4.12
is an UN-VENDORED dependency absent from this image (``requirements.txt``), so these follow its
published algorithm as recalled and are NOT pinned against it (parity unpinned for the
neighbourhood / sampling streams).  Everything downstream of them IS pinned against the reference's
own ``SMBO_ONLINE`` + ``InterleavedLocalAndRandomSearch`` / ``LocalSearch`` / ``RandomSearch`` /
``ChallengerList`` code, run unmodified in the build container over a ConfigSpace stand-in that wraps
``OracleSpace`` (``tests/golden/make_reference_golden.py`` -> ``reference_v1.npz`` ``online_*``; checked by
``tests/test_reference_golden.py::test_online_loop_matches_reference``: same picked configurations, same
``np.random`` stream states, same weights):
  * ``sample``: column-wise draws from the space's own RandomState;
  * ``neighbourhood``: the lazy generator ``ConfigSpace.util.get_one_exchange_neighbourhood
    (configuration, seed, num_neighbors=8, stdev=0.05)`` -- the two keyword values are what the
    reference binds at ``tlbo/config_space/__init__.py:10``, not ConfigSpace's own defaults.
"""
import numpy as np

from oracle.facade import zero_mean_unit_var_normalization, zero_one_normalization


class OracleSpace(object):
    def __init__(self, types, const_mask=None):
        self.types = np.asarray(types, dtype=np.int64)
        d = len(self.types)
        self.const_mask = np.zeros(d, dtype=bool) if const_mask is None else np.asarray(const_mask, dtype=bool)
        self.rng = np.random.RandomState()

    def seed(self, seed):
        self.rng = np.random.RandomState(seed)

    def sample(self, size):
        d = len(self.types)
        X = np.zeros((size, d))
        for j in range(d):
            if self.types[j] > 0:
                X[:, j] = self.rng.randint(0, int(self.types[j]), size=size)
            elif not self.const_mask[j]:
                X[:, j] = self.rng.uniform(0.0, 1.0, size=size)
        return X

    def neighbourhood(self, array, seed, num_neighbors=8, stdev=0.05):
        """Lazy generator: one neighbour per ``next``."""
        random = np.random.RandomState(seed)
        d = len(array)
        n_left = {}
        used = set()
        for j in range(d):
            if self.types[j] > 0:
                n_left[j] = int(self.types[j]) - 1
            elif self.const_mask[j]:
                n_left[j] = 0
            else:
                n_left[j] = num_neighbors
            if n_left[j] == 0 and np.isfinite(array[j]):
                used.add(j)
        usable = int(np.sum(np.isfinite(array)))
        stacks = {}
        while len(used) < usable:
            index = int(random.randint(d))
            if n_left[index] == 0:
                continue
            if not np.isfinite(array[index]):
                continue
            value = array[index]
            neighbour = None
            iteration = 0
            while True:
                if iteration > 100:
                    break
                if self.types[index] == 0:
                    cand = random.normal(value, stdev)          # UniformFloat.get_neighbors(number=1)
                    iteration += 1
                    if cand < 0 or cand > 1:
                        continue
                    neighbour = cand
                    break
                if index not in stacks:
                    choices = [float(c) for c in range(int(self.types[index])) if c != int(value)]
                    random.shuffle(choices)
                    stacks[index] = choices
                neighbour = stacks[index].pop()
                break
            if neighbour is None:
                n_left[index] = 0
                used.add(index)
                continue
            new_array = np.array(array, dtype=np.float64)
            new_array[index] = neighbour
            n_left[index] -= 1
            if n_left[index] == 0:
                used.add(index)
            # ConfigSpace 0.4.x util.pyx: "only rigorously check every tenth configuration":
            # `if np.random.random() > 0.95: new_configuration.is_valid_configuration()` -- one draw from the
            # GLOBAL legacy stream per produced neighbour (the stream the facades' bootstrap draws use), and the
            # generator is lazy, so only the neighbours the hill-climb actually inspects consume one
            np.random.random()
            yield new_array


def sort_configs_by_acq_value(acq_fn, configs, rng):
    """ei_optimization.py:106-132."""
    acq_values = acq_fn(np.asarray(configs))
    random = rng.rand(len(acq_values))
    indices = np.lexsort((random.flatten(), acq_values.flatten()))
    return [(acq_values[ind][0], configs[ind]) for ind in indices[::-1]]


def local_search_one_iter(acq_fn, space, rng, start_point, max_steps=None, counters=None):
    """ei_optimization.py:239-294."""
    incumbent = start_point
    acq_val_incumbent = acq_fn(incumbent[np.newaxis, :])[0]
    local_search_steps = 0
    while True:
        local_search_steps += 1
        changed_inc = False
        all_neighbors = space.neighbourhood(incumbent, seed=rng.randint(0, 10000))
        for neighbor in all_neighbors:
            acq_val = acq_fn(neighbor[np.newaxis, :])
            if counters is not None:
                counters['neighbours'] += 1
            if acq_val > acq_val_incumbent:
                incumbent = neighbor
                acq_val_incumbent = acq_val[0]
                changed_inc = True
                break
        if counters is not None:
            counters['steps'] += 1
        if (not changed_inc) or (max_steps is not None and local_search_steps == max_steps):
            break
    return acq_val_incumbent, incumbent


def local_search_maximize(acq_fn, space, rng, evaluated_configs, num_points, counters=None):
    """ei_optimization.py:168-237."""
    if len(evaluated_configs) == 0:
        init_points = list(space.sample(num_points))
    else:
        srt = sort_configs_by_acq_value(acq_fn, list(evaluated_configs), rng)
        init_points = [x[1] for x in srt[:int(min(len(srt), num_points))]]
    configs_acq = []
    for start_point in init_points:
        acq_val, configuration = local_search_one_iter(acq_fn, space, rng, start_point, counters=counters)
        configs_acq.append((acq_val, configuration))
    rng.shuffle(configs_acq)
    configs_acq.sort(reverse=True, key=lambda x: x[0])
    return configs_acq


def interleaved_maximize(acq_fn, space, rng, evaluated_configs, num_points, n_sls_iterations=10, rand_prob=0.0,
                         counters=None):
    """``InterleavedLocalAndRandomSearch.maximize`` + iteration of the ``ChallengerList`` up to
    exhaustion is the caller's business: returns the generator of challengers."""
    by_local = local_search_maximize(acq_fn, space, rng, evaluated_configs, n_sls_iterations, counters)
    n_rand = num_points - len(by_local)
    rand_configs = list(space.sample(n_rand)) if n_rand > 1 else [space.sample(1)[0]]
    by_random = sort_configs_by_acq_value(acq_fn, rand_configs, rng)
    allc = by_random + by_local
    allc.sort(reverse=True, key=lambda x: x[0])
    challengers = [c[1] for c in allc]

    def gen():
        index = 0
        while index < len(challengers):
            if rng.rand() < rand_prob:                  # ChooserProb.check
                yield space.sample(1)[0]
            else:
                yield challengers[index]
                index += 1
    return gen(), allc


class OracleSMBOOnline(object):
    """``SMBO_ONLINE`` (smbo_online.py) with ``evaluate`` = ``objective_function(config)``."""

    def __init__(self, objective_function, space, surrogate_model, random_seed, initial_runs=3, num_points=5000,
                 first_config=None):
        from oracle.acquisition import OracleEI
        # smbo_offline.py:211-217: the default configuration, or -- when it is not a key of the target
        # table -- ``configuration_list[0]``, is proposed while it is unevaluated
        self.first_config = None if first_config is None else np.asarray(first_config, dtype=np.float64)
        self.objective_function = objective_function
        self.space = space
        self.random_seed = random_seed
        self.rng = np.random.RandomState(random_seed)
        self.init_num = initial_runs
        self.num_points = num_points
        self.iteration_id = 0
        self.configurations = []
        self.perfs = []
        self.space.seed(random_seed)
        np.random.seed(random_seed)
        self.model = surrogate_model
        self.acquisition_function = OracleEI(self.model)
        self.acq_rng = np.random.RandomState(random_seed)
        self.chooser_rng = np.random.RandomState(random_seed)
        self.counters = dict(steps=0, neighbours=0)
        self.last_sorted = None

    def _seen(self, config):
        return any(np.array_equal(config, c) for c in self.configurations)

    def sample_random_config(self):
        sample_cnt = 0
        while True:
            sample_cnt += 1
            config = self.space.sample(1)[0]
            if not self._seen(config):
                return config
            sample_cnt += 1
            if sample_cnt >= 200:
                return config

    def choose_next(self, X, Y):
        if X.shape[0] < self.init_num:
            if self.first_config is not None and not self._seen(self.first_config):
                return self.first_config
            return self.sample_random_config()
        if self.chooser_rng.rand() < 0.001:
            config = self.sample_random_config()
            tw = self.model.target_weight
            tw.append(0. if len(tw) == 0 else tw[-1])
            return config
        self.model.train(X, Y)
        if self.model.method_id in ['tst', 'tstm']:
            y_, _, _ = zero_one_normalization(Y)
        else:
            y_, _, _ = zero_mean_unit_var_normalization(Y)
        self.acquisition_function.update(model=self.model, eta=np.min(y_), num_data=len(self.configurations))
        challengers, self.last_sorted = interleaved_maximize(self.acquisition_function, self.space, self.acq_rng,
                                                             self.configurations, self.num_points,
                                                             counters=self.counters)
        for config in challengers:
            if not self._seen(config):
                return config
        raise ValueError('The configuration in the SET is over')

    def iterate(self):
        X = np.zeros((0, len(self.space.types))) if not self.configurations else np.array(self.configurations)
        Y = np.array(self.perfs, dtype=np.float64)
        config = self.choose_next(X, Y)
        if not self._seen(config):
            self.configurations.append(config)
            self.perfs.append(float(self.objective_function(config)))
        self.iteration_id += 1
        return config
