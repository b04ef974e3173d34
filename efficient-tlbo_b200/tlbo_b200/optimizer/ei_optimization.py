"""Acquisition maximisers of the ONLINE scenario on the device path
(reference tlbo/optimizer/ei_optimization.py: ``LocalSearch`` :135-294, ``RandomSearch``
:297-357, ``InterleavedLocalAndRandomSearch`` :360-475, ``ChallengerList`` :478-524;
wired in by tlbo/framework/online/smbo_online.py:18-26).

The reference scores ONE neighbour per acquisition call inside each hill-climb
(``_one_iter``, :270-281): ~10^2 sequential calls per maximisation, each a Python loop over
K + 1 GP predicts.  Here a hill-climb step scores the incumbent's WHOLE one-exchange
neighbourhood in one fused ensemble-predict + EI launch and then takes the first improving
neighbour -- exactly the neighbour the reference's early ``break`` would have stopped at,
because the neighbourhood generator draws its neighbours from a private ``RandomState(seed)``
and the acquisition is deterministic.  (ConfigSpace's lazy generator additionally draws ONE
``np.random.random()`` from the GLOBAL legacy stream per neighbour it produces -- its "only
rigorously check every tenth configuration" test --, i.e. one per neighbour the reference's loop
inspects; the eager path consumes that many draws after scoring, see ``_one_iter``.  Integer
hyper-parameters are walked like floats here: ConfigSpace enumerates a finite, shuffled
neighbour list for ``UniformIntegerHyperparameter`` -- a documented deviation, the neighbourhood
STREAMS are parity-unpinned either way because ConfigSpace is absent from this machine.)  The ten
start points are scored in one launch, the ~4990 random configurations in another.  Every
draw from the shared ``rng`` (ordering keys, per-step generator seeds, the tie-break shuffle,
the interleaving chooser) happens in the reference's order."""
import logging

import numpy as np

from tlbo_b200.config_space import (NEIGHBOUR_STDEV, NUM_NEIGHBORS, Configuration, ConfigurationBatch,
                                    convert_configurations_to_array, one_exchange_neighbourhood_arrays)
from tlbo_b200.optimizer.ei_offline_optimizer import AcquisitionFunctionMaximizer
from tlbo_b200.optimizer.random_configuration_chooser import ChooserProb


class _ScoredBatch(object):
    """``[(acq, config), ...]`` over a ``ConfigurationBatch``: the sorted random search's result without ~5000
    tuples and Configuration objects per iteration.  Behaves as the list it stands for (len / index / iterate /
    ``+ list``); ``InterleavedLocalAndRandomSearch.maximize`` merges it with array operations."""

    def __init__(self, acq, batch):
        self.acq = acq
        self.batch = batch

    def __len__(self):
        return len(self.batch)

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            return (self.acq[i], self.batch[i])
        return _ScoredBatch(self.acq[i], self.batch[i])

    def __iter__(self):
        for i in range(len(self.batch)):
            yield (self.acq[i], self.batch[i])

    def __add__(self, other):
        return list(self) + list(other)


class _MergedChallengers(object):
    """The challengers of one maximisation in final order -- ``sorted(random_sorted + local, reverse=True, key=acq)``
    (ei_optimization.py:446-456), a STABLE descending sort -- as an index permutation over the random batch and the
    local-search results; Configurations are materialised when ``ChallengerList`` reaches them."""

    def __init__(self, scored, local):
        acq_local = np.array([float(np.asarray(a).reshape(-1)[0]) for a, _ in local], dtype=np.float64)
        vals = np.concatenate((np.asarray(scored.acq, dtype=np.float64), acq_local))
        self.order = np.argsort(-vals, kind='stable')
        self.batch = scored.batch
        self.local = [c for _, c in local]

    def __len__(self):
        return len(self.order)

    def __getitem__(self, i):
        j = int(self.order[i])
        nb = len(self.batch)
        return self.batch[j] if j < nb else self.local[j - nb]

    def __iter__(self):
        for i in range(len(self.order)):
            yield self[i]


class _SortMixin(object):
    def _sort_configs_by_acq_value(self, configs):
        """ei_optimization.py:106-132: descending by acquisition value, ties by a fresh
        ``rng.rand(N)`` key."""
        acq_values = self.acquisition_function(configs)
        random = self.rng.rand(len(acq_values))
        indices = np.lexsort((random.flatten(), acq_values.flatten()))
        if isinstance(configs, ConfigurationBatch):
            order = indices[::-1]
            return _ScoredBatch(acq_values[order, 0], configs[order])
        return [(acq_values[ind][0], configs[ind]) for ind in indices[::-1]]


class LocalSearch(_SortMixin, AcquisitionFunctionMaximizer):
    def __init__(self, acquisition_function, config_space, rng=None, max_steps=None, n_steps_plateau_walk=10):
        super().__init__(acquisition_function, config_space, rng)
        self.max_steps = max_steps
        self.n_steps_plateau_walk = n_steps_plateau_walk
        self.logger = logging.getLogger(self.__module__ + '.' + self.__class__.__name__)
        self.last_stats = None
        self.native_climb = True          # one C-ABI call per hill-climb (tlbo_local_search_climb); False: per-step Python loop

    def _maximize(self, runhistory, num_points, **kwargs):
        init_points = self._get_initial_points(num_points, runhistory)
        # the incumbents' own acquisition values do not depend on any random draw: one launch
        start_acq = self.acquisition_function(init_points) if len(init_points) else np.zeros((0, 1))
        configs_acq = []
        stats = dict(steps=0, neighbours=0, launches=1)
        for i, start_point in enumerate(init_points):
            acq_val, configuration = self._one_iter(start_point, start_acq[i], stats)
            configuration.origin = 'Local Search'
            configs_acq.append((acq_val, configuration))
        self.last_stats = stats
        self.rng.shuffle(configs_acq)                       # random tie-break
        configs_acq.sort(reverse=True, key=lambda x: x[0])
        return configs_acq

    def _get_initial_points(self, num_points, runhistory):
        if runhistory.empty():
            init_points = self.config_space.sample_configuration(size=num_points)
        else:
            configs_previous_runs = runhistory.get_all_configs()
            configs_previous_runs_sorted = self._sort_configs_by_acq_value(configs_previous_runs)
            num_configs_local_search = int(min(len(configs_previous_runs_sorted), num_points))
            init_points = list(map(lambda x: x[1], configs_previous_runs_sorted[:num_configs_local_search]))
        return init_points

    def _one_iter(self, start_point, acq_val_incumbent, stats):
        """ei_optimization.py:239-294 with the neighbourhood of each step scored in one launch."""
        incumbent = start_point
        # native driver: the whole climb in one call (same launches, accept rule and stream consumption)
        model = getattr(self.acquisition_function, 'model', None)
        space = getattr(start_point, 'space', None)
        if self.native_climb and hasattr(model, 'climb') and hasattr(space, 'const_mask'):
            af = self.acquisition_function
            af._require_eta()
            res = model.climb(start_point.get_array(), float(acq_val_incumbent[0]), self.rng, af.eta, af.par,
                              af._var_floor(), self.max_steps, space.const_mask, NUM_NEIGHBORS, NEIGHBOUR_STDEV, stats)
            if res is not None:
                x, acq = res
                if not np.array_equal(x, start_point.get_array()):
                    incumbent = Configuration(start_point.space, x, -1)
                return np.array([acq]), incumbent
        local_search_steps = 0
        while True:
            local_search_steps += 1
            if local_search_steps % 1000 == 0:
                self.logger.warning('Local search took already %d iterations. Is it maybe stuck in a infinite loop?',
                                    local_search_steps)
            changed_inc = False
            neighbours = one_exchange_neighbourhood_arrays(incumbent, seed=self.rng.randint(0, 10000))
            if len(neighbours):
                acq = self.acquisition_function(neighbours)
                stats['launches'] += 1
                better = np.nonzero(acq[:, 0] > acq_val_incumbent[0])[0]
                # the reference evaluates lazily and stops at the first improvement
                visited = int(better[0]) + 1 if len(better) else len(neighbours)
                stats['neighbours'] += visited
                # ConfigSpace's lazy generator draws one np.random.random() from the GLOBAL legacy stream per
                # neighbour it produces ("only rigorously check every tenth configuration", util.pyx); the
                # reference's loop produces exactly the `visited` neighbours it inspects.  Same stream position
                # as the facades' bootstrap draws expect.
                np.random.random(visited)
                if len(better):
                    j = int(better[0])
                    incumbent = Configuration(incumbent.space, neighbours[j], -1)
                    acq_val_incumbent = acq[j]
                    changed_inc = True
            stats['steps'] += 1
            if (not changed_inc) or (self.max_steps is not None and local_search_steps == self.max_steps):
                break
        return acq_val_incumbent, incumbent


class RandomSearch(_SortMixin, AcquisitionFunctionMaximizer):
    def _maximize(self, runhistory, num_points, _sorted=False, **kwargs):
        if num_points > 1:
            if getattr(self.config_space, 'sample_batch', None) is not None:
                # same draws, rows of one array instead of `num_points` objects
                rand_configs = self.config_space.sample_batch(num_points)
                rand_configs.origin = 'Random Search (sorted)' if _sorted else 'Random Search'
                if _sorted:
                    return self._sort_configs_by_acq_value(rand_configs)
                return [(0, c) for c in rand_configs]
            rand_configs = self.config_space.sample_configuration(size=num_points)
        else:
            rand_configs = [self.config_space.sample_configuration(size=1)]
        if _sorted:
            for c in rand_configs:
                c.origin = 'Random Search (sorted)'
            return self._sort_configs_by_acq_value(rand_configs)
        for c in rand_configs:
            c.origin = 'Random Search'
        return [(0, c) for c in rand_configs]


class InterleavedLocalAndRandomSearch(AcquisitionFunctionMaximizer):
    def __init__(self, acquisition_function, config_space, rng=None, max_steps=None, n_steps_plateau_walk=10,
                 n_sls_iterations=10, rand_prob=0.1):
        super().__init__(acquisition_function, config_space, rng)
        # the reference hands the SAME rng object to all three helpers (ei_optimization.py:391-404)
        self.random_search = RandomSearch(acquisition_function=acquisition_function, config_space=config_space,
                                          rng=rng)
        self.local_search = LocalSearch(acquisition_function=acquisition_function, config_space=config_space,
                                        rng=rng, max_steps=max_steps, n_steps_plateau_walk=n_steps_plateau_walk)
        self.n_sls_iterations = n_sls_iterations
        self.random_chooser = ChooserProb(prob=rand_prob, rng=rng if rng is not None else self.rng)

    def maximize(self, runhistory, num_points, random_configuration_chooser=None, **kwargs):
        if random_configuration_chooser is None:
            random_configuration_chooser = self.random_chooser
        next_configs_by_local_search = self.local_search._maximize(runhistory, self.n_sls_iterations, **kwargs)
        next_configs_by_random_search_sorted = self.random_search._maximize(
            runhistory, num_points - len(next_configs_by_local_search), _sorted=True)
        if isinstance(next_configs_by_random_search_sorted, _ScoredBatch):
            next_configs_by_acq_value = _MergedChallengers(next_configs_by_random_search_sorted,
                                                           next_configs_by_local_search)
        else:
            next_configs_by_acq_value = next_configs_by_random_search_sorted + next_configs_by_local_search
            next_configs_by_acq_value.sort(reverse=True, key=lambda x: x[0])
            next_configs_by_acq_value = [_[1] for _ in next_configs_by_acq_value]
        challengers = ChallengerList(next_configs_by_acq_value, self.config_space, random_configuration_chooser)
        random_configuration_chooser.next_smbo_iteration()
        return challengers

    def _maximize(self, runhistory, num_points, **kwargs):
        raise NotImplementedError()


class ChallengerList(object):
    """ei_optimization.py:478-524: interleaves freshly sampled random configurations by the
    chooser's rule while iterating over the sorted challengers."""

    def __init__(self, challengers, configuration_space, random_configuration_chooser):
        self.challengers = challengers
        self.configuration_space = configuration_space
        self._index = 0
        self._iteration = 1
        self.random_configuration_chooser = random_configuration_chooser

    def __iter__(self):
        return self

    def __next__(self):
        if self._index == len(self.challengers):
            raise StopIteration
        if self.random_configuration_chooser.check(self._iteration):
            config = self.configuration_space.sample_configuration()
            config.origin = 'Random Search'
        else:
            config = self.challengers[self._index]
            self._index += 1
        self._iteration += 1
        return config


class HistoryContainer(object):
    """The two methods of tlbo/utils/history_container.py the maximisers use."""

    def __init__(self):
        self.data = dict()

    def add(self, config, perf):
        self.data[config] = perf

    def empty(self):
        return len(self.data) == 0

    def get_all_configs(self):
        return list(self.data.keys())
