"""Offline (tabular) SMBO loop -- the CALLER of the hot path, kept as host Python
(reference tlbo/framework/smbo_offline.py:18-264).  Restated minimally so the drop-in can
be driven end to end; the numerics all happen in the facade / acquisition / maximiser."""
import numpy as np

from tlbo_b200.acquisition_function.acquisition import EI
from tlbo_b200.config_space import Configuration, convert_configurations_to_array
from tlbo_b200.optimizer.ei_offline_optimizer import OfflineSearch
from tlbo_b200.optimizer.random_configuration_chooser import ChooserProb
from tlbo_b200.utils.constants import FAILDED, MAXINT, SUCCESS
from tlbo_b200.utils.normalization import zero_mean_unit_var_normalization, zero_one_normalization


class SMBO_OFFLINE(object):
    def __init__(self, target_hpo_data, config_space, surrogate_model, acq_func='ei', source_hpo_data=None,
                 enable_init_design=False, num_src_hpo_trial=50, surrogate_type='gp', max_runs=200,
                 logging_dir='./logs', initial_runs=3, task_id=None, random_seed=None, shard=None):
        if random_seed is None:
            random_seed = np.random.RandomState(1).randint(MAXINT)
        if enable_init_design:
            raise ValueError('the clustering-based initial design is outside the B200 hot path')
        if acq_func != 'ei':
            raise ValueError('invalid acquisition function ~ %s.' % acq_func)
        self.config_space = config_space
        self.task_id = task_id
        self.random_seed = random_seed
        self.rng = np.random.RandomState(self.random_seed)
        self.source_hpo_data = source_hpo_data
        self.num_src_hpo_trial = num_src_hpo_trial
        self.surrogate_type = surrogate_type
        self.acq_func = acq_func
        self.max_iterations = max_runs
        self.iteration_id = 0
        self.default_obj_value = MAXINT

        # target table: (X, y) arrays or the reference's OrderedDict{Configuration -> perf}
        if isinstance(target_hpo_data, (tuple, list)):
            self.X_table = np.ascontiguousarray(target_hpo_data[0], dtype=np.float64)
            ys = np.asarray(target_hpo_data[1], dtype=np.float64)
            self.configuration_list = None
        else:
            self.configuration_list = list(target_hpo_data.keys())
            self.X_table = convert_configurations_to_array(self.configuration_list)
            ys = np.array(list(target_hpo_data.values()))
        self.test_ys = None
        if ys.ndim == 2:
            assert ys.shape[1] == 2
            self.test_ys = ys[:, 1]
            ys = ys[:, 0]
        self.ys = ys
        self.y_max, self.y_min = np.max(self.ys), np.min(self.ys)

        self.configurations = list()        # evaluated ROW INDICES, evaluation order
        self._default_row_cache = None
        self.failed_configurations = list()
        self.perfs = list()
        self.test_perfs = list()
        self.init_num = initial_runs
        self.initial_configurations = None

        if hasattr(self.config_space, 'seed'):
            self.config_space.seed(self.random_seed)
        np.random.seed(self.random_seed)                    # smbo_offline.py:69 -- global stream
        self.model = surrogate_model
        self.acquisition_function = EI(self.model)
        table = self.X_table if self.configuration_list is None else self.configuration_list
        self.acq_optimizer = OfflineSearch(table, self.acquisition_function, config_space,
                                           rng=np.random.RandomState(self.random_seed), shard=shard,
                                           prefetch_keys=True)
        self.random_configuration_chooser = ChooserProb(prob=0.001, rng=np.random.RandomState(self.random_seed))
        self.incumbent_value = None
        self.timing = dict(train=0.0, maximize=0.0)

    # ---- metric -------------------------------------------------------------------------
    def get_adtm(self):
        y_inc = self.get_inc_y()
        assert self.y_max != self.y_min
        return (y_inc - self.y_min) / (self.y_max - self.y_min)

    def get_inc_y(self):
        return np.min(self.perfs)

    def run(self):
        while self.iteration_id < self.max_iterations:
            self.iterate()

    def config_of(self, row):
        if self.configuration_list is not None:
            return self.configuration_list[row]
        return Configuration(self.config_space, self.X_table[row], row)

    def evaluate(self, row):
        test_perf = -1 if self.test_ys is None else self.test_ys[row]
        return self.ys[row], test_perf

    # ---- loop -----------------------------------------------------------------------------
    def iterate(self):
        if len(self.configurations) == 0:
            X = np.zeros((0, self.X_table.shape[1]))
        else:
            X = self.X_table[self.configurations]
        Y = np.array(self.perfs, dtype=np.float64)
        row = self.choose_next(X, Y)
        trial_state, trial_info = SUCCESS, None
        if row not in self.configurations and row not in self.failed_configurations:
            perf, test_perf = self.evaluate(row)
            if perf == MAXINT:
                trial_info = 'failed configuration evaluation.'
                trial_state = FAILDED
            if trial_state == SUCCESS and perf < MAXINT:
                if len(self.configurations) == 0:
                    self.default_obj_value = perf
                self.configurations.append(int(row))
                self.perfs.append(perf)
                self.test_perfs.append(test_perf)
            else:
                self.failed_configurations.append(int(row))
        else:
            if row in self.configurations:
                trial_state, perf = SUCCESS, self.perfs[self.configurations.index(row)]
            else:
                trial_state, perf = FAILDED, MAXINT
        self.iteration_id += 1
        return self.config_of(row), trial_state, perf, trial_info

    def sample_random_config(self, config_num=1):
        rows = list()
        sample_cnt = 0
        taken = lambda r: r in self.configurations or r in self.failed_configurations or r in rows
        while len(rows) < config_num:
            sample_cnt += 1
            _idx = int(self.rng.randint(self.X_table.shape[0]))
            if not taken(_idx):
                rows.append(_idx)
                sample_cnt = 0
            else:
                sample_cnt += 1
            if sample_cnt >= 200:
                rows.append(_idx)
                sample_cnt = 0
        return rows

    def _default_row(self):
        """Row of ``config_space.get_default_configuration()`` in the table (first match, as
        ``default_config not in self.configuration_list`` / list membership compares by value), row 0 when
        the space has no default or the default is not a table row."""
        if self._default_row_cache is None:
            row = 0
            get = getattr(self.config_space, 'get_default_configuration', None)
            default = get() if get is not None else None
            if default is not None:
                vec = np.asarray(default.get_array(), dtype=np.float64).ravel()
                if vec.shape[0] == self.X_table.shape[1]:
                    T = self.X_table
                    same = ((T == vec[None, :]) | (np.isnan(T) & np.isnan(vec)[None, :])).all(axis=1)
                    hits = np.nonzero(same)[0]
                    if len(hits):
                        row = int(hits[0])
            self._default_row_cache = row
        return self._default_row_cache

    def choose_next(self, X, Y):
        _config_num = X.shape[0]
        if _config_num < self.init_num:
            # smbo_offline.py:210-219: the space's default configuration if it is a table row, else the
            # table's first row; a random unevaluated row once that one has been evaluated
            default_row = self._default_row()
            if default_row not in self.configurations and default_row not in self.failed_configurations:
                return default_row
            return self.sample_random_config()[0]

        if self.random_configuration_chooser.check(self.iteration_id):
            row = self.sample_random_config()[0]
            if len(self.model.target_weight) == 0:
                self.model.target_weight.append(0.)
            else:
                self.model.target_weight.append(self.model.target_weight[-1])
            return row

        # tie-break keys / evaluated-row list are staged while the facade's fit kernel runs
        evaluated = self.configurations + self.failed_configurations
        self.model._overlap_hook = lambda: self.acq_optimizer.prepare(evaluated)
        try:
            self.model.train(X, Y)
        finally:
            self.model._overlap_hook = None
        if self.model.method_id in ['tst', 'tstm']:
            y_, _, _ = zero_one_normalization(Y)
        else:
            y_, _, _ = zero_mean_unit_var_normalization(Y)
        incumbent_value = np.min(y_)
        self.incumbent_value = incumbent_value
        self.acquisition_function.update(model=self.model, eta=incumbent_value, num_data=len(self.configurations))
        return self.acq_optimizer.best_unevaluated(evaluated)
