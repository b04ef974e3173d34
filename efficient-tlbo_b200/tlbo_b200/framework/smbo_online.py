"""Online SMBO loop -- the CALLER of the hot path in the online scenario
(reference tlbo/framework/online/smbo_online.py:11-62 on top of
tlbo/framework/smbo_offline.py:147-264; tools/online_benchmark.py:269-305).

Configurations are free points of the search space (not rows of a table), the objective is
a callable, and the acquisition maximiser is ``InterleavedLocalAndRandomSearch``.  The
reference's ``SMBO_ONLINE.evaluate`` returns one value where ``iterate`` unpacks two
(smbo_online.py:28-39 vs smbo_offline.py:162): the evident intent -- ``perf =
objective_function(config)`` -- is what is implemented here."""
import numpy as np

from tlbo_b200.acquisition_function.acquisition import EI
from tlbo_b200.config_space import convert_configurations_to_array
from tlbo_b200.optimizer.ei_optimization import HistoryContainer, InterleavedLocalAndRandomSearch
from tlbo_b200.optimizer.random_configuration_chooser import ChooserProb
from tlbo_b200.utils.constants import FAILDED, MAXINT, SUCCESS
from tlbo_b200.utils.normalization import zero_mean_unit_var_normalization, zero_one_normalization


class SMBO_ONLINE(object):
    def __init__(self, objective_function, config_space, surrogate_model, acq_func='ei', max_runs=200,
                 initial_runs=3, random_seed=None, num_points=5000):
        if random_seed is None:
            random_seed = np.random.RandomState(1).randint(MAXINT)
        if acq_func != 'ei':
            raise ValueError('invalid acquisition function ~ %s.' % acq_func)
        self.objective_function = objective_function
        self.config_space = config_space
        self.random_seed = random_seed
        self.rng = np.random.RandomState(self.random_seed)
        self.max_iterations = max_runs
        self.iteration_id = 0
        self.default_obj_value = MAXINT
        self.init_num = initial_runs
        self.num_points = num_points
        self.configurations = list()
        self.failed_configurations = list()
        self.perfs = list()
        self.history_container = HistoryContainer()
        self.config_space.seed(self.random_seed)
        np.random.seed(self.random_seed)                    # smbo_offline.py:69 -- global stream
        self.model = surrogate_model
        self.acquisition_function = EI(self.model)
        self.acq_optimizer = InterleavedLocalAndRandomSearch(          # smbo_online.py:18-26
            self.acquisition_function, self.config_space, rng=np.random.RandomState(self.random_seed),
            max_steps=None, n_steps_plateau_walk=10, n_sls_iterations=10, rand_prob=0.0)
        self.random_configuration_chooser = ChooserProb(prob=0.001, rng=np.random.RandomState(self.random_seed))
        self.incumbent_value = None

    def get_inc_y(self):
        return np.min(self.perfs)

    def run(self):
        while self.iteration_id < self.max_iterations:
            self.iterate()

    def evaluate(self, config):
        try:
            return float(self.objective_function(config))
        except Exception:
            return MAXINT

    def iterate(self):
        if len(self.configurations) == 0:
            X = np.zeros((0, len(self.config_space)))
        else:
            X = convert_configurations_to_array(self.configurations)
        Y = np.array(self.perfs, dtype=np.float64)
        config = self.choose_next(X, Y)
        trial_state, trial_info = SUCCESS, None
        if config not in (self.configurations + self.failed_configurations):
            perf = self.evaluate(config)
            if perf == MAXINT:
                trial_info = 'failed configuration evaluation.'
                trial_state = FAILDED
            if trial_state == SUCCESS and perf < MAXINT:
                if len(self.configurations) == 0:
                    self.default_obj_value = perf
                self.configurations.append(config)
                self.perfs.append(perf)
                self.history_container.add(config, perf)
            else:
                self.failed_configurations.append(config)
        else:
            if config in self.configurations:
                trial_state, perf = SUCCESS, self.perfs[self.configurations.index(config)]
            else:
                trial_state, perf = FAILDED, MAXINT
        self.iteration_id += 1
        return config, trial_state, perf, trial_info

    def sample_random_config(self, config_num=1):
        """smbo_online.py:41-55."""
        configs = list()
        sample_cnt = 0
        while len(configs) < config_num:
            sample_cnt += 1
            config = self.config_space.sample_configuration(1)
            if config not in (self.configurations + self.failed_configurations + configs):
                configs.append(config)
                sample_cnt = 0
            else:
                sample_cnt += 1
            if sample_cnt >= 200:
                configs.append(config)
                sample_cnt = 0
        return configs

    def choose_next(self, X, Y):
        _config_num = X.shape[0]
        if _config_num < self.init_num:
            default_config = self.config_space.get_default_configuration()
            if default_config is not None and default_config not in (self.configurations +
                                                                     self.failed_configurations):
                return default_config
            return self.sample_random_config()[0]
        if self.random_configuration_chooser.check(self.iteration_id):
            config = self.sample_random_config()[0]
            if len(self.model.target_weight) == 0:
                self.model.target_weight.append(0.)
            else:
                self.model.target_weight.append(self.model.target_weight[-1])
            return config
        self.model.train(X, Y)
        if self.model.method_id in ['tst', 'tstm']:
            y_, _, _ = zero_one_normalization(Y)
        else:
            y_, _, _ = zero_mean_unit_var_normalization(Y)
        incumbent_value = np.min(y_)
        self.incumbent_value = incumbent_value
        self.acquisition_function.update(model=self.model, eta=incumbent_value,
                                         num_data=len(self.history_container.data))
        sorted_configs = self.acq_optimizer.maximize(runhistory=self.history_container, num_points=self.num_points)
        for _config in sorted_configs:
            if _config not in (self.configurations + self.failed_configurations):
                return _config
        raise ValueError('The configuration in the SET is over')
