"""Oracle: EI, candidate ordering, and the offline BO loop that drives the path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Citations are relative
to ``/root/reference``.
"""
import numpy as np
from scipy.stats import norm

from oracle.facade import zero_mean_unit_var_normalization, zero_one_normalization


class OracleEI(object):
    """``EI`` (tlbo/acquisition_function/acquisition.py:112-177) with the
    array-level ``__call__`` of ``AbstractAcquisitionFunction`` (:55-77)."""

    def __init__(self, model, par=0.0):
        self.model = model
        self.par = par
        self.eta = None

    def update(self, **kwargs):
        for key in kwargs:
            setattr(self, key, kwargs[key])

    def __call__(self, X):
        if len(X.shape) == 1:
            X = X[np.newaxis, :]
        acq = self._compute(X)
        if np.any(np.isnan(acq)):
            idx = np.where(np.isnan(acq))[0]
            acq[idx, :] = -np.finfo(np.float64).max
        return acq

    def _compute(self, X):
        if len(X.shape) == 1:
            X = X[:, np.newaxis]
        m, v = self.model.predict_marginalized_over_instances(X)
        s = np.sqrt(v)
        if self.eta is None:
            raise ValueError('No current best specified. Call update('
                             'eta=<int>) to inform the acquisition function '
                             'about the current best value.')

        def calculate_f():
            z = (self.eta - m - self.par) / s
            return (self.eta - m - self.par) * norm.cdf(z) + s * norm.pdf(z)

        if np.any(s == 0.0):
            s_copy = np.copy(s)
            s[s_copy == 0.0] = 1.0
            f = calculate_f()
            f[s_copy == 0.0] = 0.0
        else:
            f = calculate_f()
        if (f < 0).any():
            raise ValueError("Expected Improvement is smaller than 0 for at least one sample.")
        return f


def sort_by_acq_value(acq_values, rng):
    """``_sort_configs_by_acq_value`` (tlbo/optimizer/ei_optimization.py:106-132):
    consumes one ``rng.rand(N)``; returns candidate indices, best first."""
    random = rng.rand(len(acq_values))
    indices = np.lexsort((random.flatten(), acq_values.flatten()))
    return indices[::-1], random


def first_unevaluated(order, evaluated):
    """tlbo/framework/smbo_offline.py:261-263 over row indices."""
    ev = set(int(e) for e in evaluated)
    for ind in order:
        if int(ind) not in ev:
            return int(ind)
    raise ValueError('The configuration in the SET (%d) is over' % len(order))


class OracleSMBOOffline(object):
    """``SMBO_OFFLINE`` (tlbo/framework/smbo_offline.py) over a candidate table
    ``X_table float64[N, d]`` with objective values ``y_table float64[N]``.

    ``default`` (a configuration vector) plays ``config_space.get_default_configuration()``: the first
    evaluation is its table row when it is one, else row 0 ("default config not in the list ->
    configuration_list[0]", :212-215).
    """

    def __init__(self, X_table, y_table, surrogate_model, random_seed, initial_runs=3, max_runs=200, default=None):
        self.X_table = np.asarray(X_table, dtype=np.float64)
        self.y_table = np.asarray(y_table, dtype=np.float64)
        self.random_seed = random_seed
        self.rng = np.random.RandomState(random_seed)
        self.max_iterations = max_runs
        self.iteration_id = 0
        self.init_num = initial_runs
        self.evaluated = list()     # row indices, evaluation order
        self.perfs = list()
        np.random.seed(random_seed)                                    # :69
        self.model = surrogate_model
        self.acquisition_function = OracleEI(self.model)
        self.acq_rng = np.random.RandomState(random_seed)              # :80-84
        self.chooser_rng = np.random.RandomState(random_seed)          # :85-88
        self.y_max, self.y_min = np.max(self.y_table), np.min(self.y_table)
        self.last_acq = None
        self.last_keys = None
        self.default_row = 0
        if default is not None:
            hits = np.nonzero((self.X_table == np.asarray(default, dtype=np.float64)[None, :]).all(axis=1))[0]
            if len(hits):
                self.default_row = int(hits[0])

    def get_adtm(self):
        return (np.min(self.perfs) - self.y_min) / (self.y_max - self.y_min)

    def sample_random_config(self):
        """:191-206 with config_num=1."""
        sample_cnt = 0
        while True:
            sample_cnt += 1
            idx = self.rng.randint(len(self.X_table))
            if idx not in self.evaluated:
                return idx
            sample_cnt += 1
            if sample_cnt >= 200:
                return idx

    def choose_next(self, X, Y):
        """:208-264."""
        if X.shape[0] < self.init_num:
            if self.default_row not in self.evaluated:
                return self.default_row
            return self.sample_random_config()
        if self.chooser_rng.rand() < 0.001:                            # ChooserProb.check
            idx = self.sample_random_config()
            tw = self.model.target_weight
            tw.append(0. if len(tw) == 0 else tw[-1])
            return idx
        self.model.train(X, Y)
        if self.model.method_id in ['tst', 'tstm']:
            y_, _, _ = zero_one_normalization(Y)
        else:
            y_, _, _ = zero_mean_unit_var_normalization(Y)
        incumbent_value = np.min(y_)
        self.acquisition_function.update(model=self.model, eta=incumbent_value, num_data=len(self.evaluated))
        acq = self.acquisition_function(self.X_table)
        self.last_acq = acq
        order, self.last_keys = sort_by_acq_value(acq, self.acq_rng)
        return first_unevaluated(order, self.evaluated)

    def iterate(self):
        """:147-189 (no failed configurations in a tabular benchmark)."""
        if len(self.evaluated) == 0:
            X = np.zeros((0, self.X_table.shape[1]))
        else:
            X = self.X_table[self.evaluated]
        Y = np.array(self.perfs, dtype=np.float64)
        idx = self.choose_next(X, Y)
        if idx not in self.evaluated:
            self.evaluated.append(idx)
            self.perfs.append(float(self.y_table[idx]))
        self.iteration_id += 1
        return idx
