"""Synthetic benchmark histories shaped like the reference's Google-Drive pickles
(``OrderedDict{Configuration -> 1 - balanced_accuracy}``, <= 20k configs per task,
30 tasks; reference README.md:17-57, test/cls/gen_data_cls.py:137-140).  No network,
so the tables are generated: SURVEY.md section 8(d)."""
import numpy as np

# random_forest space after ConfigSpace's name ordering (tlbo/config_space/space_instance.py:11-35,
# tlbo/model/util_funcs.py:14-55): bootstrap, criterion (categorical, 2 choices), max_depth (const),
# max_features (float), max_leaf_nodes, min_impurity_decrease (const), min_samples_leaf,
# min_samples_split (int), min_weight_fraction_leaf (const).
RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0], dtype=np.int64)
RF_CONST = np.array([0, 0, 1, 0, 1, 1, 0, 0, 1], dtype=bool)

# tools/utils.py:8-9 (first entries of the reference's fixed seed table)
SEEDS = [4465, 3822, 4531, 8459, 6295, 2854, 7820, 4050, 280, 6983]


def space_types(kind='rf', d=None):
    if kind == 'rf':
        return RF_TYPES.copy(), RF_CONST.copy()
    return np.zeros(int(d), dtype=np.int64), np.zeros(int(d), dtype=bool)


def sample_table(rng, T, types, const_mask, int_levels=19):
    """T configurations in the reference's array encoding
    (tlbo/config_space/util.py:11-30): floats / ints on the unit interval, categoricals as
    choice indices, constants as 0."""
    d = len(types)
    X = np.zeros((T, d), dtype=np.float64)
    for j in range(d):
        if types[j] > 0:
            X[:, j] = rng.randint(0, int(types[j]), size=T)
        elif const_mask[j]:
            X[:, j] = 0.0
        else:
            X[:, j] = rng.rand(T)
    return X


class TaskFamily(object):
    """A family of correlated smooth objectives y_t(x) in [0, 1] (random Fourier features
    with a shared and a task-specific weight part, plus categorical offsets)."""

    def __init__(self, types, const_mask, n_tasks=30, n_feat=64, corr=0.7, seed=0, noise=0.01):
        rng = np.random.RandomState(seed)
        self.types = np.asarray(types)
        self.const_mask = np.asarray(const_mask)
        self.cont = np.nonzero((self.types == 0) & (~self.const_mask))[0]
        self.cat = np.nonzero(self.types > 0)[0]
        self.n_tasks = n_tasks
        self.noise = noise
        dc = max(len(self.cont), 1)
        self.W = rng.randn(dc, n_feat) * 2.5
        self.b = rng.rand(n_feat) * 2 * np.pi
        shared = rng.randn(n_feat)
        self.coef = np.sqrt(corr) * shared[None, :] + np.sqrt(1 - corr) * rng.randn(n_tasks, n_feat)
        self.cat_off = rng.randn(n_tasks, max(len(self.cat), 1), 8) * 0.3
        self.rng = rng

    def evaluate(self, task, X, rng=None):
        Xc = X[:, self.cont] if len(self.cont) else np.zeros((X.shape[0], 1))
        phi = np.cos(Xc.dot(self.W) + self.b) * np.sqrt(2.0 / self.W.shape[1])
        f = phi.dot(self.coef[task])
        for k, j in enumerate(self.cat):
            f = f + self.cat_off[task, k, X[:, j].astype(int)]
        if rng is not None and self.noise > 0:
            f = f + rng.randn(X.shape[0]) * self.noise
        return 1.0 / (1.0 + np.exp(-1.5 * f))       # squash to (0, 1) like an error rate


def make_benchmark(n_tasks=30, T=20000, kind='rf', d=None, seed=4465, n_src_rows=50, pool=None):
    """Returns dict(types, tables=[(X_t, y_t)] per task).  ``pool`` (if given) resizes the
    candidate table of every task to that many rows."""
    types, const_mask = space_types(kind, d)
    rng = np.random.RandomState(seed)
    fam = TaskFamily(types, const_mask, n_tasks=n_tasks, seed=seed + 1)
    tables = []
    rows = int(pool) if pool else int(T)
    for t in range(n_tasks):
        X = sample_table(rng, rows, types, const_mask)
        y = fam.evaluate(t, X, rng)
        tables.append((X, y))
    return dict(types=types, const_mask=const_mask, tables=tables, family=fam)


def leave_one_out_sources(tables, target, n_source, seed):
    """Source selection of tools/offline_benchmark.py:278-284: all other tasks, shuffled by
    a RandomState(seed), first ``n_source``."""
    ids = [i for i in range(len(tables)) if i != target]
    rng = np.random.RandomState(seed)
    rng.shuffle(ids)
    ids = ids[:n_source]
    return [tables[i] for i in ids], ids
