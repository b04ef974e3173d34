"""CPU tests of the online-scenario maximisers (SURVEY 8 f rank 1): the product's BATCHED local
search (whole neighbourhood per acquisition call) against the oracle's literal restatement of
tlbo/optimizer/ei_optimization.py (one neighbour per call, early break), with a deterministic
stand-in acquisition so that no GPU is needed: same incumbents, same challengers, same
consumption of every random stream, same count of lazily inspected neighbours."""
import numpy as np
import pytest

from oracle import ei_optimization as oeo

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])
RF_CONST = np.array([0, 0, 1, 0, 1, 1, 0, 0, 1], dtype=bool)
SPACES = [(RF_TYPES, RF_CONST), (np.zeros(5, dtype=int), None), (np.array([3, 4, 2, 0]), None),
          (np.array([0, 0, 5]), np.array([1, 0, 0], dtype=bool))]


def _spaces(types, const):
    from tlbo_b200.config_space import TableSpace
    return TableSpace(types, const_mask=const), oeo.OracleSpace(types, const)


class FakeAcq(object):
    """Deterministic multi-modal acquisition with the reference call interface."""

    def __init__(self, types):
        rng = np.random.RandomState(5)
        self.c = rng.rand(len(types))
        self.w = rng.rand(len(types)) + 0.5
        self.calls, self.rows = 0, 0

    def __call__(self, configs):
        from tlbo_b200.config_space import convert_configurations_to_array
        X = convert_configurations_to_array(configs)
        if X.ndim == 1:
            X = X[np.newaxis, :]
        self.calls += 1
        self.rows += X.shape[0]
        f = np.exp(-np.sum(self.w * (X - self.c) ** 2, axis=1)) + 0.2 * np.cos(7 * X.sum(axis=1))
        return f.reshape(-1, 1)


@pytest.mark.parametrize('types,const', SPACES)
def test_sampling_and_neighbourhood_match_the_oracle(types, const):
    from tlbo_b200.config_space import (get_one_exchange_neighbourhood, one_exchange_neighbourhood_arrays,
                                        one_exchange_neighbourhood_arrays_py)
    sp, osp = _spaces(types, const)
    sp.seed(11)
    osp.seed(11)
    X = sp.sample_array(40)
    assert np.array_equal(X, osp.sample(40))
    cfgs = sp.sample_configuration(size=5)
    assert np.array_equal(np.array([c.get_array() for c in cfgs]), osp.sample(5))
    one = sp.sample_configuration()
    assert np.array_equal(one.get_array(), osp.sample(1)[0])
    n_expected = sum((int(t) - 1) if t > 0 else (0 if (const is not None and const[j]) else 8)   # num_neighbors=8: tlbo/config_space/__init__.py:10
                     for j, t in enumerate(types))
    for i, c in enumerate(cfgs):
        got = one_exchange_neighbourhood_arrays(c, seed=100 + i)
        want = list(osp.neighbourhood(c.get_array(), seed=100 + i))
        assert got.shape == (n_expected, len(types))
        assert np.array_equal(got, one_exchange_neighbourhood_arrays_py(c, seed=100 + i))    # native == Python
        assert np.array_equal(got, np.array(want).reshape(-1, len(types)))
        assert np.all((got != c.get_array()).sum(axis=1) == 1)          # one-exchange
        gen = list(get_one_exchange_neighbourhood(c, seed=100 + i))
        assert len(gen) == n_expected and np.array_equal(gen[0].get_array(), got[0])


def test_neighbourhood_of_a_space_without_free_columns_is_empty():
    from tlbo_b200.config_space import TableSpace, one_exchange_neighbourhood_arrays
    sp = TableSpace([0, 0], const_mask=[True, True])
    sp.seed(0)
    c = sp.sample_configuration()
    assert one_exchange_neighbourhood_arrays(c, 1).shape == (0, 2)
    assert list(oeo.OracleSpace([0, 0], [True, True]).neighbourhood(c.get_array(), 1)) == []


@pytest.mark.parametrize('types,const', SPACES[:3])
@pytest.mark.parametrize('n_prev', [0, 4, 25])
def test_batched_local_search_equals_the_sequential_reference(types, const, n_prev):
    from tlbo_b200.config_space import Configuration
    from tlbo_b200.optimizer.ei_optimization import HistoryContainer, InterleavedLocalAndRandomSearch
    sp, osp = _spaces(types, const)
    sp.seed(3)
    osp.seed(3)
    prev = sp.sample_array(n_prev) if n_prev else np.zeros((0, len(types)))
    if n_prev:
        osp.sample(n_prev)
    hist = HistoryContainer()
    for i in range(n_prev):
        hist.add(Configuration(sp, prev[i], -1), float(i))
    acq_p, acq_o = FakeAcq(types), FakeAcq(types)
    rng_p, rng_o = np.random.RandomState(9), np.random.RandomState(9)
    opt = InterleavedLocalAndRandomSearch(acq_p, sp, rng=rng_p, rand_prob=0.3)
    counters = dict(steps=0, neighbours=0)
    chall_o, sorted_o = oeo.interleaved_maximize(acq_o, osp, rng_o, list(prev), 300, rand_prob=0.3,
                                                 counters=counters)
    chall_p = opt.maximize(hist, 300)
    got = [c.get_array() for c in chall_p]
    want = list(chall_o)
    assert len(got) == len(want) >= 300
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    # identical consumption of the shared rng and of the space's sampler
    assert np.array_equal(rng_p.get_state()[1], rng_o.get_state()[1]) and rng_p.get_state()[2] == rng_o.get_state()[2]
    assert np.array_equal(sp._rng.get_state()[1], osp.rng.get_state()[1])
    st = opt.local_search.last_stats
    assert st['steps'] == counters['steps'] and st['neighbours'] == counters['neighbours']
    # the point of the restructuring: one acquisition call per hill-climb step instead of one per neighbour
    assert acq_p.calls <= st['steps'] + 3
    assert acq_o.calls >= counters['neighbours']
    assert st['steps'] > (10 if n_prev != 4 else 4)


def test_challenger_list_interleaves_random_configurations():
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.optimizer.ei_optimization import ChallengerList
    from tlbo_b200.optimizer.random_configuration_chooser import ChooserProb
    sp = TableSpace(np.zeros(3, dtype=int))
    sp.seed(1)
    base = sp.sample_configuration(size=20)
    rng = np.random.RandomState(2)
    out = list(ChallengerList(list(base), sp, ChooserProb(prob=0.5, rng=rng)))
    kept = [c for c in out if any(c is b for b in base)]
    assert kept == list(base) and len(out) > 20
    assert all(c.origin == 'Random Search' for c in out if not any(c is b for b in base))


def test_native_neighbourhood_reproduces_numpys_legacy_stream():
    """Many seeds / spaces: the C helper (own MT19937, masked-rejection randint / shuffle, polar normal)
    must equal the numpy-driven Python statement bit for bit, including values at the borders of the
    unit interval (many rejected normal draws) and wide categoricals (long shuffles)."""
    from tlbo_b200.config_space import (Configuration, TableSpace, one_exchange_neighbourhood_arrays,
                                        one_exchange_neighbourhood_arrays_py)
    rng = np.random.RandomState(0)
    spaces = [TableSpace(RF_TYPES, const_mask=RF_CONST), TableSpace(np.array([7, 0, 0, 13, 2, 0])),
              TableSpace(np.zeros(12, dtype=int)), TableSpace(np.array([40, 3]))]
    for sp in spaces:
        sp.seed(4)
        X = sp.sample_array(25)
        cont = np.nonzero((sp.types == 0) & ~sp.const_mask)[0]
        if len(cont):
            X[:5, cont[0]] = [0.0, 1.0, 1e-9, 0.999999, 0.5]
        for i in range(25):
            c = Configuration(sp, X[i], -1)
            for seed in (0, 1, int(rng.randint(0, 10000)), 9999, 2 ** 31 + 5):
                a = one_exchange_neighbourhood_arrays(c, seed)
                b = one_exchange_neighbourhood_arrays_py(c, seed)
                assert a.shape == b.shape and np.array_equal(a, b), (sp.types, i, seed)
    # inactive (non-finite) columns are never exchanged
    sp = TableSpace(np.array([0, 0, 3]))
    c = Configuration(sp, np.array([0.3, np.nan, 1.0]), -1)
    a, b = one_exchange_neighbourhood_arrays(c, 5), one_exchange_neighbourhood_arrays_py(c, 5)
    assert a.shape == (10, 3) and np.array_equal(a, b, equal_nan=True)   # 8 numerical + 2 categorical


class TiedAcq(FakeAcq):
    """Coarsely quantised values: many exact ties between random-search rows and local-search results."""

    def __call__(self, configs):
        return np.round(FakeAcq.__call__(self, configs), 1)


@pytest.mark.parametrize('acq_cls', [FakeAcq, TiedAcq])
def test_lazy_challenger_batch_equals_the_eager_list_path(acq_cls):
    """The sorted random search over a ConfigurationBatch + array merge (no per-row objects) yields the challengers
    of the reference's list path -- sort by lexsort((rand, acq)), concatenate, STABLE descending sort
    (ei_optimization.py:106-132,446-456) -- element for element, ties included, with the same stream consumption."""
    from tlbo_b200.config_space import Configuration, TableSpace
    from tlbo_b200.optimizer.ei_optimization import (HistoryContainer, InterleavedLocalAndRandomSearch,
                                                     _MergedChallengers)
    from tlbo_b200.optimizer.random_configuration_chooser import ChooserProb

    class EagerSpace(TableSpace):
        sample_batch = None                                   # hasattr() is False -> the list path

    out, states = [], []
    for cls in (TableSpace, EagerSpace):
        sp = cls(RF_TYPES, const_mask=RF_CONST)
        if cls is EagerSpace:
            assert not hasattr(sp, 'sample_batch') or sp.sample_batch is None
        sp.seed(3)
        hist = HistoryContainer()
        X0 = sp.sample_array(12)
        for i in range(12):
            hist.add(Configuration(sp, X0[i], -1), float(i))
        rng = np.random.RandomState(9)
        opt = InterleavedLocalAndRandomSearch(acq_cls(RF_TYPES), sp, rng=rng, n_sls_iterations=10, rand_prob=0.0)
        opt.local_search.native_climb = False
        np.random.seed(1)
        ch = opt.maximize(hist, 600, random_configuration_chooser=ChooserProb(prob=0.0, rng=np.random.RandomState(1)))
        if cls is TableSpace:
            assert isinstance(ch.challengers, _MergedChallengers)
        else:
            assert isinstance(ch.challengers, list)
        out.append(list(ch))
        states.append((rng.get_state()[1].copy(), rng.get_state()[2], sp._rng.get_state()[1].copy()))
    a, b = out
    assert len(a) == len(b) == 600
    assert all(x == y and x.origin == y.origin for x, y in zip(a, b))
    assert np.array_equal(states[0][0], states[1][0]) and states[0][1] == states[1][1]
    assert np.array_equal(states[0][2], states[1][2])
