"""Host-side logic of the drop-in, checked on the CPU (no compute calls)."""
import numpy as np
import pytest

from oracle import gp as ogp
from oracle.facade import kfold_indices


def test_start_points_equal_the_oracle_restatement():
    """GaussianProcess._start_points (product) consumes the shared / own RNG streams exactly
    like the reference's _optimize (gp.py:211-239), restated by oracle.gp.restart_points."""
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.model.model_builder import build_model
    for types, seed in (([2, 2, 0, 0, 0, 0, 0, 0, 0], 4465), ([0] * 5, 3822), ([3, 0, 2], 7)):
        m = build_model('gp', TableSpace(types), np.random.RandomState(seed))
        p0 = m._start_points()
        spec = ogp.KernelSpec(np.array(types))
        want = np.vstack([spec.default_theta()[None, :], ogp.restart_points(spec, seed)])
        assert p0.shape == want.shape == (11, len(types) + 2)
        assert (p0 == want).all()
        np.testing.assert_array_equal(m.kernel.bounds, spec.log_bounds())


def test_kfold_slices_match_sklearn():
    from sklearn.model_selection import KFold
    from tlbo_b200.facade.topo_variant1 import _kfold_val_slices
    for n in (5, 7, 10, 23, 75):
        sl = _kfold_val_slices(n, 5)
        for (lo, hi), (tr, va), (otr, ova) in zip(sl, KFold(n_splits=5).split(np.zeros((n, 1))), kfold_indices(n)):
            assert list(range(lo, hi)) == list(va) == list(ova)
            assert list(tr) == list(otr)


def test_merge_best_is_the_reference_ordering():
    from tlbo_b200.distributed import merge_best
    rng = np.random.RandomState(0)
    for trial in range(50):
        N = 40
        ei = np.round(rng.rand(N), 1)            # many ties
        key = np.round(rng.rand(N), 1)
        order = np.lexsort((key, ei))[::-1]      # ties on (ei, key) resolve to the larger row
        cuts = sorted(rng.choice(np.arange(1, N), size=3, replace=False))
        quads = []
        for lo, hi in zip([0] + cuts, cuts + [N]):
            best = max(range(lo, hi), key=lambda r: (ei[r], key[r], r))
            quads.append((ei[best], key[best], best, 0))
        quads.append((-np.inf, -np.inf, -1, 2))  # an empty shard
        got = merge_best(quads)
        assert got[2] == order[0] and got[3] == 2


def test_configuration_objects():
    from tlbo_b200.config_space import Configuration, TableSpace, convert_configurations_to_array, get_types, \
        table_to_configurations
    cs = TableSpace([2, 0, 0])
    types, bounds = get_types(cs)
    assert list(types) == [2, 0, 0] and bounds[0][0] == 2
    X = np.array([[0, .1, .2], [1, .3, .4], [0, .1, .2]])
    cfgs = table_to_configurations(cs, X)
    assert cfgs[0] == cfgs[2] and cfgs[0] != cfgs[1] and len({cfgs[0], cfgs[1], cfgs[2]}) == 2
    np.testing.assert_array_equal(convert_configurations_to_array(cfgs), X)
    assert isinstance(cfgs[1], Configuration) and cfgs[1].index == 1


def test_unsupported_surrogates_raise():
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.model.model_builder import build_model
    cs = TableSpace([0, 0])
    for kind in ('rf', 'gp_rbf'):
        with pytest.raises(ValueError):
            build_model(kind, cs, np.random.RandomState(0))
    m = build_model('gp_mcmc', cs, np.random.RandomState(0))       # model_builder.py:74-89
    assert (m.n_mcmc_walkers, m.burnin_steps, m.chain_length) == (12, 250, 250)
    with pytest.raises(ValueError):
        build_model('nonsense', cs, np.random.RandomState(0))
    with pytest.raises(ValueError):
        RGPE(cs, [], None, 1, surrogate_type='rf')


def test_synthetic_benchmark_shape():
    from tlbo_b200 import synthetic
    b = synthetic.make_benchmark(n_tasks=4, T=200, seed=1)
    assert len(b['tables']) == 4 and b['tables'][0][0].shape == (200, 9)
    X, y = b['tables'][2]
    assert set(np.unique(X[:, 0])) <= {0.0, 1.0} and (X[:, [2, 4, 5, 8]] == 0).all()
    assert (y > 0).all() and (y < 1).all()
    src, ids = synthetic.leave_one_out_sources(b['tables'], 1, 2, 5)
    assert len(src) == 2 and 1 not in ids


# ---- native host half of the per-iteration training sets (tlbo_host_prepare_jobs) ---------------------
def _py_jobs(X, y, ranges, outer, normalize, normalize_y):
    from tlbo_b200.utils.normalization import zero_mean_unit_var_normalization, zero_one_normalization
    """numpy statement: what rgpe/_prepare_single_surrogate/fit_models_async do"""
    y = y.copy(); outs = []
    n = len(y)
    for (lo, hi), og in zip(ranges, outer):
        rows = [i for i in range(n) if not (lo <= i < hi)]
        if og:
            if (y[rows] == y[rows[0]]).all():
                y[rows[0]] += 1e-4
        if hi == lo:
            yj = y          # the caller's array itself
            Xj = X
        else:
            yj = y[rows]; Xj = X[rows, :]
        if normalize == 1:
            if (yj == yj[0]).all(): yj[0] += 1e-4
            yj, _, _ = zero_mean_unit_var_normalization(yj)
        elif normalize == 2:
            if (yj == yj[0]).all(): yj[0] += 1e-4
            yj, _, _ = zero_one_normalization(yj)
        Xi = np.array(Xj, dtype=np.float64); Xi[~np.isfinite(Xi)] = -1
        yy = np.asarray(yj, dtype=np.float64)
        if normalize_y:
            m = float(np.mean(yy)); s = float(np.std(yy))
            if s == 0: s = 1
            yy = (yy - m) / s
        else:
            m, s = 0.0, 1.0
        outs.append((Xi, yy.flatten(), m, s))
    return y, outs


def _c_jobs(X, y, ranges, outer, normalize, normalize_y):
    from tlbo_b200._cabi import check, lib, ptr
    y = y.copy(); n, d = X.shape; B = len(ranges); ncap = ((n + 7)//8)*8
    lo = np.array([r[0] for r in ranges], dtype=np.int32); hi = np.array([r[1] for r in ranges], dtype=np.int32)
    og = np.array(outer, dtype=np.int32)
    Xo = np.full((B, ncap, d), 7.0); yo = np.full((B, ncap), 7.0); no = np.zeros(B, np.int32); ym = np.zeros(B); ys = np.zeros(B)
    check(lib.tlbo_host_prepare_jobs(ptr(np.ascontiguousarray(X)), ptr(y), n, d, B, ptr(lo), ptr(hi), ptr(og), normalize, normalize_y, ncap, ptr(Xo), ptr(yo), ptr(no), ptr(ym), ptr(ys)))
    return y, [(Xo[b, :no[b]], yo[b, :no[b]], ym[b], ys[b]) for b in range(B)], Xo, yo, no



def test_native_job_preparation_is_bit_identical_to_numpy():
    """``tlbo_host_prepare_jobs`` against the numpy statement of build_single_surrogate + GaussianProcess._train's
    host half (base_facade.py:93-108, rgpe.py:55-70, gp.py:115-122): guards (including the in-place mutation
    of the caller's y), both normalisations with numpy's pairwise-summation mean / std, NaN imputation,
    zero padding -- bit for bit, over RGPE folds and sklearn KFold ranges, n up to 210."""
    rng = np.random.RandomState(0)
    for trial in range(700):
        n = int(rng.randint(5, 210)); d = int(rng.randint(1, 12))
        X = rng.rand(n, d)
        if trial % 7 == 0:
            X[rng.rand(n, d) < 0.1] = np.nan
        kind = trial % 5
        if kind == 0:
            y = rng.randn(n) * 10 ** rng.uniform(-4, 4)
        elif kind == 1:
            y = np.full(n, rng.randn())                      # all equal
        elif kind == 2:
            y = np.round(rng.rand(n), 1)                     # many ties
        elif kind == 3:
            y = np.full(n, 0.5); y[:n // 5] = rng.rand(n // 5)   # every fold but the first leaves all-equal rows
        else:
            y = rng.rand(n) + 1e6
        if trial % 2 == 0:      # rgpe.py folds: n // 5 rows each, the last takes the remainder
            f = n // 5
            ranges = [(0, 0)] + [(i * f, n if i == 4 else (i + 1) * f) for i in range(5)]
            outer = [0, 1, 1, 1, 1, 1]
        else:                   # sklearn KFold(5)
            sizes = np.full(5, n // 5); sizes[:n % 5] += 1
            cur = 0; ranges = [(0, 0)]
            for s in sizes:
                ranges.append((cur, cur + int(s))); cur += int(s)
            outer = [0] * 6
        normalize = [1, 1, 2, 0][trial % 4]
        y1, o1 = _py_jobs(X, y, ranges, outer, normalize, 1)
        y2, o2, Xo, yo, no = _c_jobs(X, y, ranges, outer, normalize, 1)
        assert np.array_equal(y1, y2), trial
        for (a, b, m, s), (a2, b2, m2, s2) in zip(o1, o2):
            assert np.array_equal(a, a2) and np.array_equal(b, b2, equal_nan=True), trial
            assert (m == m2 or (m != m and m2 != m2)) and (s == s2 or (s != s and s2 != s2)), trial
        for b in range(len(ranges)):
            assert (Xo[b, no[b]:] == 0).all() and (yo[b, no[b]:] == 0).all()


def test_prepare_jobs_wrapper_layout():
    from tlbo_b200 import ops
    rng = np.random.RandomState(3)
    X, y = rng.rand(23, 4), rng.rand(23)
    prep = ops.prepare_jobs(X, y, [(0, 0), (0, 4), (20, 23)], [0, 1, 1], 'standardize')
    assert (prep.B, prep.ncap, prep.d) == (3, 24, 4) and list(prep.n) == [23, 19, 20]
    nx, ny = 3 * 24 * 4, 3 * 24
    assert list(prep.buf[nx + ny:].view(np.int32)[:3]) == [23, 19, 20]
    yj = prep.buf[nx:nx + ny].reshape(3, 24)
    assert abs(yj[0, :23].mean()) < 1e-12 and abs(yj[0, :23].std() - 1) < 1e-12
    assert np.array_equal(prep.buf[:nx].reshape(3, 24, 4)[1, :19], X[4:])


def test_thread_local_global_stream_equals_numpys():
    """``utils/global_stream.thread_stream``: two threads drawing through the module-level functions, interleaved,
    each see the stream a lone process would after ``np.random.seed``; threads outside keep numpy's own stream."""
    import threading
    from tlbo_b200.utils.global_stream import thread_stream
    want = {}
    for s in (5, 6):
        np.random.seed(s)
        want[s] = (np.random.standard_normal((3, 4)), np.random.random(5), np.random.standard_normal(2))
    np.random.seed(99)
    ref_outside = np.random.random(3)
    got = {}
    bar = threading.Barrier(2)

    def work(s):
        with thread_stream():
            np.random.seed(s)
            bar.wait()
            a = np.random.standard_normal((3, 4))
            bar.wait()
            b = np.random.random(5)
            bar.wait()
            c = np.random.standard_normal(2)
            got[s] = (a, b, c)
    np.random.seed(99)
    ts = [threading.Thread(target=work, args=(s,)) for s in (5, 6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for s in (5, 6):
        for x, y in zip(got[s], want[s]):
            assert np.array_equal(x, y)
    assert np.array_equal(np.random.random(3), ref_outside)        # the process-global stream was not touched
    assert np.random.seed.__module__.startswith('numpy') or np.random.seed.__name__ == 'seed'


def test_host_turns_are_exclusive_and_given_up_while_waiting():
    """``thread_stream(turns=True)``: only one trial thread runs its host segment at a time; inside ``waiting()`` (and
    inside a large ``standard_normal`` draw) the turn is free for the others; draws stay the private stream's."""
    import threading
    import time
    from tlbo_b200.utils.global_stream import thread_stream, waiting
    inside = [0]
    overlap = [0]
    met = threading.Barrier(2)
    got = {}

    def work(s):
        with thread_stream(turns=True):
            np.random.seed(s)
            for _ in range(20):
                inside[0] += 1
                overlap[0] = max(overlap[0], inside[0])
                time.sleep(0.0005)
                inside[0] -= 1
                with waiting():
                    time.sleep(0.0005)
            with waiting():
                met.wait(timeout=20)                 # both threads are inside their blocks here: no deadlock
            got[s] = np.random.standard_normal((70, 70))
    ts = [threading.Thread(target=work, args=(s,)) for s in (11, 12)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(30)
    assert not any(t.is_alive() for t in ts)
    assert overlap[0] == 1
    for s in (11, 12):
        assert np.array_equal(got[s], np.random.RandomState(s).standard_normal((70, 70)))
    with waiting():                                   # a thread without a turn: no-op
        pass


def test_slsqp_simplex_equals_scipy_minimize():
    """``utils/scipy_solver.slsqp_simplex`` (the SLSQP core driven directly) returns bit for bit what
    ``scipy.optimize.minimize`` returns for the reference's problem set-up (tlbo/utils/scipy_solver.py:81-115):
    pairwise-logistic ranking losses of several shapes, including one that ends on an exit mode other than 0."""
    from scipy.optimize import minimize
    from tlbo_b200.utils import scipy_solver as ss
    if ss._slsqp_core is None:
        pytest.skip('scipy without _slsqplib: scipy_solve goes through minimize itself')
    rng = np.random.RandomState(3)
    for n, m, maxiter in [(12, 3, 100), (30, 8, 100), (45, 30, 100), (45, 30, 3), (20, 1, 100)]:
        A = rng.standard_normal((n, m))
        b = rng.standard_normal(n)
        iu = np.triu_indices(n, 1)
        sgn = np.sign(b[iu[0]] - b[iu[1]])
        D = A[iu[0]] - A[iu[1]]
        calls = []

        def fg(x):
            z = -sgn * (D @ x)
            calls.append(np.array(x))
            return float(np.sum(np.logaddexp(0., z))), D.T @ (-sgn / (1. + np.exp(-z)))
        x1, ok1 = ss.slsqp_simplex(fg, m, maxiter=maxiter)
        n_lean = len(calls)
        eye_m, ones_m = np.eye(m), np.array([1.] * m)
        res = minimize(lambda x: fg(x)[0], np.array([1. / m] * m), method='SLSQP', jac=lambda x: fg(x)[1],
                       constraints=[{'type': 'eq', 'fun': lambda x: np.array([sum(x) - 1]), 'jac': lambda x: ones_m},
                                    {'type': 'ineq', 'fun': lambda x: np.array(x), 'jac': lambda x: eye_m}],
                       options={'ftol': 1e-8, 'disp': False, 'maxiter': maxiter})
        assert np.array_equal(x1, res.x), (n, m, x1, res.x)
        assert ok1 == res.success
        assert n_lean <= len(calls) - n_lean          # never more evaluations than minimize makes
