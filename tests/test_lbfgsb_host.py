"""The device optimiser's single-thread statement (csrc/lbfgsb_dense.cuh), compiled for the
host, against scipy.optimize.fmin_l_bfgs_b -- the routine the reference calls
(tlbo/model/gp.py:244).  Same start, bounds and defaults (m=10, factr=1e7, pgtol=1e-5)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.optimize

from oracle import gp as ogp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, 'tests', 'hostsim', 'lbfgsb_host.cpp')
LIB = os.path.join(ROOT, 'tests', 'hostsim', 'liblbfgsb_host.so')
HDR = os.path.join(ROOT, 'efficient-tlbo_b200', 'csrc', 'lbfgsb_dense.cuh')

FN = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double))


@pytest.fixture(scope='module')
def hostlib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-x', 'c++', '-I',
                               os.path.dirname(HDR), SRC, '-o', LIB])
    lib = ctypes.CDLL(LIB)
    lib.lbfgsb_host_minimize.restype = ctypes.c_int
    return lib


def _minimize(lib, fun, x0, lo, hi):
    n = len(x0)
    evals = []

    def cb(xp, gp):
        x = np.array([xp[i] for i in range(n)])
        f, g = fun(x)
        for i in range(n):
            gp[i] = g[i]
        evals.append((x, f))
        return float(f)
    cfn = FN(cb)
    xo = (ctypes.c_double * n)()
    fo = ctypes.c_double()
    nfev, nit, status = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    arr = lambda a: (ctypes.c_double * n)(*[float(v) for v in a])
    lib.lbfgsb_host_minimize(n, arr(x0), arr(lo), arr(hi), cfn, xo, ctypes.byref(fo), ctypes.byref(nfev),
                             ctypes.byref(nit), ctypes.byref(status), None, 0)
    return np.array(list(xo)), fo.value, nfev.value, nit.value, status.value, evals


def _rosen(x):
    return scipy.optimize.rosen(x), scipy.optimize.rosen_der(x)


@pytest.mark.parametrize('n,seed', [(2, 0), (5, 1), (11, 2), (20, 3)])
def test_rosenbrock_with_active_bounds(hostlib, n, seed):
    rng = np.random.RandomState(seed)
    x0 = rng.uniform(-1.5, 1.5, size=n)
    lo = np.full(n, -1.0)
    hi = np.full(n, 0.8)          # optimum (1, .., 1) is outside: bounds become active
    xs, fs, info = scipy.optimize.fmin_l_bfgs_b(_rosen, x0, bounds=list(zip(lo, hi)))
    x, f, nfev, nit, status, evals = _minimize(hostlib, _rosen, x0, lo, hi)
    assert abs(f - fs) <= 1e-9 * max(1.0, abs(fs))
    np.testing.assert_allclose(x, xs, rtol=1e-6, atol=1e-7)
    assert abs(nfev - info['funcalls']) <= 2 and abs(nit - info['nit']) <= 2


@pytest.mark.parametrize('n,types', [(20, [2, 2, 0, 0, 0, 0, 0, 0, 0]), (40, [0, 0, 0, 0, 0])])
def test_gp_nll_restarts_track_scipy(hostlib, n, types):
    """The 11 restarts of GaussianProcess._optimize on the oracle's NLL: same optimum per
    restart (identical evaluation counts for most of them)."""
    types = np.array(types)
    rng = np.random.RandomState(n)
    spec = ogp.KernelSpec(types)
    X = rng.rand(n, len(types))
    for j, t in enumerate(types):
        if t > 0:
            X[:, j] = rng.randint(0, t, size=n)
    y = np.sin(3 * X[:, 2]) + X[:, 3] ** 2 + 0.05 * rng.randn(n)
    y = (y - y.mean()) / y.std()
    fun = lambda t: ogp.nll(t, spec, X, y)
    lb = spec.log_bounds()
    p0 = np.vstack([spec.default_theta()[None, :], ogp.restart_points(spec, 4465)])
    same_count = 0
    for start in p0:
        xs, fs, info = scipy.optimize.fmin_l_bfgs_b(fun, start, bounds=[(a, b) for a, b in lb])
        x, f, nfev, nit, status, evals = _minimize(hostlib, fun, start, lb[:, 0], lb[:, 1])
        assert abs(f - fs) <= 1e-6 * max(1.0, abs(fs)), (f, fs)
        same_count += int(nfev == info['funcalls'])
    assert same_count >= len(p0) - 3
