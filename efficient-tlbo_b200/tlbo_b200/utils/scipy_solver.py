"""Simplex-constrained ranking-loss weight solver (reference tlbo/utils/scipy_solver.py:81-115,
``loss_type=3``).  scipy's SLSQP keeps control on the host exactly as in the reference; every
objective / gradient callback is ONE launch of the pairwise-logistic CUDA kernel with the
prediction matrix resident on the device."""
import numpy as np
from scipy.optimize import minimize

try:                                     # scipy >= 1.16: the SLSQP core behind ``minimize`` (reverse communication)
    from scipy.optimize._slsqplib import slsqp as _slsqp_core
    from scipy.linalg.lapack import HAS_ILP64 as _ILP64
except Exception:                        # pragma: no cover -- other scipy layouts go through ``minimize`` itself
    _slsqp_core = None


def slsqp_simplex(fg, m, ftol=1e-8, maxiter=100):
    """``minimize(f, [1/m]*m, method='SLSQP', jac=f', constraints=[sum(x) == 1, x >= 0], options={'ftol': ftol})``
    driven directly: the SAME core routine (``scipy.optimize._slsqplib.slsqp``) fed the same numbers in the same
    order as ``scipy.optimize._slsqp_py._minimize_slsqp`` feeds it, without ``minimize``'s per-call set-up
    (constraint triage, ``ScalarFunction``) and per-iteration wrappers (constraint lambdas, Jacobian stacking) --
    about 1 ms of interpreter time per TOPO iteration.  The constraint values are formed as the reference's lambdas
    form them (``sum(x) - 1`` is Python's left-to-right sum, not numpy's pairwise one); the constraint normals are
    constant.  ``fg(x) -> (f, g)``.  Returns ``(x, success)``; iterates, multipliers and exit mode are bit-identical
    to ``minimize``'s (tests/test_host_logic.py::test_slsqp_simplex_equals_scipy_minimize)."""
    n, meq = int(m), 1
    mtot = n + 1
    x = np.array([1. / n] * n)
    xl = np.full(n, np.nan)
    xu = np.full(n, np.nan)
    state = {"acc": ftol, "alpha": 0.0, "f0": 0.0, "gs": 0.0, "h1": 0.0, "h2": 0.0, "h3": 0.0, "h4": 0.0, "t": 0.0,
             "t0": 0.0, "tol": 10.0 * ftol, "exact": 0, "inconsistent": 0, "reset": 0, "iter": 0,
             "itermax": int(maxiter), "line": 0, "m": mtot, "meq": meq, "mode": 0, "n": n}
    indices = np.zeros([max(mtot + 2 * n + 2, 1)], dtype=np.int64 if _ILP64 else np.int32)
    buffer_size = (n * (n + 1) // 2 + 3 * mtot * n - (mtot + 5 * n + 7) * meq + 9 * mtot + 8 * n * n + 35 * n
                   + meq * meq + 28)
    buffer = np.zeros(max(buffer_size, 1), dtype=np.float64)
    fx, g = fg(x)
    mult = np.zeros([max(1, mtot + 2 * n + 2)], dtype=np.float64)
    C = np.zeros([mtot, n], dtype=np.float64, order='F')
    C[0, :] = 1.0
    C[1:, :] = np.eye(n)
    d = np.zeros([mtot], dtype=np.float64)

    def constraints():
        d[0] = sum(x) - 1
        d[1:] = x
    constraints()
    while True:
        _slsqp_core(state, fx, g, C, d, x, mult, xl, xu, buffer, indices)
        mode = state['mode']
        if mode == 1:
            fx = fg(x)[0]
            constraints()
        elif mode == -1:
            g = fg(x)[1]
        else:
            break
    return x, mode == 0


def scipy_solve(A, b, loss_type=3, debug=False):
    if loss_type != 3:
        raise ValueError('only the pairwise logistic ranking loss (loss_type=3) is on the B200 hot path')
    from tlbo_b200 import ops
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    n, m = A.shape
    dev_loss = ops.PairLogistic(A, b)

    # same constraint functions as the reference; the constant Jacobians are built once
    eye_m, ones_m = np.eye(m), np.array([1.] * m)
    ineq_cons = {'type': 'ineq', 'fun': lambda x: np.array(x), 'jac': lambda x: eye_m}
    eq_cons = {'type': 'eq', 'fun': lambda x: np.array([sum(x) - 1]), 'jac': lambda x: ones_m}
    x0 = np.array([1. / m] * m)

    if _slsqp_core is not None:
        def fg(x):
            r = dev_loss(x)
            return r[0], r[2]
        x, success = slsqp_simplex(fg, m)
        status = False if np.isnan(x).any() else True
        if not success and status:
            x[x < 0.] = 0.
            x[x > 1.] = 1.
            if sum(x) > 1.5:
                status = False
        return x, status

    def f(x):
        return dev_loss(x)[0]

    def f_der(x):
        return dev_loss(x)[2]

    res = minimize(f, x0, method='SLSQP', jac=f_der, constraints=[eq_cons, ineq_cons],
                   options={'ftol': 1e-8, 'disp': False})
    status = False if np.isnan(res.x).any() else True
    if not res.success and status:
        res.x[res.x < 0.] = 0.
        res.x[res.x > 1.] = 1.
        if sum(res.x) > 1.5:
            status = False
    return res.x, status
