"""Simplex-constrained ranking-loss weight solver (reference tlbo/utils/scipy_solver.py:81-115,
``loss_type=3``).  scipy's SLSQP keeps control on the host exactly as in the reference; every
objective / gradient callback is ONE launch of the pairwise-logistic CUDA kernel with the
prediction matrix resident on the device."""
import numpy as np
from scipy.optimize import minimize


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
