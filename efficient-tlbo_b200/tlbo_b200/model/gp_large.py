"""Large-n Gaussian-process path (SCoT's single GP on K x 50 + n ~ 1.5 k stacked rows).

The register / shared-memory panel kernels of csrc/gp_fit.cu keep ONE surrogate per CTA and stop at a few hundred
rows.  Above ``LARGE_N`` rows a fit evaluates the reference's negative log marginal likelihood + gradient
(tlbo/model/gp.py:164-197 over sklearn's ``log_marginal_likelihood``; kernels tlbo/model/gp_kernels.py) with
device-wide FP64 operations -- the kernel matrix and its derivative terms as batched element-wise tensors, the
Cholesky factorisation / inverse through cuSOLVER and the contractions through cuBLAS, all via ``torch`` on the
current stream -- driven by scipy's L-BFGS-B on the host exactly as the reference drives it (11 restarts,
gp.py:199-248).  The fitted state lands in the same ``SurrogatePack`` slot layout the posterior kernels read, and
the pool posterior is the two-step form of the reference: cross-covariance block ``K(X*, X)`` (element-wise,
HBM-bound), then ``L^-1 K*^T`` as a dense FP64 GEMM (cuBLAS) and a row reduction -- chunked over the pool.

This path is LIBRARY code (cuBLAS / cuSOLVER), not a hand-written kernel: it exists so that SCoT runs end to end
on the device; a blocked multi-CTA factorisation of our own is the stated next step."""
import math

import numpy as np

LARGE_N = 256                 # rows above which a fit / pool posterior takes this path
HORSESHOE_SCALE = 0.1
VERY_SMALL_NUMBER = 1e-10


def _split_columns(types):
    types = np.asarray(types)
    cont = [j for j in range(len(types)) if types[j] == 0]
    cat = [j for j in range(len(types)) if types[j] != 0]
    return cont, cat


class _Problem(object):
    """Device tensors of one training set, columns split into continuous / categorical."""

    def __init__(self, X, y, types):
        import torch
        cont, cat = _split_columns(types)
        self.cont, self.cat = cont, cat
        dev = torch.device('cuda')
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.n = X.shape[0]
        self.Xc = torch.from_numpy(np.ascontiguousarray(X[:, cont])).to(dev)
        self.Xk = torch.from_numpy(np.ascontiguousarray(X[:, cat])).to(dev)
        self.y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(dev)
        # theta-independent pair tensors: squared differences per continuous column, mismatch per categorical
        self.D2 = (self.Xc[:, None, :] - self.Xc[None, :, :]) ** 2 if cont else None           # [n, n, dc]
        self.NE = (self.Xk[:, None, :] != self.Xk[None, :, :]).to(torch.float64) if cat else None
        self.eye = torch.eye(self.n, dtype=torch.float64, device=dev)
        # activity mask of gp_kernels.py:21-29 (spaces with conditions): 1 iff the two rows have the same set of
        # columns <= -1
        self.active = None
        if getattr(types, 'has_conditions', False):
            Xd = torch.from_numpy(X).to(dev) <= -1.0
            self.active = (~((Xd[:, None, :] != Xd[None, :, :]).any(dim=2))).to(torch.float64)


def _priors(theta):
    """log-priors of GaussianProcess._nll and their gradients (lognormal on c: gp_base_prior.py:389-449;
    horseshoe(0.1) on the noise: :289-355); same operation order as csrc/gp_nll.cuh:nll_finish."""
    v0 = math.exp(theta[0])
    if v0 <= 0.0:
        lp0, gp0 = -1e25, 0.0
    else:
        lv = math.log(v0)
        lp0 = -(lv * lv) / 2.0 - math.log(2.5066282746310002 * v0)
        gp0 = -(1.0 + lv) / v0 * v0
    vn = math.exp(theta[-1])
    s2 = HORSESHOE_SCALE * HORSESHOE_SCALE
    if vn == 0.0:
        lpn, gpn = math.inf, math.inf
    else:
        a = math.log(1.0 + 3.0 * (s2 / (vn * vn)))
        lpn = math.log(a + VERY_SMALL_NUMBER)
        b = 3.0 * s2 + vn * vn
        b *= math.log(3.0 * s2 * (1.0 / (vn * vn)) + 1.0)
        b = max(b, 1e-14)
        gpn = -(6.0 * s2) / b
    return lp0 + lpn, gp0, gpn


def kernel_parts(P, theta):
    """(Km, t, E) with Km = c * matern * hamming (no noise term), t = sqrt(5) r, E = exp(-t - hs)."""
    import torch
    dc, dk = len(P.cont), len(P.cat)
    c = math.exp(theta[0])
    n = P.n
    if dc:
        il2 = torch.tensor(np.exp(-2.0 * np.asarray(theta[1:1 + dc])), dtype=torch.float64, device=P.y.device)
        s = (P.D2 * il2).sum(-1)
        t = torch.sqrt(5.0 * s)
    else:
        t = torch.zeros(n, n, dtype=torch.float64, device=P.y.device)
    if dk:
        hk = torch.tensor(0.5 * np.exp(-2.0 * np.asarray(theta[1 + dc:1 + dc + dk])), dtype=torch.float64,
                          device=P.y.device)
        hs = (P.NE * hk).sum(-1)
    else:
        hs = torch.zeros(n, n, dtype=torch.float64, device=P.y.device)
    E = torch.exp(-t - hs)
    if P.active is not None:
        E = E * P.active                   # every term of K and dK carries E
    Km = c * ((1.0 + t + t * t / 3.0) * E)
    return Km, t, E


def nll_grad(P, theta):
    """GaussianProcess._nll (gp.py:164-197): (f, g) as host floats; (1e25, 0) on a failed Cholesky or a
    non-finite value, as the reference."""
    import torch
    theta = np.asarray(theta, dtype=np.float64)
    H = theta.shape[0]
    dc, dk = len(P.cont), len(P.cat)
    c, noise = math.exp(theta[0]), math.exp(theta[-1])
    Km, t, E = kernel_parts(P, theta)
    K = Km + noise * P.eye
    L, info = torch.linalg.cholesky_ex(K)
    if int(info.item()) != 0:
        return 1e25, np.zeros(H)
    alpha = torch.cholesky_solve(P.y[:, None], L)
    lml = -0.5 * (P.y * alpha[:, 0]).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * P.n * math.log(2.0 * math.pi)
    W = alpha @ alpha.T - torch.cholesky_inverse(L)
    g = torch.zeros(H, dtype=torch.float64, device=P.y.device)
    g[0] = 0.5 * (W * Km).sum()
    if dc:
        # gp_kernels.py:468-496: dK/dlog l_p = c * hamming * 5/3 * D_p * (1 + t) * exp(-t), D_p = (dx_p / l_p)^2
        il2 = torch.tensor(np.exp(-2.0 * theta[1:1 + dc]), dtype=torch.float64, device=P.y.device)
        WC = W * (c * (5.0 / 3.0) * (t + 1.0) * E)
        g[1:1 + dc] = 0.5 * torch.einsum('ij,ijp->p', WC, P.D2) * il2
    if dk:
        # gp_kernels.py:723-736: dK/dlog l_q = K * [x_q != x'_q] / l_q^3 (the reference's scaling)
        il3 = torch.tensor(np.exp(-3.0 * theta[1 + dc:1 + dc + dk]), dtype=torch.float64, device=P.y.device)
        g[1 + dc:1 + dc + dk] = 0.5 * torch.einsum('ij,ijq->q', W * Km, P.NE) * il3
    g[H - 1] = 0.5 * noise * torch.diagonal(W).sum()
    out = torch.cat([lml[None], g]).cpu().numpy()
    lp, gp0, gpn = _priors(theta)
    f = float(out[0]) + lp
    grad = out[1:].copy()
    grad[0] += gp0
    grad[-1] += gpn
    if not (np.isfinite(f) and np.all(np.isfinite(grad))):
        return 1e25, np.zeros(H)
    return -f, -grad


def fit_large(mdl, X, y, p0, do_optimize=True):
    """gp.py:98-151 + :199-248 for one surrogate: noise-bump probe at the default theta, the L-BFGS-B restarts
    (scipy, bounds = the kernel's log bounds, first strict improvement wins), final factorisation into the
    model's pack slot.  ``y`` is already GP-level normalised.  Returns per-restart (theta, f, nfev)."""
    import torch
    from scipy import optimize
    P = _Problem(X, y, mdl.types)
    p0 = np.array(p0, dtype=np.float64)
    theta0 = p0[0].copy()
    bumps = 0
    for _ in range(10):                       # gp.py:127-140
        Km, _, _ = kernel_parts(P, theta0)
        _, info = torch.linalg.cholesky_ex(Km + math.exp(theta0[-1]) * P.eye)
        if int(info.item()) == 0:
            break
        theta0[-1] = math.log(math.exp(theta0[-1]) + 1.0)
        bumps += 1
    p0[0] = theta0
    results = []
    if do_optimize:
        bounds = [(float(lo), float(hi)) for lo, hi in mdl.kernel.bounds]
        theta_star, f_star = None, np.inf
        for p in p0:
            th, f, d = optimize.fmin_l_bfgs_b(lambda v: nll_grad(P, v), p, bounds=bounds)
            results.append((np.array(th), float(f), int(d['funcalls'])))
            if f < f_star:
                f_star, theta_star = f, np.array(th)
        if theta_star is None:
            raise RuntimeError('GP fit failed: no finite optimum (status 2)')
    else:
        theta_star = theta0
    factorize_into_pack(mdl, P, theta_star)
    return theta_star, results, bumps


def factorize_into_pack(mdl, P, theta):
    """The final ``GaussianProcessRegressor.fit`` at theta* (gp.py:146): L, alpha, L^-1 into the pack slot in the
    layout of csrc/gp_fit.cu:gp_factorize_kernel (columns permuted continuous-first, continuous columns
    pre-divided by their length scale)."""
    import torch
    pack, slot = mdl._pack, mdl._slot
    n, dc, dk = P.n, len(P.cont), len(P.cat)
    d = dc + dk
    c, noise = math.exp(theta[0]), math.exp(theta[-1])
    Km, _, _ = kernel_parts(P, theta)
    L, info = torch.linalg.cholesky_ex(Km + noise * P.eye)
    if int(info.item()) != 0:
        raise np.linalg.LinAlgError('Cholesky of the kernel matrix failed at the selected hyper-parameters')
    alpha = torch.cholesky_solve(P.y[:, None], L)[:, 0]
    Linv = torch.linalg.solve_triangular(L, P.eye, upper=False)
    ls_c = np.exp(np.asarray(theta[1:1 + dc]))
    ls_k = np.exp(np.asarray(theta[1 + dc:1 + dc + dk]))
    hyp = np.zeros(pack.hyp.shape[1])
    hyp[0], hyp[1], hyp[2], hyp[3], hyp[4] = c, noise, mdl.mean_y_, mdl.std_y_, c + noise
    hyp[8:8 + dc] = 1.0 / ls_c
    hyp[8 + dc:8 + d] = 1.0 / (2.0 * ls_k * ls_k)
    pack.hyp[slot].copy_(torch.from_numpy(hyp))
    Xs = pack.Xs[slot]
    Xs.zero_()
    if dc:
        Xs[:n, :dc] = P.Xc * torch.from_numpy(1.0 / ls_c).to(P.Xc.device)
    if dk:
        Xs[:n, dc:d] = P.Xk
    pack.alpha[slot].zero_()
    pack.alpha[slot, :n] = alpha
    pack.Linv[slot].zero_()
    pack.Linv[slot, :n, :n] = Linv
    pack.theta[slot].copy_(torch.from_numpy(np.asarray(theta, dtype=np.float64)))
    pack.n[slot] = n


def posterior_pool(pack, slot, types, X_dev, chunk=8192):
    """Posterior mean / variance ``[1, N]`` of the large surrogate in ``pack`` slot ``slot`` over the device pool
    ``X_dev [N, d]`` (gp.py:250-305: un-standardised, GP-level 1e-10 clip): per chunk the cross-covariance block,
    the mean ``K* alpha`` and ``|L^-1 k*|^2`` as one FP64 GEMM against the resident inverse factor."""
    import torch
    cont, cat = _split_columns(types)
    dc, dk = len(cont), len(cat)
    n = int(pack.n[slot].item())
    hyp = pack.hyp[slot]
    c, ymean, ystd, diag = float(hyp[0]), float(hyp[2]), float(hyp[3]), float(hyp[4])
    Xs = pack.Xs[slot, :n]
    alpha = pack.alpha[slot, :n]
    LinvT = pack.Linv[slot, :n, :n].T.contiguous()
    il = hyp[8:8 + dc]
    hk = hyp[8 + dc:8 + dc + dk]
    N = int(X_dev.shape[0])
    mu = torch.empty(1, N, dtype=torch.float64, device=X_dev.device)
    var = torch.empty(1, N, dtype=torch.float64, device=X_dev.device)
    ci = torch.tensor(cont, dtype=torch.long, device=X_dev.device)
    ki = torch.tensor(cat, dtype=torch.long, device=X_dev.device)
    for r0 in range(0, N, chunk):
        Xq = X_dev[r0:r0 + chunk]
        Xq = torch.where(torch.isfinite(Xq), Xq, torch.full_like(Xq, -1.0))
        if dc:
            t = math.sqrt(5.0) * torch.cdist(Xq[:, ci] * il, Xs[:, :dc], compute_mode='donot_use_mm_for_euclid_dist')
        else:
            t = torch.zeros(Xq.shape[0], n, dtype=torch.float64, device=X_dev.device)
        hs = torch.zeros_like(t)
        for q in range(dk):
            hs += (Xq[:, ki[q], None] != Xs[None, :, dc + q]).to(torch.float64) * hk[q]
        Ks = c * ((1.0 + t + t * t / 3.0) * torch.exp(-t - hs))
        if getattr(types, 'has_conditions', False):
            qpat = Xq <= -1.0                                             # original column order
            tpat = torch.zeros(n, Xq.shape[1], dtype=torch.bool, device=X_dev.device)
            if dc:
                tpat[:, ci] = Xs[:, :dc] <= -il
            if dk:
                tpat[:, ki] = Xs[:, dc:dc + dk] <= -1.0
            Ks = Ks * (~((qpat[:, None, :] != tpat[None, :, :]).any(dim=2))).to(torch.float64)
        mean = Ks @ alpha
        V = Ks @ LinvT
        yv = torch.clamp(diag - (V * V).sum(1), min=0.0)
        sd = torch.sqrt(yv)
        v = torch.clamp(sd * sd, min=VERY_SMALL_NUMBER)
        mu[0, r0:r0 + chunk] = mean * ystd + ymean
        var[0, r0:r0 + chunk] = v * (ystd * ystd)
    return mu, var
