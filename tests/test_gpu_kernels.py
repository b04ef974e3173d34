"""GPU parity tests of the CUDA kernels (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): mean / variance / EI within 1e-6 relative in FP64;
ranking-loss counts and selected candidate indices bit-exact."""
import numpy as np
import pytest

from oracle import gp as ogp
from oracle.acquisition import OracleEI, first_unevaluated, sort_by_acq_value
from oracle.facade import loss_der3, loss_func3, rank_loss_vectorised

pytestmark = pytest.mark.gpu

RF_TYPES = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])
RTOL = 1e-6


def _data(rng, n, types, const_cols=(2, 4, 5, 8)):
    d = len(types)
    X = rng.rand(n, d)
    for j in range(d):
        if types[j] > 0:
            X[:, j] = rng.randint(0, types[j], size=n)
    if len(types) == 9:
        X[:, list(const_cols)] = 0.0
    y = np.sin(3 * X[:, 3 % d]) + 0.5 * X[:, 6 % d] ** 2 + 0.3 * X[:, 0] + 0.05 * rng.randn(n)
    return X, y


def _norm(y):
    m, s = np.mean(y), np.std(y)
    if s == 0:
        s = 1
    return (y - m) / s, m, s


def _theta(rng, spec):
    lb = spec.log_bounds()
    th = np.array([rng.uniform(-1, 1.5)] + list(rng.uniform(-2.0, 0.08, size=spec.H - 2)) + [rng.uniform(-12, -3)])
    return np.clip(th, lb[:, 0], lb[:, 1])


@pytest.mark.parametrize('n,types', [(5, RF_TYPES), (50, RF_TYPES), (75, RF_TYPES), (33, np.zeros(5, int)),
                                     (100, np.zeros(20, int)), (200, np.zeros(5, int))])
def test_nll_grad_matches_oracle(n, types):
    from tlbo_b200 import ops
    rng = np.random.RandomState(n)
    spec = ogp.KernelSpec(types)
    Xs, ys, ths = [], [], []
    for b in range(4):
        X, y = _data(rng, n, types)
        yn, _, _ = _norm(y)
        Xs.append(X); ys.append(yn); ths.append(_theta(rng, spec))
    ths[0] = spec.default_theta()
    f, g, st = ops.gp_nll_grad_batched(Xs, ys, types, ths)
    for b in range(4):
        fo, go = ogp.nll(ths[b], spec, Xs[b], ys[b])
        if fo >= 1e25:
            assert f[b] == 1e25 and not g[b].any()
            continue
        assert st[b] == 0
        assert abs(f[b] - fo) <= 1e-8 * max(1.0, abs(fo)), (b, f[b], fo)
        np.testing.assert_allclose(g[b], go, rtol=1e-6, atol=1e-7 * max(1.0, np.abs(go).max()))


def test_nll_cholesky_failure_returns_1e25():
    """gp.py:183-186: LinAlgError -> (1e25, zeros).  Duplicate rows + minimal noise."""
    from tlbo_b200 import ops
    types = np.zeros(3, int)
    rng = np.random.RandomState(0)
    X = np.repeat(rng.rand(4, 3), 5, axis=0)
    y = rng.randn(20)
    spec = ogp.KernelSpec(types)
    th = np.array([2.0, 0.08, 0.08, 0.08, -25.0])
    fo, go = ogp.nll(th, spec, X, y)
    f, g, st = ops.gp_nll_grad_batched([X], [y], types, [th])
    assert (fo >= 1e25) == (f[0] >= 1e25)


@pytest.mark.parametrize('n,types', [(3, RF_TYPES), (50, RF_TYPES), (75, RF_TYPES), (200, np.zeros(5, int))])
def test_factorize_matches_oracle(n, types):
    from tlbo_b200 import ops
    rng = np.random.RandomState(n + 1)
    spec = ogp.KernelSpec(types)
    B = 3
    Xs, ys, ths, yms, yss, ogs = [], [], [], [], [], []
    for b in range(B):
        X, y = _data(rng, n, types)
        g = ogp.OracleGP(types, 1)
        th = _theta(rng, spec)
        g.train(X, y, do_optimize=False, theta=th)
        Xs.append(X); ys.append(g.y); ths.append(th); yms.append(g.mean_y); yss.append(g.std_y); ogs.append(g)
    pack = ops.SurrogatePack(B, n, len(types))
    st, L, Kinv = ops.gp_factorize_batched(Xs, ys, types, ths, yms, yss, pack, want_dense=True)
    assert not st.any()
    for b in range(B):
        np.testing.assert_allclose(L[b, :n, :n], ogs[b].L, rtol=1e-9, atol=1e-12)
        scale = np.abs(ogs[b].K_inv).max()
        np.testing.assert_allclose(Kinv[b, :n, :n], ogs[b].K_inv, rtol=1e-6, atol=1e-9 * scale)
        al = pack.alpha[b, :n].cpu().numpy()
        np.testing.assert_allclose(al, ogs[b].alpha, rtol=1e-6, atol=1e-9 * np.abs(ogs[b].alpha).max())
    # small-query predict
    Xq, _ = _data(rng, 37, types)
    mu, var = ops.gp_predict(pack, types, Xq)
    mu, var = mu.cpu().numpy(), var.cpu().numpy()
    for b in range(B):
        mo, vo = ogs[b].predict(Xq)
        np.testing.assert_allclose(mu[b], mo.ravel(), rtol=RTOL, atol=1e-9)
        np.testing.assert_allclose(var[b], vo.ravel(), rtol=RTOL, atol=1e-12)


def _fitted_pack(rng, types, ns, ncap=None):
    from tlbo_b200 import ops
    spec = ogp.KernelSpec(types)
    Xs, ys, ths, yms, yss, ogs = [], [], [], [], [], []
    for n in ns:
        X, y = _data(rng, n, types)
        g = ogp.OracleGP(types, 1)
        th = _theta(rng, spec)
        g.train(X, y, do_optimize=False, theta=th)
        Xs.append(X); ys.append(g.y); ths.append(th); yms.append(g.mean_y); yss.append(g.std_y); ogs.append(g)
    M = len(ns)
    pack = ops.SurrogatePack(M, ncap or max(ns), len(types))
    # surrogates of different sizes: factorise size groups separately into their slots
    for m in range(M):
        st = ops.gp_factorize_batched([Xs[m]], [ys[m]], types, [ths[m]], [yms[m]], [yss[m]], pack, slot0=m)
        assert not st.any()
    return pack, ogs


def _oracle_ensemble(ogs, w, skip, Xc, mode=0):
    if mode == 0:
        mu = np.zeros((Xc.shape[0], 1)); var = np.zeros((Xc.shape[0], 1))
        # accumulation order of the kernel = pack order
        for m, g in enumerate(ogs):
            if skip is not None and skip[m]:
                continue
            a, b = g.predict(Xc)
            mu += w[m] * a
            var += w[m] * w[m] * b
        return mu, var
    mu, var = ogs[0].predict(Xc)
    den = 0.75
    for m in range(1, len(ogs)):
        a, _ = ogs[m].predict(Xc)
        mu += w[m] * a
        den += w[m]
    return mu / den, var


class _Ens(object):
    def __init__(self, ogs, w, skip, mode, floor):
        self.ogs, self.w, self.skip, self.mode, self.floor = ogs, w, skip, mode, floor

    def predict_marginalized_over_instances(self, X):
        mu, var = _oracle_ensemble(self.ogs, self.w, self.skip, X, self.mode)
        var[var < self.floor] = self.floor
        var[np.isnan(var)] = self.floor
        return mu, var


@pytest.mark.parametrize('types,ns,N,mode', [
    (RF_TYPES, [50] * 5 + [20], 3000, 0),
    (RF_TYPES, [50] * 29 + [75], 20000, 0),
    (RF_TYPES, [7, 50, 50, 50], 1000, 1),
    (np.zeros(5, int), [100] * 3, 4099, 0),
    (np.zeros(20, int), [200, 50], 2500, 0),
    (RF_TYPES, [3, 50], 17, 0),
])
def test_fused_predict_ei_argmax(types, ns, N, mode):
    import torch
    from tlbo_b200 import ops
    rng = np.random.RandomState(len(ns) * 1000 + N)
    pack, ogs = _fitted_pack(rng, types, ns)
    M = len(ns)
    w = rng.rand(M)
    w[rng.rand(M) < 0.3] = 0.0
    w[-1] = max(w[-1], 0.2)
    w = w / w.sum() if mode == 0 else w
    skip = (rng.rand(M) < 0.2).astype(np.int32)
    skip[-1] = 0
    if mode == 1:
        skip[:] = 0
    Xc, _ = _data(rng, N, types)
    # some candidates coincide with training rows (variance cancellation path)
    Xc[:min(ns[-1], N)] = ogs[-1].X[:min(ns[-1], N)]
    eta = float(np.min(ogs[-1].y)) * ogs[-1].std_y + ogs[-1].mean_y
    ens = _Ens(ogs, w, skip, mode, 1e-10)
    ei_o = OracleEI(ens)
    ei_o.update(eta=eta)
    acq = ei_o(Xc)
    rs = np.random.RandomState(7)
    order, keys = sort_by_acq_value(acq, rs)
    evaluated = list(order[:3]) + [int(order[10 % N])]
    want = first_unevaluated(order, evaluated)

    dev = torch.device('cuda')
    res, mu, var, ei = ops.predict_ei_fused(
        pack, types, torch.from_numpy(w).to(dev), torch.from_numpy(skip).to(dev), torch.from_numpy(Xc).to(dev),
        eta, keys=torch.from_numpy(keys).to(dev),
        excluded=torch.from_numpy(np.sort(np.array(evaluated, dtype=np.int64))).to(dev), mode=mode, want_rows=True)
    mo, vo = ens.predict_marginalized_over_instances(Xc)
    np.testing.assert_allclose(mu.cpu().numpy(), mo.ravel(), rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(var.cpu().numpy(), vo.ravel(), rtol=RTOL, atol=1e-13)
    np.testing.assert_allclose(ei.cpu().numpy(), acq.ravel(), rtol=RTOL, atol=1e-12 + 1e-6 * acq.max())
    assert res[3] == 0
    # selected index: bit-exact whenever the oracle's own winner is separated from the
    # runner-up by more than the FP64 parity tolerance; otherwise the GPU pick must be
    # within tolerance of the best EI.
    eo = acq.ravel().copy()
    eo[evaluated] = -np.inf
    top2 = np.sort(eo)[-2:]
    if top2[1] - top2[0] > 1e-9 * max(top2[1], 1e-300):
        assert res[2] == want
    else:
        assert eo[res[2]] >= top2[1] * (1 - 1e-6)
    assert abs(res[0] - eo[res[2]]) <= 1e-6 * max(eo[res[2]], 1e-12) + 1e-12
    assert res[1] == keys[res[2]]


def test_fused_argmax_bit_exact_given_identical_ei():
    """Index selection given the GPU's own EI column: lexsort((key, ei))[::-1] + first
    unevaluated, bit-exact (ties on EI broken by the larger key)."""
    import torch
    from tlbo_b200 import ops
    types = RF_TYPES
    rng = np.random.RandomState(3)
    pack, ogs = _fitted_pack(rng, types, [50, 50, 30])
    N = 50000
    Xc, _ = _data(rng, N, types)
    # duplicate rows -> exact EI ties
    Xc[N // 2:] = Xc[:N - N // 2]
    w = np.array([0.3, 0.2, 0.5])
    dev = torch.device('cuda')
    keys = np.random.RandomState(11).rand(N)
    eta = float(np.min(ogs[-1].y)) * ogs[-1].std_y + ogs[-1].mean_y
    for n_ex in (0, 1, 40):
        res0, mu, var, ei = ops.predict_ei_fused(pack, types, torch.from_numpy(w).to(dev), None,
                                                 torch.from_numpy(Xc).to(dev), eta,
                                                 keys=torch.from_numpy(keys).to(dev), want_rows=True)
        e = ei.cpu().numpy()
        order = np.lexsort((keys, e))[::-1]
        ex = sorted(int(i) for i in order[:n_ex])
        res = ops.predict_ei_fused(pack, types, torch.from_numpy(w).to(dev), None, torch.from_numpy(Xc).to(dev), eta,
                                   keys=torch.from_numpy(keys).to(dev),
                                   excluded=torch.tensor(ex, dtype=torch.int64, device=dev) if ex else None)
        assert res[2] == first_unevaluated(order, ex)
        assert res[0] == e[res[2]]


def test_fused_sharded_offsets_agree_with_single_call():
    import torch
    from tlbo_b200 import ops
    types = np.zeros(5, int)
    rng = np.random.RandomState(5)
    pack, ogs = _fitted_pack(rng, types, [40, 60])
    N = 10007
    Xc, _ = _data(rng, N, types)
    dev = torch.device('cuda')
    w = torch.tensor([0.4, 0.6], dtype=torch.float64, device=dev)
    keys = torch.from_numpy(np.random.RandomState(1).rand(N)).to(dev)
    X = torch.from_numpy(Xc).to(dev)
    full = ops.predict_ei_fused(pack, types, w, None, X, 0.1, keys=keys)
    parts = []
    bounds = [0, 2500, 5000, 7500, N]
    for a, b in zip(bounds[:-1], bounds[1:]):
        parts.append(ops.predict_ei_fused(pack, types, w, None, X[a:b].contiguous(), 0.1,
                                          keys=keys[a:b].contiguous(), idx_offset=a))
    best = max(parts, key=lambda r: (r[0], r[1], r[2]))
    assert best[:3] == full[:3]


@pytest.mark.parametrize('n,S,Mr', [(5, 30, 6), (75, 30, 35), (200, 30, 8), (1, 2, 2)])
def test_rank_loss_counts_bit_exact(n, S, Mr):
    from tlbo_b200 import ops
    rng = np.random.RandomState(n)
    y = rng.randn(n)
    y[rng.rand(n) < 0.2] = y[0]          # ties in y
    ys = rng.randn(S, Mr, n)
    ys[:, :, ::3] = np.round(ys[:, :, ::3], 1)   # ties in the samples
    lo = rng.randint(0, n, size=Mr)
    hi = np.minimum(lo + rng.randint(0, n + 1, size=Mr), n)
    lo[0], hi[0] = 0, n
    got = ops.rank_loss_counts(y, ys, lo, hi)
    for s in range(S):
        for m in range(Mr):
            assert got[s, m] == rank_loss_vectorised(y, ys[s, m], lo[m], hi[m])
    got_u = ops.rank_loss_counts(y, ys[:1], np.zeros(Mr, int), np.full(Mr, n), upper_only=True)
    iu = np.triu_indices(n, 1)
    for m in range(Mr):
        a = (y[:, None] < y[None, :])[iu]
        b = (ys[0, m][:, None] < ys[0, m][None, :])[iu]
        assert got_u[0, m] == np.count_nonzero(a ^ b)


@pytest.mark.parametrize('n,m', [(3, 5), (20, 29), (75, 30), (200, 30)])
def test_pairwise_logistic_loss_grad(n, m):
    from tlbo_b200 import ops
    rng = np.random.RandomState(n * m)
    A = rng.randn(n, m)
    y = rng.rand(n)
    y[1] = y[0]
    pl = ops.PairLogistic(A, y)
    for _ in range(3):
        w = rng.rand(m)
        w /= w.sum()
        loss, npairs, grad = pl(w)
        lo = loss_func3(y, A.dot(w))
        go = loss_der3(y, A, w)
        assert abs(loss - lo) <= 1e-10 * max(1, abs(lo))
        np.testing.assert_allclose(grad, go, rtol=1e-8, atol=1e-12)


def test_device_mt19937_equals_numpy_legacy_random_sample():
    """``tlbo_mt19937_random_sample``: the tie-break keys ``rng.rand(N)`` generated on the device with the MT19937
    state resident in HBM -- bit-identical doubles, identical final state, across block boundaries, odd stream
    positions (a double straddling two state blocks) and consecutive calls."""
    import torch
    from tlbo_b200 import ops
    for seed, sizes, pre in ((4465, [5, 311, 312, 313, 1, 100000, 7], 0), (1, [624 * 3, 2, 999], 3), (2 ** 31 + 7, [10 ** 6], 1)):
        ref = np.random.RandomState(seed)
        for _ in range(pre):
            ref.randint(0, 10000)            # odd / arbitrary stream position
        dev = ops.DeviceRandomState(ref)
        for n in sizes:
            out = torch.empty(n, dtype=torch.float64, device='cuda')
            dev.rand_into(out)
            want = ref.rand(n)
            assert np.array_equal(out.cpu().numpy(), want), (seed, n)
        back = np.random.RandomState(0)
        torch.cuda.synchronize()
        dev.to_host(back)
        a, b = back.get_state(), ref.get_state()
        assert np.array_equal(a[1], b[1]) and a[2] == b[2]
        assert back.rand() == ref.rand()


def test_device_mt19937_sliced_jump_ahead_equals_numpy():
    """Multi-CTA key generation (``tlbo_mt19937_jump`` + ``tlbo_mt19937_random_sample_sliced``): slice c starts from
    the state jumped ahead by c * S doubles (polynomial jump-ahead, tlbo_b200/utils/mt_jump.py).  Consecutive
    draws of the hinted length, a draw of another length in between (sequential walk + re-jump), odd stream
    positions: bit-identical doubles and final state."""
    import torch
    from tlbo_b200 import ops
    for seed, n, pre in ((4465, 5 * 65536 + 777, 0), (7, 2 * 65536, 3), (2 ** 31 + 7, 1000003, 1)):
        ref = np.random.RandomState(seed)
        for _ in range(pre):
            ref.randint(0, 10000)
        dev = ops.DeviceRandomState(ref, n_hint=n)
        assert dev._slices >= 2
        out = torch.empty(n, dtype=torch.float64, device='cuda')
        for rep in range(3):
            dev.rand_into(out)
            assert np.array_equal(out.cpu().numpy(), ref.rand(n)), (seed, n, rep)
            if rep == 1:
                small = torch.empty(1001, dtype=torch.float64, device='cuda')
                dev.rand_into(small)
                assert np.array_equal(small.cpu().numpy(), ref.rand(1001))
        back = np.random.RandomState(0)
        torch.cuda.synchronize()
        dev.to_host(back)
        a, b = back.get_state(), ref.get_state()
        assert np.array_equal(a[1], b[1]) and a[2] == b[2]
        assert back.rand() == ref.rand()


def test_mt_jump_polynomial_matches_numpy_stream():
    """Host side of the jump-ahead: the window jumped by 2 * 312000 words equals numpy's key array after 312000
    doubles (characteristic polynomial by Berlekamp-Massey, t^J mod phi, XOR of shifted windows)."""
    from tlbo_b200.utils import mt_jump as mj
    rs = np.random.RandomState(5)
    rs.random_sample(1)                       # first draw regenerates the block: key = block 1, pos = 2
    key = rs.get_state()[1].copy()
    rs.random_sample(312 * 1000)
    assert np.array_equal(mj.jump_window(key, mj.bit_positions(mj.jump_poly(624 * 1000))), rs.get_state()[1])
    assert bin(mj.charpoly()).count('1') == 135 and mj.charpoly().bit_length() == 19938


def test_merge_best_matches_host_rule():
    """tlbo_merge_best (the exchange step of the row-sharded pool) against the host statement of the rule,
    including ties on EI / key and shards without an unevaluated candidate."""
    import torch
    from tlbo_b200 import ops
    from tlbo_b200.distributed import merge_best
    rng = np.random.RandomState(5)
    for G in (1, 2, 8, 37):
        for case in range(6):
            q = np.zeros((G, 4))
            q[:, 0] = rng.choice([0.0, 0.25, 0.5], size=G) if case % 2 else rng.rand(G)
            q[:, 1] = rng.choice([0.1, 0.9], size=G) if case % 3 == 0 else rng.rand(G)
            q[:, 2] = rng.permutation(G * 10)[:G]
            q[rng.rand(G) < 0.3, 2] = -1
            if case == 5:
                q[:, 2] = -1
            q[:, 3] = rng.randint(0, 3, size=G)
            want = merge_best(q)
            got = ops.merge_best(torch.from_numpy(q).cuda()).cpu().numpy()
            assert int(got[3]) == want[3]
            if want[2] < 0:
                assert got[2] < 0
            else:
                assert (float(got[0]), float(got[1]), int(got[2])) == (want[0], want[1], want[2])


@pytest.mark.parametrize('types,ns,N,mode,ncap,nmax', [
    (RF_TYPES, [50] * 29 + [33], 26, 0, 72, 50),          # the online hill-climb step: staged slabs, rows in the arguments
    (RF_TYPES, [50] * 5 + [20], 200, 0, 56, 0),           # 1 800 doubles of rows: read through the pointer, nmax unknown
    (RF_TYPES, [7, 50, 50, 50], 31, 1, 64, 50),           # TST combination (mode 1), first entry initialises
    (np.zeros(5, int), [100] * 3, 17, 0, 104, 100),       # staging plan beyond the shared-memory budget: global path
    (np.zeros(20, int), [60, 50], 9, 2, 64, 60),          # POGPE combination (mode 2)
    (RF_TYPES, [3, 50, 0], 5, 0, 56, 50),                 # an empty slot (n = 0) is dropped
])
def test_ensemble_ei_small_matches_oracle_and_fused(types, ns, N, mode, ncap, nmax):
    """``tlbo_ensemble_ei_small`` (one CTA per (row, surrogate), last CTA of a row combines): EI of a small query set
    against the oracle's ensemble EI and -- value for value -- against ``tlbo_gp_predict`` + the fused kernel in
    all-resident mode; the counters are left zero and untouched rows keep their sentinel."""
    import torch
    from tlbo_b200 import ops
    rng = np.random.RandomState(len(ns) * 100 + N)
    live = [n for n in ns if n > 0]
    pack, ogs = _fitted_pack(rng, types, live, ncap=ncap)
    if len(live) < len(ns):                                # append empty slots
        full = ops.SurrogatePack(len(ns), ncap, len(types))
        for m in range(len(live)):
            full.copy_slot_from(m, pack, m)
        pack = full
    M = len(ns)
    w = rng.rand(M) + 0.1
    skip = np.zeros(M, dtype=np.int32)
    if mode == 0:
        w = w / w.sum()
        if M > 3:
            skip[1] = 1
    order = np.arange(M, dtype=np.int32)
    if mode == 0 and M > 2:
        order = np.concatenate([[M - 1], np.arange(M - 1)]).astype(np.int32)      # RGPE: target first
        if ns[-1] == 0:
            order = np.arange(M, dtype=np.int32)
    Xq, _ = _data(rng, N, types)
    Xq[0] = ogs[0].X[0]                                    # a training row: variance cancellation path
    eta = float(np.min(ogs[-1].y)) * ogs[-1].std_y + ogs[-1].mean_y
    dev = torch.device('cuda')
    dw, dskip, dorder = (torch.from_numpy(a).to(dev) for a in (w, skip, order))
    sq = ops.SmallQuery(max(256, N), len(types), M)
    sq.X_np[:N] = Xq
    sq.ei_np[:] = -7.0
    pack.c.nmax = nmax
    ops.ensemble_ei_small(pack, types, dw, dskip, dorder, mode, sq, N, eta, 0.0, 1e-10)
    ops.stream_synchronize()
    got = sq.ei_np[:N].copy()
    assert np.all(sq.ei_np[N:] == -7.0)
    assert int(sq.counter.abs().sum().item()) == 0
    # (a) the same posteriors through tlbo_gp_predict and the fused kernel in all-resident mode
    Xd = torch.from_numpy(Xq).to(dev)
    mu, var = ops.gp_predict(pack, types, Xd)
    logical = ops.SurrogatePack(M, 8, len(types))
    logical.n.copy_(pack.n.clamp(max=1))
    res, _, _, ei2 = ops.predict_ei_fused(logical, types, dw, dskip, Xd, eta, var_floor=1e-10, mode=mode, order=dorder,
                                          want_rows=True, nmax=8, cache=(mu, var))
    np.testing.assert_allclose(got, ei2.cpu().numpy(), rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(sq.mu[:M * N].view(M, N).cpu().numpy(), mu.cpu().numpy(), rtol=1e-10, atol=1e-12)
    # (b) the oracle (modes 0 / 1; accumulation order does not change the value beyond rounding)
    if mode in (0, 1) and 0 not in ns:
        ens = _Ens(ogs, w, skip if mode == 0 else None, mode, 1e-10)
        if mode == 1:
            ens = _Ens([ogs[i] for i in order], w[order], None, 1, 1e-10)
        ei_o = OracleEI(ens)
        ei_o.update(eta=eta)
        acq = ei_o(Xq).ravel()
        np.testing.assert_allclose(got, acq, rtol=RTOL, atol=1e-12 + 1e-6 * acq.max())


@pytest.mark.parametrize('S,K,n_folds,n', [(30, 29, 5, 41), (30, 3, 0, 4), (30, 29, 5, 75), (7, 1, 2, 10)])
def test_rgpe_weights_kernel_bit_exact(S, K, n_folds, n):
    """``tlbo_rgpe_weights`` against numpy's statement of rgpe.py:106-133 (np.argmin ties, histogram / num_sample,
    median-vs-95th-percentile dilution flags) on count tables with many ties."""
    import torch
    from tlbo_b200 import ops
    rng = np.random.RandomState(S * 100 + K)
    for trial in range(6):
        hi = [n * n, 40, 3, 2][trial % 4]                       # small ranges: many exact ties
        counts = rng.randint(0, hi + 1, size=(S, K + n_folds)).astype(np.int64)
        losses = np.empty((S, K + 1), dtype=np.int64)
        losses[:, :K] = counts[:, :K]
        losses[:, K] = counts[:, K:].sum(axis=1) if n_folds else n * n
        w_want = np.bincount(np.argmin(losses, axis=1), minlength=K + 1) / S
        srt = np.sort(losses, axis=0)
        skip_want = np.array([bool(m > srt[int(S * 0.95), -1]) for m in srt[int(S * 0.5), :K]] + [False], dtype=np.int32)
        w = torch.zeros(K + 1, dtype=torch.float64, device='cuda')
        skip = torch.zeros(K + 1, dtype=torch.int32, device='cuda')
        ops.rgpe_weights(torch.from_numpy(counts).cuda(), K, n_folds, n * n, (w, skip))
        assert np.array_equal(w.cpu().numpy(), w_want), trial
        assert np.array_equal(skip.cpu().numpy(), skip_want), trial
