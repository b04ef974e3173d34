"""Golden fixtures (tests/golden/hotpath_v1.npz, minted by tests/golden/make_golden.py):
the oracle must keep reproducing them (CPU), and the CUDA path must match them (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))
G = np.load(os.path.join(HERE, 'golden', 'hotpath_v1.npz'))


def test_oracle_reproduces_golden():
    import make_golden
    fresh = make_golden.build()
    assert sorted(fresh.keys()) == sorted(G.files)
    for k in G.files:
        np.testing.assert_allclose(np.asarray(fresh[k], dtype=np.float64), np.asarray(G[k], dtype=np.float64),
                                   rtol=1e-10, atol=1e-12, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['rf', 'c5'])
def test_cuda_matches_golden_gp(tag):
    import torch
    from tlbo_b200 import ops
    types, X, theta = G[tag + '_types'], G[tag + '_X'], G[tag + '_theta']
    n = X.shape[0]
    f, g, st = ops.gp_nll_grad_batched([X], [G[tag + '_ynorm']], types, [theta])
    assert abs(f[0] - float(G[tag + '_nll'])) <= 1e-8 * abs(float(G[tag + '_nll']))
    np.testing.assert_allclose(g[0], G[tag + '_grad'], rtol=1e-6, atol=1e-8)
    pack = ops.SurrogatePack(1, n, len(types))
    st, L, Kinv = ops.gp_factorize_batched([X], [G[tag + '_ynorm']], types, [theta], [float(G[tag + '_ymean'])],
                                           [float(G[tag + '_ystd'])], pack, want_dense=True)
    np.testing.assert_allclose(L[0, :n, :n], G[tag + '_L'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(Kinv[0, :n, :n], G[tag + '_Kinv'], rtol=1e-6, atol=1e-9 * np.abs(G[tag + '_Kinv']).max())
    mu, var = ops.gp_predict(pack, types, G[tag + '_Xq'])
    np.testing.assert_allclose(mu[0].cpu().numpy(), G[tag + '_mu'].ravel(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(var[0].cpu().numpy(), G[tag + '_var'].ravel(), rtol=1e-6, atol=1e-13)
    dev = torch.device('cuda')
    order = G[tag + '_order']
    res, m, v, ei = ops.predict_ei_fused(pack, types, torch.ones(1, dtype=torch.float64, device=dev), None,
                                         torch.from_numpy(G[tag + '_Xq']).to(dev), float(G[tag + '_eta']),
                                         keys=torch.from_numpy(G[tag + '_keys']).to(dev),
                                         excluded=torch.from_numpy(np.sort(order[:2]).astype(np.int64)).to(dev),
                                         want_rows=True)
    np.testing.assert_allclose(ei.cpu().numpy(), G[tag + '_ei'].ravel(), rtol=1e-6, atol=1e-12)
    assert res[2] == int(G[tag + '_pick'])


@pytest.mark.gpu
def test_cuda_matches_golden_weights():
    from tlbo_b200 import ops
    got = ops.rank_loss_counts(G['rl_y'], G['rl_ys'], G['rl_lo'], G['rl_hi'])
    assert (got == G['rl_counts']).all()
    loss, npairs, grad = ops.PairLogistic(G['pl_A'], G['rl_y'])(G['pl_w'])
    assert abs(loss - float(G['pl_loss'])) <= 1e-12
    np.testing.assert_allclose(grad, G['pl_grad'], rtol=1e-9, atol=1e-13)


# ---- MCMC-marginalised GP path (tests/golden/mcmc_v1.npz, minted by make_golden_mcmc.py) ----------
GM = np.load(os.path.join(HERE, 'golden', 'mcmc_v1.npz'))


def test_oracle_reproduces_golden_mcmc():
    import make_golden_mcmc
    fresh = make_golden_mcmc.build()
    assert sorted(fresh.keys()) == sorted(GM.files)
    for k in GM.files:
        a, b = np.asarray(fresh[k], dtype=np.float64), np.asarray(GM[k], dtype=np.float64)
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), k
        np.testing.assert_allclose(a[fin], b[fin], rtol=1e-9, atol=1e-11, err_msg=k)


@pytest.mark.gpu
@pytest.mark.parametrize('tag', ['rf', 'c4'])
def test_cuda_matches_golden_mcmc(tag):
    import torch
    from oracle.gp import KernelSpec
    from oracle.gp_mcmc import prior_bounds_raw
    from tlbo_b200 import ops
    types, X, yn = GM[tag + '_types'], GM[tag + '_X'], GM[tag + '_ynorm']
    spec = KernelSpec(types)
    p0 = GM[tag + '_p0']
    W, H = p0.shape
    nsteps = GM[tag + '_zz'].shape[0] // 2
    batch = ops.upload_batch([X], [yn])
    # _ll at the initial ensemble
    st, _, lp0, _ = ops.gp_mcmc_chain(batch, types, W, 0, p0[None], None, None, prior_bounds_raw(spec),
                                      spec.default_theta()).finish()
    want = GM[tag + '_lp0']
    fin = np.isfinite(want)
    assert st[0] == 0 and np.array_equal(np.isfinite(lp0[0]), fin)
    np.testing.assert_allclose(lp0[0][fin], want[fin], rtol=1e-9, atol=1e-9)
    # the 40-step chain from the golden random tables
    tabs = ops.McmcTables([tuple(GM[tag + k] for k in ('_sidx', '_cidx', '_zz', '_fac', '_logu'))])
    h = ops.gp_mcmc_chain(batch, types, W, nsteps, p0[None], tabs, None, prior_bounds_raw(spec), spec.default_theta())
    st, pos, lp, nacc = h.finish()
    assert np.array_equal(nacc[0], GM[tag + '_nacc'])
    np.testing.assert_allclose(pos[0], GM[tag + '_pos'], rtol=1e-9, atol=1e-12)
    # walker GPs at the golden positions + marginalised posterior
    pack = ops.SurrogatePack(W, X.shape[0], len(types))
    theta_dev = ops.h2d(np.clip(GM[tag + '_pos'], -50, 50))
    fst = ops.gp_factorize_walkers(batch, types, theta_dev, [float(GM[tag + '_ymean'])], [float(GM[tag + '_ystd'])],
                                   pack, 0, W)
    assert not fst.cpu().numpy().any()
    mu, var = ops.gp_predict(pack, types, GM[tag + '_Xq'])
    gmu, gvar = ops.mcmc_marginalize(mu, var, 1, W)
    np.testing.assert_allclose(gmu[0].cpu().numpy(), GM[tag + '_mu'], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(gvar[0].cpu().numpy(), GM[tag + '_var'], rtol=1e-6, atol=1e-13)
