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
