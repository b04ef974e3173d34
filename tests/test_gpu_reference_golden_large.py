"""The CUDA path (through the C ABI) against the REFERENCE'S OWN outputs at BASELINE sizes
(tests/golden/reference_v2_large.npz, produced by tests/golden/make_reference_golden_large.py from the
unmodified reference in the build container): d = 20, n = 75, n = 200 (the shared-memory panel fit path), and
whole SMBO_OFFLINE loops with K = 29 sources over a 20 000-row table up to n = 75 target observations."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
R = np.load(os.path.join(HERE, 'golden', 'reference_v2_large.npz'))
TAGS = ['c20', 'rf75', 'rf200']


def _norm(y):
    m, s = float(np.mean(y)), float(np.std(y))
    return (y - m) / s, m, s


@pytest.mark.parametrize('tag', TAGS)
def test_cuda_nll_matches_reference(tag):
    from tlbo_b200 import ops
    types, X, yn = R['kern_%s_types' % tag], R['kern_%s_X' % tag], R['nll_%s_ynorm' % tag]
    thetas = R['kern_%s_thetas' % tag]
    f, g, st = ops.gp_nll_grad_batched([X] * len(thetas), [yn] * len(thetas), types, list(thetas))
    assert not np.asarray(st).any()
    np.testing.assert_allclose(f, R['nll_%s_f' % tag], rtol=1e-8, atol=0)
    np.testing.assert_allclose(g, R['nll_%s_g' % tag], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('tag', TAGS)
def test_cuda_posterior_and_ei_match_reference(tag):
    import torch
    from tlbo_b200 import ops
    types, X, Xq = R['kern_%s_types' % tag], R['kern_%s_X' % tag], R['kern_%s_Xq' % tag]
    yn, ym, ys = _norm(R['kern_%s_y' % tag])
    n = X.shape[0]
    pack = ops.SurrogatePack(1, n, len(types))
    st = ops.gp_factorize_batched([X], [yn], types, [R['kern_%s_thetas' % tag][1]], [ym], [ys], pack)
    assert not np.asarray(st).any()
    mu, var = ops.gp_predict(pack, types, Xq)
    np.testing.assert_allclose(mu[0].cpu().numpy(), R['pred_%s_fixed_mu' % tag].ravel(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(var[0].cpu().numpy(), R['pred_%s_fixed_var' % tag].ravel(), rtol=1e-6, atol=1e-12)
    st = ops.gp_factorize_batched([X], [yn], types, [R['fit_%s_theta_star' % tag]], [ym], [ys], pack)
    assert not np.asarray(st).any()
    dev = torch.device('cuda')
    res, m, v, ei = ops.predict_ei_fused(pack, types, torch.ones(1, dtype=torch.float64, device=dev), None,
                                         torch.from_numpy(Xq).to(dev), float(R['ei_%s_eta' % tag]), var_floor=1e-5,
                                         want_rows=True)
    np.testing.assert_allclose(m.cpu().numpy(), R['pred_%s_mu' % tag].ravel(), rtol=1e-6, atol=1e-9)
    ref = R['ei_%s_acq' % tag].ravel()
    np.testing.assert_allclose(ei.cpu().numpy(), ref, rtol=1e-6, atol=1e-12)
    assert ref[int(res[2])] == ref.max()


@pytest.mark.parametrize('tag', TAGS)
def test_cuda_fit_matches_reference_restarts(tag):
    """Device L-BFGS-B from the reference's own 11 start points at n = 60 (d = 20), 75 and 200."""
    from tlbo_b200 import ops
    from tlbo_b200.model.gp import KernelInfo
    types, X = R['kern_%s_types' % tag], R['kern_%s_X' % tag]
    yn, ym, ys = _norm(R['kern_%s_y' % tag])
    p0, f_ref = R['fit_%s_p0' % tag], R['fit_%s_f' % tag]
    lb = KernelInfo(types).bounds
    pack = ops.SurrogatePack(1, X.shape[0], len(types))
    st, rt, rf, ri = ops.gp_fit_batched([X], [yn], types, p0, lb[:, 0], lb[:, 1], [ym], [ys], pack, return_restarts=True)
    assert not np.asarray(st).any()
    close = np.abs(rf[0] - f_ref) <= 1e-5 * np.maximum(1.0, np.abs(f_ref))
    print('%s: %d of %d restarts at the reference optimum; evaluations device %s reference %s' % (
        tag, close.sum(), len(f_ref), ri[0, :, 0].tolist(), R['fit_%s_funcalls' % tag].tolist()))
    assert close.sum() >= len(f_ref) - 3, (rf[0], f_ref)
    assert rf[0].min() <= f_ref.min() + 1e-6 * max(1.0, abs(f_ref.min()))


def _state_vec():
    st = np.random.get_state()
    return np.concatenate([st[1].astype(np.int64), [st[2], st[3]], [np.float64(st[4]).view(np.int64)]])


def _product(method, types, sources, seed):
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.facade.topo_variant1 import OBTLV
    cs = TableSpace(types)
    return cs, dict(rgpe=RGPE, topo=OBTLV)[method](cs, sources, None, seed, surrogate_type='gp', num_src_hpo_trial=50)


class _ForcedFits(object):
    def __init__(self, monkeypatch):
        import tlbo_b200.facade.base_facade as BF
        import tlbo_b200.model.gp as G
        self.queue = []
        orig = G.fit_models_async

        def forced(models, Xs, Ys, do_optimize=True):
            for m in models:
                m.kernel.theta = np.array(self.queue.pop(0), dtype=np.float64)
            return orig(models, Xs, Ys, do_optimize=False)
        monkeypatch.setattr(BF, 'fit_models_async', forced)
        monkeypatch.setattr(BF.BaseFacade, '_native_prep_ok', lambda self, X, y: False)


@pytest.mark.parametrize('method', ['rgpe', 'topo'])
def test_cuda_big_loop_matches_reference_given_theta(method, monkeypatch):
    """K = 29, 20 000-row table, the reference's FULL trajectory (76 RGPE / 40 OBTLV iterations, n up to 75):
    with the device GPs placed at the reference's theta the CUDA path reproduces its rows, 30 x 30 integer
    ranking-loss tables, dilution flags and np.random stream states exactly, weights to 1e-6."""
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    forced = _ForcedFits(monkeypatch)
    K, seed, types = int(R['big_K']), int(R['big_seed']), R['big_types']
    p = 'big_%s_' % method
    sources = [(R['big_source_X'][k].copy(), R['big_source_y'][k].copy()) for k in range(K)]
    thetas = [th for th in R[p + 'thetas']]
    nsrc = int(R[p + 'n_source_fits'])
    forced.queue = thetas[:nsrc]
    cs, fac = _product(method, types, sources, seed)
    assert not forced.queue
    np.testing.assert_allclose(fac.eta_list, R[p + 'eta_list'], rtol=1e-13)
    smbo = SMBO_OFFLINE((R['big_target_X'], R['big_target_y']), cs, fac, source_hpo_data=None, surrogate_type='gp',
                        initial_runs=3, random_seed=seed, max_runs=1000)
    used, loss_rows = nsrc, 0
    for it, row in enumerate(R[p + 'rows']):
        nfit = int(R[p + 'fits_per_iter'][it])
        forced.queue = thetas[used:used + nfit]
        used += nfit
        smbo.iterate()
        assert not forced.queue, 'iteration %d: fewer device fits than the reference made' % it
        got = smbo.configurations[-1]
        assert got == int(row), 'iteration %d: device picked row %d, reference %d' % (it, got, row)
        np.testing.assert_array_equal(_state_vec(), R[p + 'np_random_state'][it], err_msg='np.random stream, it %d' % it)
        np.testing.assert_allclose(np.asarray(fac.w, dtype=np.float64).ravel(), R[p + 'w'][it], rtol=1e-6, atol=1e-7,
                                   err_msg='weights, iteration %d' % it)
        if method == 'rgpe' and nfit:
            np.testing.assert_array_equal(np.asarray(fac.ignored_flag, dtype=bool), R[p + 'ignored'][it])
            if nfit == 6:
                np.testing.assert_array_equal(fac.last_losses, R[p + 'losses'][loss_rows:loss_rows + 30])
            loss_rows += 30
    np.testing.assert_array_equal(np.asarray(smbo.perfs), R[p + 'perfs'])


@pytest.mark.parametrize('method', ['rgpe', 'topo'])
def test_cuda_big_loop_with_device_fit_tracks_reference(method):
    """No teacher forcing: the device's own 11-restart MAP fits of all 29 + 6 surrogates.  Two L-BFGS-B
    implementations agree to optimiser tolerance only in general, and a trajectory of arg-max picks amplifies a
    rounding-level difference at the first near-tie.  Measured on B200 (profiles/r2_gpu_reference_golden_large.log):
    OBTLV 40 of 40 picks identical; RGPE 62 of 76 as a prefix, after which the two runs evaluate the SAME rows in a
    different order and end with the same incumbent (an earlier build of this round, with single-pivot sweeps in
    natural order, coincided on all 76).  Asserted: at least half of the trajectory as an identical prefix, at least
    90 % of the evaluated rows in common, an incumbent as good as the reference's."""
    from tlbo_b200.framework.smbo_offline import SMBO_OFFLINE
    K, seed, types = int(R['big_K']), int(R['big_seed']), R['big_types']
    p = 'big_%s_' % method
    sources = [(R['big_source_X'][k].copy(), R['big_source_y'][k].copy()) for k in range(K)]
    cs, fac = _product(method, types, sources, seed)
    smbo = SMBO_OFFLINE((R['big_target_X'], R['big_target_y']), cs, fac, source_hpo_data=None, surrogate_type='gp',
                        initial_runs=3, random_seed=seed, max_runs=1000)
    rows = R[p + 'rows']
    for _ in rows:
        smbo.iterate()
    got = np.asarray(smbo.configurations)
    same = int(np.argmax(np.concatenate([got != rows, [True]])))
    common = len(set(got.tolist()) & set(rows.tolist()))
    print('%s (K=29, N=20000): %d of %d picks identical as a prefix, %d rows in common; best perf device %.6f reference %.6f'
          '; first difference: device %s reference %s'
          % (method, same, len(rows), common, min(smbo.perfs), R[p + 'perfs'].min(), got[same:same + 4].tolist(),
             rows[same:same + 4].tolist()))
    assert same >= len(rows) // 2, (got.tolist(), rows.tolist())
    assert common >= int(0.9 * len(rows))
    y = R['big_target_y']
    span = y.max() - y.min()
    assert min(smbo.perfs) <= R[p + 'perfs'].min() + 0.05 * span


@pytest.mark.parametrize('tag', ['rf200', 'c20'])
def test_large_n_path_matches_reference(tag, monkeypatch):
    """model/gp_large.py (the device-wide path SCoT's stacked GP takes above LARGE_N rows) against the reference's
    own numbers: NLL + gradient, the 11-restart fit (scipy L-BFGS-B over the device likelihood), the pack slot
    it fills (read back through the posterior kernels) and the two-step pool posterior."""
    import torch
    from tlbo_b200 import ops
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.model import gp_large
    from tlbo_b200.model.model_builder import build_model
    types, X, Xq = R['kern_%s_types' % tag], R['kern_%s_X' % tag], R['kern_%s_Xq' % tag]
    yn, ym, ys = _norm(R['kern_%s_y' % tag])
    P = gp_large._Problem(X, yn, types)
    for i, theta in enumerate(R['kern_%s_thetas' % tag]):
        f, g = gp_large.nll_grad(P, theta)
        assert abs(f - R['nll_%s_f' % tag][i]) <= 1e-9 * abs(R['nll_%s_f' % tag][i])
        np.testing.assert_allclose(g, R['nll_%s_g' % tag][i], rtol=1e-6, atol=1e-7)
    # full train through the model interface with the threshold lowered so that n = 60 / 200 takes the large path
    monkeypatch.setattr(gp_large, 'LARGE_N', 32)
    m = build_model('gp', TableSpace(types), np.random.RandomState(7))
    m.train(X, R['kern_%s_y' % tag].copy())
    assert 'large-n' in m.fit_info_['path']
    f_ref = R['fit_%s_f' % tag]
    f_got = m.fit_info_['restart_f']
    close = np.abs(f_got - f_ref) <= 1e-5 * np.maximum(1.0, np.abs(f_ref))
    print('%s large-n fit: %d of %d restarts at the reference optimum; evaluations %s reference %s' % (
        tag, close.sum(), len(f_ref), m.fit_info_['restart_info'][:, 0].tolist(), R['fit_%s_funcalls' % tag].tolist()))
    assert close.sum() >= len(f_ref) - 2
    assert abs(f_got.min() - f_ref.min()) <= 1e-6 * max(1.0, abs(f_ref.min()))
    # posterior at the REFERENCE's theta* through the slot the large path fills
    P2 = gp_large._Problem(X, yn, types)
    m.mean_y_, m.std_y_ = ym, ys
    gp_large.factorize_into_pack(m, P2, R['fit_%s_theta_star' % tag])
    mu, var = m.predict(Xq)
    np.testing.assert_allclose(mu, R['pred_%s_mu' % tag], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(var, R['pred_%s_var' % tag], rtol=1e-6, atol=1e-12)
    pm, pv = gp_large.posterior_pool(m._pack, m._slot, types, torch.from_numpy(Xq).cuda(), chunk=100)
    np.testing.assert_allclose(pm[0].cpu().numpy(), R['pred_%s_mu' % tag].ravel(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(pv[0].cpu().numpy(), R['pred_%s_var' % tag].ravel(), rtol=1e-6, atol=1e-12)
