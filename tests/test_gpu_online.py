"""GPU parity of the online scenario (SURVEY 8 f rank 1; BASELINE configs[2] shape): the product's
``SMBO_ONLINE`` -- ensemble facade on the device, EI through the fused kernel, batched local search +
sorted random search -- against the oracle's literal sequential restatement, iteration by iteration.
The oracle's GPs are teacher-forced with the device-selected hyper-parameters (fit parity is tested
elsewhere); chosen configurations must be IDENTICAL arrays, challengers' EI within 1e-6."""
import numpy as np
import pytest

import oracle.facade as ofacade
from oracle import ei_optimization as oeo
from oracle.gp import OracleGP

pytestmark = pytest.mark.gpu


class ForcedGP(OracleGP):
    queue = []

    def train(self, X, Y, do_optimize=True, theta=None):
        return super().train(X, Y, do_optimize=False, theta=ForcedGP.queue.pop(0))


@pytest.fixture
def forced(monkeypatch):
    monkeypatch.setattr(ofacade, 'OracleGP', ForcedGP)
    ForcedGP.queue = []
    yield ForcedGP
    ForcedGP.queue = []


@pytest.mark.parametrize('method,K,iters', [('rgpe', 3, 9), ('topo', 2, 8)])
def test_online_loop_matches_oracle(method, K, iters, forced):
    from tlbo_b200 import synthetic
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.facade.topo_variant1 import OBTLV
    from tlbo_b200.framework.smbo_online import SMBO_ONLINE
    seed = 4465
    b = synthetic.make_benchmark(n_tasks=K + 1, T=200, seed=seed)
    sources, _ = synthetic.leave_one_out_sources(b['tables'], 0, K, seed)
    types, const, fam = b['types'], b['const_mask'], b['family']

    def objective(config):
        x = config.get_array() if hasattr(config, 'get_array') else np.asarray(config)
        return float(fam.evaluate(0, x[np.newaxis, :])[0])

    cs = TableSpace(types, const_mask=const)
    prod = dict(rgpe=RGPE, topo=OBTLV)[method](cs, sources, None, seed, surrogate_type='gp', num_src_hpo_trial=50)
    forced.queue = [prod._pack.theta[k].cpu().numpy() for k in range(K)]
    ora = dict(rgpe=ofacade.OracleRGPE, topo=ofacade.OracleOBTLV)[method](types, sources, seed)
    assert not forced.queue
    smbo = SMBO_ONLINE(objective, cs, prod, random_seed=seed, max_runs=iters, num_points=400)
    state = np.random.get_state()
    osmbo = oeo.OracleSMBOOnline(objective, oeo.OracleSpace(types, const), ora, seed, num_points=400)
    np.random.set_state(state)
    for it in range(iters):
        s0 = np.random.get_state()
        cfg, st, perf, _ = smbo.iterate()
        s1 = np.random.get_state()
        n_prev = len(osmbo.configurations)
        if n_prev >= 3:
            forced.queue = [prod._pack.theta[K + j].cpu().numpy() for j in range(6 if n_prev >= 5 else 1)]
        np.random.set_state(s0)
        ocfg = osmbo.iterate()
        s2 = np.random.get_state()
        assert not forced.queue
        assert s1[2] == s2[2] and (s1[1] == s2[1]).all()
        assert np.array_equal(cfg.get_array(), ocfg), (it, cfg.get_array(), ocfg)
        assert perf == osmbo.perfs[-1]
    st = smbo.acq_optimizer.local_search.last_stats
    assert st['steps'] > 0 and st['launches'] <= st['steps'] + 1
    # every random stream ended in the same state
    assert (smbo.acq_optimizer.rng.get_state()[1] == osmbo.acq_rng.get_state()[1]).all()
    assert (cs._rng.get_state()[1] == osmbo.space.rng.get_state()[1]).all()


def test_online_trials_on_threads_select_the_configurations_of_their_solo_runs():
    """Three online-scenario RGPE trials on three threads (a CUDA stream and a private copy of the process-global
    np.random stream each, tlbo_b200/utils/global_stream.py) propose exactly the configurations each proposes alone:
    the hill-climb entry points (tlbo_local_search_climb / tlbo_ensemble_ei_small) are thread-safe."""
    import threading
    import torch
    from tlbo_b200 import synthetic
    from tlbo_b200.config_space import TableSpace
    from tlbo_b200.facade.rgpe import RGPE
    from tlbo_b200.framework.smbo_online import SMBO_ONLINE
    from tlbo_b200.utils.global_stream import thread_stream
    K, iters = 3, 6

    def run_trial(trial):
        seed = 4465 + trial
        b = synthetic.make_benchmark(n_tasks=K + 1, T=200, seed=seed)
        sources, _ = synthetic.leave_one_out_sources(b['tables'], 0, K, seed)
        types, const, fam = b['types'], b['const_mask'], b['family']

        def objective(config):
            return float(fam.evaluate(0, np.asarray(config.get_array())[np.newaxis, :])[0])
        cs = TableSpace(types, const_mask=const)
        cs.seed(seed)
        fac = RGPE(cs, sources, None, seed, surrogate_type='gp', num_src_hpo_trial=50)
        smbo = SMBO_ONLINE(objective, cs, fac, random_seed=seed, max_runs=10 ** 6, num_points=400)
        for _ in range(iters):
            smbo.iterate()
        return [c.get_array().tobytes() for c in smbo.configurations]

    solo = [run_trial(t) for t in range(3)]
    got, errs = {}, []
    bar = threading.Barrier(3)

    def work(t):
        try:
            with thread_stream(), torch.cuda.stream(torch.cuda.Stream()):
                bar.wait()
                got[t] = run_trial(t)
                torch.cuda.current_stream().synchronize()
        except Exception as ex:                      # pragma: no cover
            errs.append(repr(ex))
            bar.abort()
    ts = [threading.Thread(target=work, args=(t,)) for t in range(3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for t in range(3):
        assert got[t] == solo[t]
    assert solo[0] != solo[1]
