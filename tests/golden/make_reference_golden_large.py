"""Mints tests/golden/reference_v2_large.npz: the REFERENCE'S OWN modules at BASELINE sizes.

Same method as make_reference_golden.py (unmodified reference from /root/reference, stand-ins only for its
absent third-party imports), at the sizes BASELINE.json's configs are quoted on instead of the small
shapes of reference_v1.npz:

  kern_/nll_/fit_/pred_/ei_  tags  c20    20 continuous columns, n = 60          (config 4, d = 20)
                                   rf75   RF-shaped space, n = 75                 (config 1/2, last iteration)
                                   rf200  RF-shaped space, n = 200                (config 4, n = 200: the
                                          shared-memory / global panel fit path)
  big_rgpe_*, big_topo_*     SMBO_OFFLINE loops with K = 29 sources x 50 observations over a 20 000-row
                             target table (config 1/2 shape): rows, weights, dilution flags, integer
                             ranking-loss tables, every fitted theta, np.random stream states

Build container only (needs /root/reference):   python tests/golden/make_reference_golden_large.py [iters_rgpe iters_topo]
"""
import collections
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as M      # noqa: E402  (sets the thread env vars, numpy aliases, import paths)
import numpy as np                     # noqa: E402

ref_gp = M.ref_gp


def gp_level_large(out):
    rng = np.random.RandomState(20261022)
    for tag, cs, types, n, full_fit in (('c20', M.cont_space(20), np.zeros(20, int), 60, True),
                                        ('rf75', M.rf_space(), M.RF_TYPES, 75, True),
                                        ('rf200', M.rf_space(), M.RF_TYPES, 200, True)):
        t0 = time.time()
        X, y = M.data(rng, n, types)
        Xq, _ = M.data(rng, 256, types)
        Xq[:5] = X[:5]
        model = M.build_model('gp', cs, np.random.RandomState(7))
        H = len(model.kernel.theta)
        thetas = np.array([np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.5, 0.05, H - 2),
                                           [rng.uniform(-9, -4)]]) for _ in range(3)])
        k = model.kernel.clone_with_theta(thetas[0])
        out.update({'kern_%s_types' % tag: types, 'kern_%s_X' % tag: X, 'kern_%s_y' % tag: y,
                    'kern_%s_Xq' % tag: Xq, 'kern_%s_thetas' % tag: thetas, 'kern_%s_K' % tag: k(X),
                    'kern_%s_Ks' % tag: k(Xq[:32], X)})
        M.quiet(model._train, X, y.copy(), do_optimize=False)
        model._all_priors = model._get_all_priors(add_bound_priors=False)
        vals = [model._nll(th.copy()) for th in thetas]
        out['nll_%s_f' % tag] = np.array([v[0] for v in vals])
        out['nll_%s_g' % tag] = np.array([v[1] for v in vals])
        out['nll_%s_ynorm' % tag] = model.gp.y_train_.copy()
        model.gp.kernel.theta = thetas[1]
        model.gp.fit(X, model.gp.y_train_.copy())
        model.is_trained = True
        mu, var = model.predict(Xq)
        out['pred_%s_fixed_mu' % tag], out['pred_%s_fixed_var' % tag] = mu, var
        model = M.build_model('gp', cs, np.random.RandomState(7))
        rec = []
        orig = ref_gp.optimize.fmin_l_bfgs_b

        def spy(func, x0, **kw):
            r = orig(func, x0, **kw)
            rec.append((np.array(x0, dtype=np.float64), np.array(r[0]), float(r[1]), int(r[2]['funcalls'])))
            return r
        ref_gp.optimize.fmin_l_bfgs_b = spy
        try:
            M.quiet(model.train, X, y.copy())
        finally:
            ref_gp.optimize.fmin_l_bfgs_b = orig
        out['fit_%s_p0' % tag] = np.array([r[0] for r in rec])
        out['fit_%s_theta' % tag] = np.array([r[1] for r in rec])
        out['fit_%s_f' % tag] = np.array([r[2] for r in rec])
        out['fit_%s_funcalls' % tag] = np.array([r[3] for r in rec])
        out['fit_%s_theta_star' % tag] = np.array(model.hypers)
        mu, var = model.predict(Xq)
        out['pred_%s_mu' % tag], out['pred_%s_var' % tag] = mu, var
        ei = M.EI(model)
        eta = float(np.min(y))
        ei.update(model=model, eta=eta, num_data=n)
        out['ei_%s_eta' % tag] = eta
        out['ei_%s_acq' % tag] = ei._compute(Xq)
        print(tag, 'n', n, 'funcalls', out['fit_%s_funcalls' % tag].tolist(), '%.1fs' % (time.time() - t0), flush=True)


def loop_level_large(out, iters_by_method):
    types = M.RF_TYPES
    N, K, n_src, seed = 20000, 29, 50, 4465
    rng = np.random.RandomState(2029)
    Xt, yt = M.data(rng, N, types, task=0)
    src = [M.data(rng, n_src, types, task=1 + (t % 7)) for t in range(K)]
    out['big_target_X'], out['big_target_y'] = Xt, yt
    out['big_source_X'] = np.array([s[0] for s in src])
    out['big_source_y'] = np.array([s[1] for s in src])
    out['big_types'], out['big_seed'], out['big_K'] = types, seed, K
    for method, cls in (('rgpe', M.RGPE), ('topo', M.OBTLV)):
        iters = iters_by_method[method]
        if iters <= 0:
            continue
        cs = M.rf_space()
        as_dict = lambda X, y: collections.OrderedDict((M.Configuration(cs, vector=x), float(v)) for x, v in zip(X, y))
        sources = [as_dict(*s) for s in src]
        target = as_dict(Xt, yt)
        thetas = []
        orig_train = ref_gp.GaussianProcess._train

        def spy_train(self, X, y, do_optimize=True):
            r = orig_train(self, X, y, do_optimize)
            thetas.append(np.array(self.hypers, dtype=np.float64))
            return r
        losses, orig_argmin = [], np.argmin

        def spy_argmin(a, *args, **kw):
            if isinstance(a, list) and len(a) == K + 1 and not args and not kw:
                losses.append([float(v) for v in a])
            return orig_argmin(a, *args, **kw)
        ref_gp.GaussianProcess._train = spy_train
        np.argmin = spy_argmin
        try:
            t0 = time.time()
            fac = M.quiet(cls, cs, sources, list(target.keys()), seed, surrogate_type='gp', num_src_hpo_trial=n_src)
            n_source_fits = len(thetas)
            print(method, 'facade built: %d source fits, %.1fs' % (n_source_fits, time.time() - t0), flush=True)
            smbo = M.quiet(M.SMBO_OFFLINE, target, cs, fac, random_seed=seed, max_runs=iters, source_hpo_data=sources,
                           num_src_hpo_trial=n_src, surrogate_type='gp', initial_runs=3,
                           logging_dir='/tmp/tlbo_ref_logs')
            keys = list(target.keys())
            index_of = {id(k): i for i, k in enumerate(keys)}
            rows, states, fits_per_iter, w_hist, flags, secs = [], [], [], [], [], []
            for it in range(iters):
                before = len(thetas)
                t1 = time.time()
                cfg, _, _, _ = M.quiet(smbo.iterate)
                secs.append(time.time() - t1)
                rows.append(index_of[id(cfg)] if id(cfg) in index_of else keys.index(cfg))
                fits_per_iter.append(len(thetas) - before)
                st = np.random.get_state()
                states.append(np.concatenate([st[1].astype(np.int64), [st[2], st[3]], [np.float64(st[4]).view(np.int64)]]))
                w_hist.append(np.array(fac.w, dtype=np.float64).ravel() if fac.w is not None else np.zeros(K + 1))
                flags.append(np.array(getattr(fac, 'ignored_flag', [False] * K), dtype=bool))
                print(method, 'it', it, 'row', rows[-1], '%.1fs' % secs[-1], flush=True)
        finally:
            ref_gp.GaussianProcess._train = orig_train
            np.argmin = orig_argmin
        p = 'big_%s_' % method
        out[p + 'rows'] = np.array(rows)
        out[p + 'perfs'] = np.array(smbo.perfs, dtype=np.float64)
        out[p + 'thetas'] = np.array(thetas)
        out[p + 'n_source_fits'] = n_source_fits
        out[p + 'fits_per_iter'] = np.array(fits_per_iter)
        out[p + 'np_random_state'] = np.array(states)
        out[p + 'w'] = np.array(w_hist)
        out[p + 'ignored'] = np.array(flags)
        out[p + 'target_weight'] = np.array(fac.target_weight, dtype=np.float64)
        out[p + 'losses'] = np.array(losses, dtype=np.int64).reshape(-1, K + 1)
        out[p + 'eta_list'] = np.array(fac.eta_list, dtype=np.float64)
        out[p + 'reference_seconds_per_iteration'] = np.array(secs)


def main():
    it_rgpe = int(sys.argv[1]) if len(sys.argv) > 1 else 45
    it_topo = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = {}
    gp_level_large(out)
    loop_level_large(out, dict(rgpe=it_rgpe, topo=it_topo))
    path = os.path.join(HERE, 'reference_v2_large.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes,', len(out), 'arrays')


if __name__ == '__main__':
    main()
