"""BASELINE config 4: batched GP predict sweep -- 30 surrogates x n in {50,100,200} observations x
N candidates, d in {5, 20} (all continuous) -- fused mean/variance/EI/arg-max, candidates per
second, with a parity spot check against the oracle on a subsample.
usage: predict_sweep.py [--max-n-cand 1e7] > profiles/r1_predict_sweep.json"""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'efficient-tlbo_b200'))
import numpy as np, torch
from tlbo_b200 import ops
from oracle import gp as ogp

ap = argparse.ArgumentParser(); ap.add_argument('--max-n-cand', type=float, default=1e7); ap.add_argument('--M', type=int, default=30)
args = ap.parse_args()
dev = torch.device('cuda')
peak = ops.fp64_peak_probe()
out = {'fp64_peak_tflops': {'dfma': peak[0], 'dmma': peak[1]}, 'rows': []}
for d in (5, 20):
    types = np.zeros(d, int)
    spec = ogp.KernelSpec(types)
    for n in (50, 100, 200):
        rng = np.random.RandomState(100 * d + n)
        M = args.M
        pack = ops.SurrogatePack(M, n, d)
        ogs = []
        Xs, ys, ths, yms, yss = [], [], [], [], []
        for m in range(M):
            X = rng.rand(n, d); y = np.sin(X.sum(1) * (1 + 0.1 * m)) + 0.05 * rng.randn(n)
            th = np.concatenate([[rng.uniform(-0.5, 1.0)], rng.uniform(-1.2, 0.05, d), [rng.uniform(-9, -5)]])
            ym, ysd = y.mean(), y.std()
            Xs.append(X); ys.append((y - ym) / ysd); ths.append(th); yms.append(ym); yss.append(ysd)
        st = ops.gp_factorize_batched(Xs, ys, types, ths, yms, yss, pack)
        assert not st.any()
        w = torch.full((M,), 1.0 / M, dtype=torch.float64, device=dev)
        flops_per_cand = 30.0 + M * (n * (3 * d + 12) + 2 * n + n * n + 2 * n)
        for N in (10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7):
            if N > args.max_n_cand:
                continue
            Xc = torch.rand(N, d, dtype=torch.float64, device=dev)
            keys = torch.rand(N, dtype=torch.float64, device=dev)
            ws = ops.FusedWorkspace()
            for _ in range(2):
                res = ops.predict_ei_fused(pack, types, w, None, Xc, 0.1, keys=keys, ws=ws, nmax=n)
            reps = 3 if N >= 10 ** 6 else 10
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.predict_ei_fused(pack, types, w, None, Xc, 0.1, keys=keys, ws=ws, nmax=n, sync=False)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            row = dict(d=d, n=n, M=M, N=N, ms=ms, candidates_per_s=N / (ms * 1e-3),
                       tflops_algorithmic=flops_per_cand * N / (ms * 1e-3) / 1e12,
                       frac_of_dmma_peak=flops_per_cand * N / (ms * 1e-3) / 1e12 / peak[1],
                       hbm_gbs_algorithmic=N * 8.0 * (d + 1) / (ms * 1e-3) / 1e9)
            if N == 10 ** 4:
                # parity spot check on 200 rows against the oracle
                r, mu, var, ei = ops.predict_ei_fused(pack, types, w, None, Xc[:200].contiguous(), 0.1, want_rows=True, nmax=n)
                Xh = Xc[:200].cpu().numpy()
                mo = np.zeros(200); vo = np.zeros(200)
                for m in range(M):
                    g = ogp.OracleGP(types, 1); g.X = Xs[m]; g.y = ys[m]; g.spec = spec; g.mean_y = yms[m]; g.std_y = yss[m]
                    g._fit_at(ths[m]); g.is_trained = True
                    a, b = g.predict(Xh)
                    mo += a.ravel() / M; vo += b.ravel() / M / M
                row['parity_max_rel_mu'] = float(np.max(np.abs(mu.cpu().numpy() - mo) / np.maximum(np.abs(mo), 1e-9)))
                row['parity_max_rel_var'] = float(np.max(np.abs(var.cpu().numpy() - np.maximum(vo, 1e-10)) / np.maximum(vo, 1e-10)))
            out['rows'].append(row)
            print('d=%2d n=%3d N=%8d  %8.3f ms  %.3e cand/s  %.2f TFLOP/s (%.1f%% of DMMA peak)' % (
                d, n, N, ms, row['candidates_per_s'], row['tflops_algorithmic'], 100 * row['frac_of_dmma_peak']), file=sys.stderr)
            del Xc, keys
print(json.dumps(out, indent=1))
