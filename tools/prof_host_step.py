"""cProfile of the host side of the RGPE offline step (what holds the interpreter lock per BO iteration)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402


class A:
    pass


args = A()
args.method = 'rgpe'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
fac, smbo = bench.make_product(args, wl)
for _ in range(5):
    smbo.iterate()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    smbo.iterate()
torch.cuda.synchronize()
print('plain ms/step', (time.perf_counter() - t0) / 20 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    smbo.iterate()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(45)
