"""PinnedStage.upload inside the running RGPE step, piece by piece (tools/micro_upload.py is the stand-alone figure)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from tlbo_b200 import ops


class A:
    pass


args = A()
args.method = 'rgpe'; args.pool = 100000; args.sources = 29; args.n0 = 40; args.surrogate = 'gp'; args.mcmc_steps = [250, 250]
wl = bench.build_workload(0, 4465, args.pool, args.sources, args.n0)
fac, smbo = bench.make_product(args, wl)
for _ in range(5):
    smbo.iterate()
T = time.perf_counter
acc = {}


def upload(self, tag, arr):
    t0 = T()
    arr = np.ascontiguousarray(arr)
    nbytes = max(arr.nbytes, 8)
    buf = self.bufs.get(tag)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(ops.round_up(max(2 * nbytes, 1 << 16), 4096), dtype=torch.uint8).pin_memory()
        self.bufs[tag] = buf
    ta = T()
    src = torch.from_numpy(arr)
    tb = T()
    v1 = buf[:arr.nbytes]
    tc = T()
    v2 = v1.view(src.dtype)
    td = T()
    view = v2.reshape(src.shape)
    t1 = T()
    b = acc.setdefault(tag + ' views', [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    b[0] += 1; b[1] += ta - t0; b[2] += tb - ta; b[3] += tc - tb; b[4] += td - tc; b[5] += t1 - td
    view.numpy()[...] = arr
    t2 = T()
    dev = ops._dev()
    t3 = T()
    out = view.to(dev, non_blocking=True)
    t4 = T()
    a = acc.setdefault(tag, [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += t1 - t0; a[2] += t2 - t1; a[3] += t3 - t2; a[4] += t4 - t3
    return out


ops.PinnedStage.upload = upload
N = 40
torch.cuda.synchronize()
t0 = T()
for _ in range(N):
    smbo.iterate()
torch.cuda.synchronize()
print('ms/step %.3f' % ((T() - t0) / N * 1e3))
for tag, a in acc.items():
    if tag.endswith(' views'):
        print('%-18s contig+get %.1f  from_numpy %.1f  slice %.1f  view %.1f  reshape %.1f us' % (
            tag, *[x / a[0] * 1e6 for x in a[1:]]))
        continue
    print('%-12s calls/step %.1f  views %.1f us  memcpy %.1f us  _dev %.1f us  .to %.1f us' % (
        tag, a[0] / N, a[1] / a[0] * 1e6, a[2] / a[0] * 1e6, a[3] / a[0] * 1e6, a[4] / a[0] * 1e6))
