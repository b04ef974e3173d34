"""Host cost of PinnedStage.upload for the bootstrap-draw array (490 KB), step by step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'efficient-tlbo_b200')):
    sys.path.insert(0, p)
import numpy as np, torch
from tlbo_b200 import ops
T = time.perf_counter
st = ops.PinnedStage()
z = np.random.standard_normal((30, 34, 60))
small = np.random.standard_normal(700)
for arr, name in ((z, 'z 490 KB'), (small, 'rank inputs 5.6 KB')):
    for _ in range(5):
        st.upload('a', arr)
    torch.cuda.synchronize()
    t = T()
    for _ in range(200):
        st.upload('a', arr)
    print('%-20s upload %.1f us' % (name, (T() - t) / 200 * 1e6))
    torch.cuda.synchronize()
    buf = st.bufs['a']
    src = torch.from_numpy(arr)
    view = buf[:arr.nbytes].view(src.dtype).reshape(src.shape)
    t = T()
    for _ in range(200):
        view.numpy()[...] = arr
    print('   memcpy into pinned %.1f us' % ((T() - t) / 200 * 1e6))
    dev = ops._dev()
    t = T()
    for _ in range(200):
        d = view.to(dev, non_blocking=True)
    print('   .to(non_blocking)  %.1f us' % ((T() - t) / 200 * 1e6))
    torch.cuda.synchronize()
    out = torch.empty(view.shape, dtype=view.dtype, device=dev)
    t = T()
    for _ in range(200):
        out.copy_(view, non_blocking=True)
    print('   copy_ into a fixed device buffer %.1f us' % ((T() - t) / 200 * 1e6))
    torch.cuda.synchronize()
    t = T()
    for _ in range(200):
        ops._dev()
    print('   _dev() %.1f us' % ((T() - t) / 200 * 1e6))
