import sys, time
sys.path.insert(0, '/root/repo/efficient-tlbo_b200')
import numpy as np, torch
from tlbo_b200 import ops
from tlbo_b200._cabi import ptr
T = time.perf_counter
def bench(name, f, n=2000):
    for _ in range(50): f()
    torch.cuda.synchronize(); t0 = T()
    for _ in range(n): f()
    dt = (T() - t0) / n * 1e6
    torch.cuda.synchronize()
    print('%-40s %6.2f us' % (name, dt))
pinned = torch.empty(4000, dtype=torch.float64, pin_memory=True)
bench('pinned.to(non_blocking)', lambda: pinned.to('cuda', non_blocking=True))
bench('torch.empty(pin_memory) 32KB', lambda: torch.empty(4000, dtype=torch.float64, pin_memory=True))
bench('torch.empty(cuda) 1MB', lambda: torch.empty(1 << 20, dtype=torch.uint8, device='cuda'))
bench('torch.zeros(cuda) 6', lambda: torch.zeros(6, dtype=torch.int32, device='cuda'))
types = np.array([2, 2, 0, 0, 0, 0, 0, 0, 0])
bench('_types_arr', lambda: ops._types_arr(types))
lo = np.zeros(11)
bench('_cached_const(11)', lambda: ops._cached_const(lo))
t = torch.empty(4, device='cuda')
bench('ptr() x 20', lambda: [ptr(t) for _ in range(20)])
def ctx():
    with ops._call('gp_fit', 2):
        pass
bench('_call ctx', ctx)
bench('_stream()', lambda: ops._stream())
bench('ops._dev()', lambda: ops._dev())
blob = torch.empty(100000, dtype=torch.uint8, device='cuda')
bench('views x5', lambda: (blob[:800].view(torch.float64).view(10, 10), blob[800:1600].view(torch.float64), blob[2048:4096], blob[4096:4200].view(torch.int32), blob[4200:4224].view(torch.int32)))
import copy
from tlbo_b200.model.model_builder import build_model
from tlbo_b200.config_space import TableSpace
m = build_model('gp', TableSpace(types), np.random.RandomState(1))
bench('fresh_copy', lambda: m.fresh_copy())
