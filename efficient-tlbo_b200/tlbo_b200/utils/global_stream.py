"""Per-thread stand-ins for numpy's PROCESS-GLOBAL legacy random stream.

The reference seeds and draws from ``np.random`` directly (``np.random.seed`` in tlbo/framework/smbo_offline.py:69,
``np.random.normal`` in tlbo/facade/rgpe.py:77,97, ConfigSpace's ``np.random.random()``), so one process = one BO
trial.  To run several independent trials in ONE process -- one thread and one CUDA stream each, so that their fit
kernels (66 CTAs apiece) share a GPU's 148 SMs -- every trial thread gets its own ``RandomState`` behind the
module-level functions the hot path calls: ``RandomState(seed)`` produces the stream ``np.random.seed(seed)`` would
have produced globally, so each trial sees exactly the draws it would see alone in a process.  Threads that have not
entered ``thread_stream()`` fall through to numpy's real global functions."""
import contextlib
import threading

import numpy as np

_NAMES = ('seed', 'standard_normal', 'random', 'normal', 'get_state', 'set_state', 'rand', 'randn', 'random_sample')
_tls = threading.local()
_lock = threading.Lock()
_installed = [0]
_orig = {}


def _dispatch(name):
    def f(*a, **k):
        rs = getattr(_tls, 'rs', None)
        if rs is None:
            return _orig[name](*a, **k)
        return getattr(rs, name)(*a, **k)
    f.__name__ = name
    return f


@contextlib.contextmanager
def thread_stream():
    """Inside the block, the calling thread's ``np.random.<seed|standard_normal|random|...>`` calls go to a private
    legacy ``RandomState`` (initially seeded by the OS like numpy's own; the BO loop re-seeds it)."""
    with _lock:
        if _installed[0] == 0:
            for n in _NAMES:
                _orig[n] = getattr(np.random, n)
                setattr(np.random, n, _dispatch(n))
        _installed[0] += 1
    _tls.rs = np.random.RandomState()
    try:
        yield _tls.rs
    finally:
        _tls.rs = None
        with _lock:
            _installed[0] -= 1
            if _installed[0] == 0:
                for n in _NAMES:
                    setattr(np.random, n, _orig[n])
