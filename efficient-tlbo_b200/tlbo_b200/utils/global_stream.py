"""Per-thread stand-ins for numpy's PROCESS-GLOBAL legacy random stream.

The reference seeds and draws from ``np.random`` directly (``np.random.seed`` in tlbo/framework/smbo_offline.py:69,
``np.random.normal`` in tlbo/facade/rgpe.py:77,97, ConfigSpace's ``np.random.random()``), so one process = one BO
trial.  To run several independent trials in ONE process -- one thread and one CUDA stream each, so that their fit
kernels (66 CTAs apiece) share a GPU's 148 SMs -- every trial thread gets its own ``RandomState`` behind the
module-level functions the hot path calls: ``RandomState(seed)`` produces the stream ``np.random.seed(seed)`` would
have produced globally, so each trial sees exactly the draws it would see alone in a process.  Threads that have not
entered ``thread_stream()`` fall through to numpy's real global functions.

**Host turns.**  Trial threads that all run interpreter code at once hand the interpreter lock back and forth at every
C-ABI call (tens per BO step); with ``thread_stream(turns=True)`` a trial runs its host segment -- from the read-back of
one step's selected row to the last launch of the next step -- while holding one process-wide lock, and gives it up
only where it would block anyway: ``waiting()`` wraps the device read-backs (``ops.d2h``) and the large bootstrap
draws (numpy generates them without the interpreter lock).  The trials then pipeline: one prepares and launches while
the others wait for their kernels."""
import contextlib
import threading
import time

import numpy as np

_NAMES = ('seed', 'standard_normal', 'random', 'normal', 'get_state', 'set_state', 'rand', 'randn', 'random_sample')
_tls = threading.local()
_lock = threading.Lock()
_installed = [0]
_orig = {}


_turn = threading.Lock()
trace = None          # tools/host_turn_trace.py: list receiving (release time, re-acquire time) of every waiting() block


@contextlib.contextmanager
def waiting():
    """Give up the calling trial thread's host turn for the duration of a blocking call (no-op for threads without
    one)."""
    held = getattr(_tls, 'turn', False)
    if held:
        if trace is not None:
            t0 = time.perf_counter()
        _turn.release()
    try:
        yield
    finally:
        if held:
            _turn.acquire()
            if trace is not None:
                trace.append((t0, time.perf_counter()))


def _dispatch(name):
    big = name == 'standard_normal'

    def f(*a, **k):
        rs = getattr(_tls, 'rs', None)
        if rs is None:
            return _orig[name](*a, **k)
        if big and getattr(_tls, 'turn', False):
            size = a[0] if a else k.get('size')
            if size is not None and np.prod(size) >= 4096:
                with waiting():
                    return rs.standard_normal(*a, **k)
        return getattr(rs, name)(*a, **k)
    f.__name__ = name
    return f


@contextlib.contextmanager
def thread_stream(turns=False):
    """Inside the block, the calling thread's ``np.random.<seed|standard_normal|random|...>`` calls go to a private
    legacy ``RandomState`` (initially seeded by the OS like numpy's own; the BO loop re-seeds it).  ``turns``: the
    thread also takes part in the host-turn protocol above (every barrier / synchronisation of its own must then sit
    inside ``waiting()``)."""
    with _lock:
        if _installed[0] == 0:
            for n in _NAMES:
                _orig[n] = getattr(np.random, n)
                setattr(np.random, n, _dispatch(n))
        _installed[0] += 1
    _tls.rs = np.random.RandomState()
    if turns:
        _turn.acquire()
        _tls.turn = True
    try:
        yield _tls.rs
    finally:
        if turns:
            _tls.turn = False
            _turn.release()
        _tls.rs = None
        with _lock:
            _installed[0] -= 1
            if _installed[0] == 0:
                for n in _NAMES:
                    setattr(np.random, n, _orig[n])
