"""Thin, typed wrappers over the C ABI (one function per entry point of
``include/tlbo_b200.h``).  torch is used for device memory and streams only."""
import ctypes
import os

import numpy as np
import torch

from tlbo_b200 import _cabi
from tlbo_b200._cabi import HYP_STRIDE, TlboPack, check, ldl, lib, ptr
from tlbo_b200.utils.global_stream import waiting

F64 = torch.float64


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError('tlbo_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback.')
    return torch.device('cuda', torch.cuda.current_device())


_RAW_STREAM = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_synchronize():
    """Wait for torch's current stream (``tlbo_stream_synchronize``)."""
    check(lib.tlbo_stream_synchronize(_stream()))


def _stream():
    """The launching stream (torch's CURRENT stream, so ``with torch.cuda.stream(s)`` is honoured) as a raw
    ``cudaStream_t``.  ``torch.cuda.current_stream()`` builds a Python Stream object (~15 us -- three of
    them per acquisition call of the online local search); the raw query costs well under a microsecond."""
    if _RAW_STREAM is not None:
        return ctypes.c_void_p(_RAW_STREAM(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


TYPES_HAS_CONDITIONS = 0x40000000        # include/tlbo_b200.h: bit 30 of h_types[0]


def _types_arr(types):
    """``h_types`` of the C ABI: int32[d], with the space's ``has_conditions`` fact (a ``TypesArray`` attribute) as
    bit 30 of element 0."""
    t = np.array(np.asarray(types), dtype=np.int32)
    if getattr(types, 'has_conditions', False):
        t[0] |= TYPES_HAS_CONDITIONS
    return np.ascontiguousarray(t)


def round_up(n, m):
    return ((int(n) + m - 1) // m) * m


def batch_cap(n):
    """Row capacity (stride) of a training batch whose largest set has ``n`` rows: a multiple of 8 with room
    for the extra y row of the register-panel kernels, whose tile is chosen from the capacity
    (n + 1 <= 16 * ceil(cap / 16), csrc/gp_nll.cuh:panel_tile_for) -- so n = 57..63 runs 4 x 4 tiles."""
    cap = round_up(max(int(n), 1), 8)
    if cap % 16 == 0 and int(n) == cap:
        cap += 8
    return cap


class _Stats(object):
    """Launch / transfer accounting for bench.py: how many of OUR kernels were launched,
    how many bytes crossed PCIe, and (when ``timing``) a CUDA-event pair around every
    C-ABI call on the launching stream."""

    def __init__(self):
        self.timing = False
        self.reset()

    def reset(self):
        self.launches = {}
        self.h2d = 0
        self.d2h = 0
        self.events = []

    def total_launches(self):
        return sum(self.launches.values())

    def kernel_ms(self):
        """name -> (calls, total ms) from the recorded event pairs (synchronises)."""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.events:
            c, t = out.get(name, (0, 0.0))
            out[name] = (c + 1, t + e0.elapsed_time(e1))
        return out


STATS = _Stats()


class _call(object):
    def __init__(self, name, kernels):
        self.name, self.kernels = name, kernels

    def __enter__(self):
        STATS.launches[self.name] = STATS.launches.get(self.name, 0) + self.kernels
        if STATS.timing:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if STATS.timing:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            STATS.events.append((self.name, self.e0, e1))
        return False


def h2d(a, dtype=None):
    """Host array -> device tensor (counted)."""
    if torch.is_tensor(a):
        if a.is_cuda:
            return a if dtype is None else a.to(dtype)
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    STATS.h2d += t.numel() * t.element_size()
    return t.to(_dev())


class PinnedStage(object):
    """Reusable pinned staging buffers for uploads that must not block the host while a
    kernel is in flight (pageable copies synchronise).  One buffer per tag; a tag may be
    reused once the consumer of the previous upload has been synchronised."""

    def __init__(self):
        self.bufs = {}

    def upload(self, tag, arr):
        arr = np.ascontiguousarray(arr)
        nbytes = max(arr.nbytes, 8)
        buf = self.bufs.get(tag)
        if buf is None or buf.numel() < nbytes:
            # geometric growth: pinning is a ~0.3 ms driver call, and e.g. the bootstrap-draw
            # buffer grows by a few KB every BO iteration
            buf = torch.empty(round_up(max(2 * nbytes, 1 << 16), 4096), dtype=torch.uint8).pin_memory()
            self.bufs[tag] = buf
        src = torch.from_numpy(arr)
        if not arr.nbytes:
            return src.to(_dev())
        view = buf[:arr.nbytes].view(src.dtype).reshape(src.shape)
        view.numpy()[...] = arr                      # plain memcpy (see OfflineSearch.prepare)
        STATS.h2d += arr.nbytes
        return view.to(_dev(), non_blocking=True)


def merge_best(quads, out=None):
    """``tlbo_merge_best``: lexicographic merge of ``quads [G, 4]`` (device) into a device 4-vector."""
    if out is None:
        out = torch.empty(4, dtype=F64, device=quads.device)
    with _call('merge_best', 1):
        check(lib.tlbo_merge_best(ptr(quads), int(quads.shape[0]), ptr(out), _stream()))
    return out


def d2h(t):
    """Device tensor -> numpy (counted; synchronises)."""
    STATS.d2h += t.numel() * t.element_size()
    with waiting():                    # a trial thread gives its host turn away while it blocks (utils/global_stream.py)
        return t.cpu().numpy()


class SurrogatePack(object):
    """Device-resident batch of fitted GP surrogates (``struct tlbo_pack``)."""

    def __init__(self, M, ncap, d):
        dev = _dev()
        ncap = batch_cap(max(ncap, 8))
        self.M, self.ncap, self.d = int(M), int(ncap), int(d)
        self.n = torch.zeros(M, dtype=torch.int32, device=dev)
        self.Xs = torch.zeros(M, ncap, d, dtype=F64, device=dev)
        self.hyp = torch.zeros(M, HYP_STRIDE, dtype=F64, device=dev)
        self.alpha = torch.zeros(M, ncap, dtype=F64, device=dev)
        self.Linv = torch.zeros(M, ncap, ldl(ncap), dtype=F64, device=dev)
        self.theta = torch.zeros(M, d + 2, dtype=F64, device=dev)
        self.c = TlboPack(self.M, self.ncap, self.d, 0, self.n.data_ptr(), self.Xs.data_ptr(),
                          self.hyp.data_ptr(), self.alpha.data_ptr(), self.Linv.data_ptr(),
                          self.theta.data_ptr())

    def ref(self):
        return ctypes.byref(self.c)

    def copy_slot_from(self, dst_slot, other, src_slot):
        """Copy one fitted surrogate from another pack (other.ncap <= self.ncap)."""
        assert other.d == self.d and other.ncap <= self.ncap
        nc = other.ncap
        self.n[dst_slot] = other.n[src_slot]
        self.Xs[dst_slot].zero_()
        self.Xs[dst_slot, :nc] = other.Xs[src_slot]
        self.hyp[dst_slot] = other.hyp[src_slot]
        self.alpha[dst_slot].zero_()
        self.alpha[dst_slot, :nc] = other.alpha[src_slot]
        self.Linv[dst_slot].zero_()
        self.Linv[dst_slot, :nc, :nc] = other.Linv[src_slot, :, :nc]
        self.theta[dst_slot] = other.theta[src_slot]


class PackView(object):
    """``M`` consecutive slots of a pack, starting at ``slot0`` (no copy)."""

    def __init__(self, pack, slot0, M):
        assert 0 <= slot0 and slot0 + M <= pack.M
        self.base, self.slot0 = pack, int(slot0)
        self.M, self.ncap, self.d = int(M), pack.ncap, pack.d
        s = self.slot0
        self.n, self.Xs, self.hyp = pack.n[s:s + M], pack.Xs[s:s + M], pack.hyp[s:s + M]
        self.alpha, self.Linv, self.theta = pack.alpha[s:s + M], pack.Linv[s:s + M], pack.theta[s:s + M]
        self.c = TlboPack(self.M, self.ncap, self.d, 0, self.n.data_ptr(), self.Xs.data_ptr(),
                          self.hyp.data_ptr(), self.alpha.data_ptr(), self.Linv.data_ptr(),
                          self.theta.data_ptr())

    def ref(self):
        return ctypes.byref(self.c)


def pack_view(pack, slot0, M):
    if slot0 == 0 and M == pack.M:
        return pack
    return PackView(pack, slot0, M)


def _pad_batch(X_list, y_list, ncap, d):
    """Host-side batch assembly; ONE upload ([X | y | n] in a single float64 buffer)."""
    B = len(X_list)
    nx, ny = B * ncap * d, B * ncap
    buf = np.zeros(nx + ny + (B + 1) // 2 + 1, dtype=np.float64)
    Xh = buf[:nx].reshape(B, ncap, d)
    yh = buf[nx:nx + ny].reshape(B, ncap)
    nh = buf[nx + ny:].view(np.int32)
    for b, (X, y) in enumerate(zip(X_list, y_list)):
        n = X.shape[0]
        Xh[b, :n] = X
        yh[b, :n] = y
        nh[b] = n
    dbuf = h2d(buf)
    return dbuf[:nx].view(B, ncap, d), dbuf[nx:nx + ny].view(B, ncap), dbuf[nx + ny:].view(torch.int32)[:B]


class PreparedBatch(object):
    """Padded upload buffer of a batched fit, filled by the native ``tlbo_host_prepare_jobs`` (the host half
    of ``build_single_surrogate`` + ``GaussianProcess._train`` for the 1 + 5 training sets of an iteration,
    bit-identical to the numpy statement in ``facade/base_facade.py`` / ``model/gp.py``)."""

    def __init__(self, buf, B, ncap, d, n, ymean, ystd):
        self.buf, self.B, self.ncap, self.d = buf, B, ncap, d
        self.n, self.ymean, self.ystd = n, ymean, ystd


def prepare_jobs(X, y, ranges, outer_guard, normalize, normalize_y=True):
    """``tlbo_host_prepare_jobs``: job b trains on every row of (X, y) except ``ranges[b] = (lo, hi)``.
    ``y`` (float64, contiguous) is mutated in place exactly where the reference mutates the caller's array
    (the all-equal guards).  ``normalize``: 'none' | 'standardize' | 'scale'."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    assert y.dtype == np.float64 and y.flags.c_contiguous and y.ndim == 1
    n, d = X.shape
    B = len(ranges)
    ncap = batch_cap(n)
    nx, ny = B * ncap * d, B * ncap
    # pinned: the upload in gp_fit_batched is then an asynchronous copy (a pageable one blocks the host for ~30 us
    # right in front of the fit launch); torch's caching host allocator recycles the block
    pinned = torch.empty(nx + ny + (B + 1) // 2 + 1, dtype=F64, pin_memory=torch.cuda.is_available())
    buf = pinned.numpy()
    nh = buf[nx + ny:].view(np.int32)
    nh[:] = 0
    lo = np.array([r[0] for r in ranges], dtype=np.int32)
    hi = np.array([r[1] for r in ranges], dtype=np.int32)
    og = np.array(outer_guard, dtype=np.int32)
    ym = np.empty(B, dtype=np.float64)
    ys = np.empty(B, dtype=np.float64)
    check(lib.tlbo_host_prepare_jobs(ptr(X), ptr(y), n, d, B, ptr(lo), ptr(hi), ptr(og),
                                     {'none': 0, 'standardize': 1, 'scale': 2}[normalize], 1 if normalize_y else 0,
                                     ncap, ptr(buf), ptr(buf[nx:]), ptr(nh), ptr(ym), ptr(ys)))
    pb = PreparedBatch(buf, B, ncap, d, nh[:B].copy(), ym, ys)
    pb.pinned = pinned
    return pb


class FitHandle(object):
    """Result buffers of an enqueued ``tlbo_gp_fit_batched``; ``finish`` reads them back
    (one synchronisation)."""

    def __init__(self, status, rt, rf, ri, keep):
        self.status, self.rt, self.rf, self.ri, self._keep = status, rt, rf, ri, keep

    def finish(self, return_restarts=False):
        st = d2h(self.status)
        self._keep = None
        if return_restarts:
            return st, d2h(self.rt), d2h(self.rf), d2h(self.ri)
        return st


_CONST_CACHE = {}
_P0_CACHE = {}          # id(start-point array) -> (array, device copy, restarts)
_WS_BYTES = {}          # (B, R, ncap, d) -> tlbo_gp_fit_workspace_bytes


def _cached_const(a):
    """Small read-only device constants (start points, bounds): uploaded once per value."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    # per STREAM: an entry is uploaded on the stream that first asks for it, and a trial running on another stream
    # (several trials per process) must not consume it before that copy has completed there
    key = (a.shape, a.tobytes(), _stream().value)
    t = _CONST_CACHE.get(key)
    if t is None:
        if len(_CONST_CACHE) > 1024:
            _CONST_CACHE.clear()
        t = h2d(a)
        _CONST_CACHE[key] = t
    return t


def cached_upload(a):
    """Small host arrays that repeat from call to call (ensemble weights / skip flags / slot order
    within one maximisation): uploaded once per distinct value."""
    a = np.ascontiguousarray(a)
    key = (a.dtype.str, a.shape, a.tobytes(), _stream().value)
    t = _CONST_CACHE.get(key)
    if t is None:
        if len(_CONST_CACHE) > 1024:
            _CONST_CACHE.clear()
        t = h2d(a)
        _CONST_CACHE[key] = t
    return t


def gp_fit_batched(X_list, y_list, types, p0, lo, hi, ymean, ystd, pack, slot0=0, do_optimize=True,
                   return_restarts=False, sync=True, prepared=None):
    """``tlbo_gp_fit_batched``.  X_list/y_list: per-surrogate host arrays (y already
    normalised); p0 [R, H] start points; returns status int32[B] (host) and optionally
    the per-restart results -- or, with ``sync=False``, a ``FitHandle`` right after the
    launch so that host work can overlap the fit.  ``prepared``: a ``PreparedBatch`` (then
    X_list / y_list / ymean / ystd are taken from it)."""
    if prepared is not None:
        B, d, ncap = prepared.B, prepared.d, prepared.ncap
        ymean, ystd = prepared.ymean, prepared.ystd
    else:
        B = len(X_list)
        d = int(X_list[0].shape[1])
        ncap = batch_cap(max(x.shape[0] for x in X_list))
    H = d + 2
    if ncap > pack.ncap:
        raise ValueError('pack row capacity %d < %d' % (pack.ncap, ncap))
    if prepared is not None:
        nx, ny = B * ncap * d, B * ncap
        pinned = getattr(prepared, 'pinned', None)
        if pinned is not None and pinned.is_pinned():
            dbuf = pinned.to(_dev(), non_blocking=True)
            STATS.h2d += pinned.numel() * 8
        else:
            dbuf = h2d(prepared.buf)
        dX, dy, dn = dbuf[:nx].view(B, ncap, d), dbuf[nx:nx + ny].view(B, ncap), dbuf[nx + ny:].view(torch.int32)[:B]
    else:
        dX, dy, dn = _pad_batch(X_list, y_list, ncap, d)
    dev = dX.device
    # the start points of a facade are ONE array object reused by every fit (same seed -> same draws): its device copy
    # is looked up by identity, not by hashing a kilobyte of floats per launch
    hit = _P0_CACHE.get(id(p0))
    if hit is not None and hit[0] is p0:
        dp0, R = hit[1], hit[2]
    else:
        p0c = np.ascontiguousarray(np.asarray(p0, dtype=np.float64).reshape(-1, H))
        R = p0c.shape[0]
        dp0 = _cached_const(p0c)
        if len(_P0_CACHE) > 64:
            _P0_CACHE.clear()
        _P0_CACHE[id(p0)] = (p0, dp0, R)
    dlo = _cached_const(lo)
    dhi = _cached_const(hi)
    hym = np.ascontiguousarray(ymean, dtype=np.float64)
    hys = np.ascontiguousarray(ystd, dtype=np.float64)
    STATS.h2d += hym.nbytes + hys.nbytes
    wkey = (B, R, ncap, d)
    wbytes = _WS_BYTES.get(wkey)
    if wbytes is None:
        wbytes = _WS_BYTES[wkey] = int(lib.tlbo_gp_fit_workspace_bytes(B, R, ncap, d))
    # ONE allocation for every per-launch buffer: rt | rf | work | ri | status (the factorisation kernel writes the
    # status of every surrogate on all of its paths: no zero fill needed)
    o_rf = 8 * B * R * H
    o_work = round_up(o_rf + 8 * B * R, 256)
    o_ri = round_up(o_work + wbytes, 16)
    o_st = o_ri + 16 * B * R
    blob = torch.empty(o_st + 4 * B, dtype=torch.uint8, device=dev)
    rt = blob[:o_rf].view(F64).view(B, R, H)
    rf = blob[o_rf:o_rf + 8 * B * R].view(F64).view(B, R)
    work = blob[o_work:o_work + wbytes]
    ri = blob[o_ri:o_st].view(torch.int32).view(B, R, 4)
    status = blob[o_st:o_st + 4 * B].view(torch.int32)
    t = _types_arr(types)
    with _call('gp_fit', 2):            # gp_fit_kernel + gp_factorize_kernel
        check(lib.tlbo_gp_fit_batched(ptr(dX), ptr(dy), ptr(dn), B, ncap, d, ptr(t), ptr(dp0), R, ptr(dlo),
                                      ptr(dhi), ptr(hym), ptr(hys), 1 if do_optimize else 0, pack.ref(), slot0,
                                      ptr(rt), ptr(rf), ptr(ri), ptr(status), ptr(work), _stream()))
    h = FitHandle(status, rt, rf, ri, (dX, dy, dn, work, hym, hys))
    if not sync:
        return h
    return h.finish(return_restarts)


def gp_factorize_batched(X_list, y_list, types, thetas, ymean, ystd, pack, slot0=0, want_dense=False):
    """``tlbo_gp_factorize_batched``: factorise at given theta and fill pack slots."""
    B = len(X_list)
    d = int(X_list[0].shape[1])
    ncap = batch_cap(max(x.shape[0] for x in X_list))
    if ncap > pack.ncap:
        raise ValueError('pack row capacity %d < %d' % (pack.ncap, ncap))
    dX, dy, dn = _pad_batch(X_list, y_list, ncap, d)
    dev = dX.device
    dth = h2d(np.ascontiguousarray(np.asarray(thetas, dtype=np.float64).reshape(B, d + 2)))
    hym = np.ascontiguousarray(ymean, dtype=np.float64)
    hys = np.ascontiguousarray(ystd, dtype=np.float64)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    wbytes = int(lib.tlbo_gp_fit_workspace_bytes(B, 1, ncap, d))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    dL = torch.zeros(B, ncap, ncap, dtype=F64, device=dev) if want_dense else None
    dK = torch.zeros(B, ncap, ncap, dtype=F64, device=dev) if want_dense else None
    t = _types_arr(types)
    with _call('gp_factorize', 1):
        check(lib.tlbo_gp_factorize_batched(ptr(dX), ptr(dy), ptr(dn), B, ncap, d, ptr(t), ptr(dth), ptr(hym),
                                            ptr(hys), pack.ref(), slot0, ptr(dL), ptr(dK), ptr(status), ptr(work),
                                            _stream()))
    st = d2h(status)
    if want_dense:
        return st, d2h(dL), d2h(dK)
    return st


def gp_nll_grad_batched(X_list, y_list, types, thetas):
    """``tlbo_gp_nll_grad_batched``: one ``GaussianProcess._nll`` evaluation per surrogate."""
    B = len(X_list)
    d = int(X_list[0].shape[1])
    H = d + 2
    ncap = batch_cap(max(x.shape[0] for x in X_list))
    dX, dy, dn = _pad_batch(X_list, y_list, ncap, d)
    dev = dX.device
    dth = h2d(np.ascontiguousarray(np.asarray(thetas, dtype=np.float64).reshape(B, H)))
    nll = torch.zeros(B, dtype=F64, device=dev)
    grad = torch.zeros(B, H, dtype=F64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    wbytes = int(lib.tlbo_gp_fit_workspace_bytes(B, 1, ncap, d))
    work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    t = _types_arr(types)
    with _call('gp_nll_grad', 1):
        check(lib.tlbo_gp_nll_grad_batched(ptr(dX), ptr(dy), ptr(dn), B, ncap, d, ptr(t), ptr(dth), ptr(nll),
                                           ptr(grad), ptr(status), ptr(work), _stream()))
    return d2h(nll), d2h(grad), d2h(status)


def gp_predict(pack, types, Xq, out=None):
    """``tlbo_gp_predict``: per-surrogate mean / variance at a small query set.
    Xq: device tensor (or pinned host tensor, read over UVA) or host array [Nq, d].  Returns device
    tensors mu, var [M, Nq] (``out`` = preallocated contiguous buffers with at least that many elements)."""
    dev = _dev()
    if not torch.is_tensor(Xq):
        Xq = h2d(np.ascontiguousarray(Xq, dtype=np.float64))
    Nq = int(Xq.shape[0])
    if out is not None:
        mu = out[0].view(-1)[:pack.M * Nq].view(pack.M, Nq)
        var = out[1].view(-1)[:pack.M * Nq].view(pack.M, Nq)
    else:
        mu = torch.empty(pack.M, Nq, dtype=F64, device=dev)
        var = torch.empty(pack.M, Nq, dtype=F64, device=dev)
    if Nq == 0:
        return mu, var
    t = _types_arr(types)
    with _call('gp_predict_small', 1):
        check(lib.tlbo_gp_predict(pack.ref(), ptr(t), ptr(Xq), Nq, ptr(mu), ptr(var), _stream()))
    return mu, var


class FusedWorkspace(object):
    """Reusable scratch + result buffers of ``tlbo_predict_ei_fused`` (``pinned_best``: the 4-double
    result lands in mapped pinned host memory, readable after a stream synchronisation)."""

    def __init__(self, pinned_best=False):
        dev = _dev()
        self.work = torch.empty(int(lib.tlbo_predict_ei_workspace_bytes(0)), dtype=torch.uint8, device=dev)
        self.best = torch.zeros(4, dtype=F64).pin_memory() if pinned_best else torch.zeros(4, dtype=F64, device=dev)


class SmallQuery(object):
    """Buffers of the low-overhead acquisition call on a small ad-hoc query set (the online
    scenario's neighbourhoods): the kernel reads the rows from a mapped pinned staging buffer and EI comes
    back through mapped pinned host memory (UVA loads / stores), so one call is ONE launch and one stream
    synchronisation (or a spin on the EI rows) -- no memcpy, no allocation."""

    def __init__(self, cap, d, M):
        dev = _dev()
        self.cap = int(cap)
        self.X = torch.zeros(self.cap, d, dtype=F64).pin_memory()
        self.X_np = self.X.numpy()
        self.Xd = torch.zeros(self.cap, d, dtype=F64, device=dev)
        self.ei = torch.zeros(self.cap, dtype=F64).pin_memory()
        self.ei_np = self.ei.numpy()
        self.mu_rows = torch.zeros(self.cap, dtype=F64, device=dev)
        self.var_rows = torch.zeros(self.cap, dtype=F64, device=dev)
        self.mu = torch.zeros(M * self.cap, dtype=F64, device=dev)
        self.var = torch.zeros(M * self.cap, dtype=F64, device=dev)
        self.ws = FusedWorkspace(pinned_best=True)
        self.best_np = self.ws.best.numpy()
        self.counter = torch.zeros(self.cap, dtype=torch.int32, device=dev)       # tlbo_ensemble_ei_small's tickets


def ensemble_ei_small(view, types, w, skip, order, mode, sq, N, eta, par, var_floor):
    """``tlbo_ensemble_ei_small`` over the first ``N`` rows of ``sq.X`` (pinned host staging; small sets travel
    inside the kernel arguments): ONE launch, EI (NaN kept) in ``sq.ei_np[:N]`` once the stream is synchronised."""
    t = _types_arr(types)
    with _call('ensemble_ei_small', 1):
        check(lib.tlbo_ensemble_ei_small(view.ref(), ptr(t), ptr(w), ptr(skip), ptr(order), int(mode), ptr(sq.X), ptr(sq.X),
                                         int(N), float(eta), float(par), float(var_floor), ptr(sq.mu), ptr(sq.var),
                                         ptr(sq.ei), ptr(sq.counter), _stream()))


def predict_ei_fused(pack, types, w, skip, cand, eta, par=0.0, var_floor=1e-10, keys=None, excluded=None,
                     idx_offset=0, mode=0, order=None, want_rows=False, ws=None, sync=True, nmax=0, cache=None,
                     rows=None):
    """``tlbo_predict_ei_fused``.  All array arguments are DEVICE tensors (cand [N, d],
    w [M], skip int32[M] or None, keys [N] or None, excluded sorted int64 or None).
    Returns (ei, key, idx, n_negative) of the best non-excluded row (host floats when
    ``sync``), plus device mu/var/ei [N] when ``want_rows``."""
    dev = cand.device
    N = int(cand.shape[0])
    ws = ws or FusedWorkspace()
    mu = var = ei = None
    if rows is not None:
        mu, var, ei = rows            # caller-provided per-row outputs (device or pinned host, >= N elements)
    elif want_rows:
        mu = torch.empty(N, dtype=F64, device=dev)
        var = torch.empty(N, dtype=F64, device=dev)
        ei = torch.empty(N, dtype=F64, device=dev)
    n_excl = 0 if excluded is None else int(excluded.numel())
    t = _types_arr(types)
    pack.c.nmax = int(nmax) if nmax and nmax > 0 else 0
    with _call('predict_ei_fused', 2):    # predict_ei_fused_kernel + argmax_final_kernel
        if cache is None:
            check(lib.tlbo_predict_ei_fused(pack.ref(), ptr(t), ptr(w), ptr(skip), ptr(order), int(mode), ptr(cand), N,
                                            float(eta), float(par), float(var_floor), ptr(keys),
                                            ptr(excluded) if n_excl else None, n_excl, int(idx_offset), ptr(ws.best),
                                            ptr(mu), ptr(var), ptr(ei), ptr(ws.work), _stream()))
        else:
            cmu, cvar = cache
            assert cmu.shape[1] == N and cvar.shape == cmu.shape and cmu.is_contiguous() and cvar.is_contiguous()
            check(lib.tlbo_predict_ei_fused_cached(pack.ref(), ptr(t), ptr(w), ptr(skip), ptr(order), int(mode),
                                                   ptr(cand), N, float(eta), float(par), float(var_floor), ptr(keys),
                                                   ptr(excluded) if n_excl else None, n_excl, int(idx_offset),
                                                   ptr(ws.best), ptr(mu), ptr(var), ptr(ei), ptr(cmu), ptr(cvar),
                                                   int(cmu.shape[0]), ptr(ws.work), _stream()))
    if not sync:
        return ws.best, mu, var, ei
    b = d2h(ws.best)
    res = (float(b[0]), float(b[1]), int(b[2]), int(b[3]))
    if want_rows or rows is not None:
        return res, mu, var, ei
    return res


def predict_pool_posteriors(pack, types, cand, nmax=0, ws=None):
    """``tlbo_predict_pool_posteriors``: per-surrogate posterior mean / variance of every slot of
    ``pack`` over a resident candidate pool; returns device tensors ``mu, var [M, N]``."""
    dev = cand.device
    N = int(cand.shape[0])
    ws = ws or FusedWorkspace()
    mu = torch.empty(pack.M, N, dtype=F64, device=dev)
    var = torch.empty(pack.M, N, dtype=F64, device=dev)
    t = _types_arr(types)
    pack.c.nmax = int(nmax) if nmax and nmax > 0 else 0
    with _call('predict_pool_posteriors', 2):
        check(lib.tlbo_predict_pool_posteriors(pack.ref(), ptr(t), ptr(cand), N, ptr(mu), ptr(var), ptr(ws.work),
                                               _stream()))
    return mu, var


def rank_loss_counts(y, ys, row_lo, row_hi, upper_only=False):
    """``tlbo_rank_loss_counts``.  y [n], ys [S, Mr, n] host arrays or device tensors;
    returns int64[S, Mr] on the host."""
    dev = _dev()

    def up(a, dt):
        if torch.is_tensor(a) and a.is_cuda:
            return a.to(dtype=dt).contiguous()
        return h2d(a, dt)
    dy = up(y, F64)
    dys = up(ys, F64)
    S, Mr, n = dys.shape
    dlo = up(np.asarray(row_lo, dtype=np.int32), torch.int32)
    dhi = up(np.asarray(row_hi, dtype=np.int32), torch.int32)
    counts = torch.zeros(S, Mr, dtype=torch.int64, device=dev)
    with _call('rank_loss_counts', 1):
        check(lib.tlbo_rank_loss_counts(ptr(dy), ptr(dys), int(n), int(S), int(Mr), ptr(dlo), ptr(dhi),
                                        1 if upper_only else 0, ptr(counts), _stream()))
    return d2h(counts)


def rank_loss_counts_normal(y, z, loc, scale, src, row_lo, row_hi, sync=True):
    """``tlbo_rank_loss_counts_normal``: z [S, Mr, n] standard-normal draws (host), loc / scale
    device tensors [*, n] (outputs of ``gp_predict``), src[Mr] their row per column.  Returns
    int64[S, Mr] on the host, or the device tensor when ``sync`` is False."""
    dev = _dev()

    def up(a, dt):
        if torch.is_tensor(a) and a.is_cuda:
            return a
        return h2d(np.ascontiguousarray(a, dtype=dt))
    dy = up(y, np.float64)
    dz = up(z, np.float64)
    S, Mr, n = dz.shape
    dsrc = up(src, np.int32)
    dlo = up(row_lo, np.int32)
    dhi = up(row_hi, np.int32)
    assert loc.is_contiguous() and scale.is_contiguous() and loc.shape[1] == n
    counts = torch.zeros(S, Mr, dtype=torch.int64, device=dev)
    with _call('rank_loss_counts', 1):
        check(lib.tlbo_rank_loss_counts_normal(ptr(dy), ptr(dz), ptr(loc), ptr(scale), ptr(dsrc), int(n), int(S),
                                               int(Mr), ptr(dlo), ptr(dhi), ptr(counts), _stream()))
    if not sync:
        return counts
    return d2h(counts)


def rgpe_weights(counts, K, n_folds, n_sq, out):
    """``tlbo_rgpe_weights``: RGPE weights and weight-dilution flags from the device pair-count table
    ``counts int64[S, K + n_folds]`` into ``out = (w float64[K + 1], skip int32[K + 1])`` (device tensors)."""
    w, skip = out
    S = int(counts.shape[0])
    assert counts.is_contiguous() and int(counts.shape[1]) == K + n_folds and w.numel() == K + 1 == skip.numel()
    with _call('rgpe_weights', 1):
        check(lib.tlbo_rgpe_weights(ptr(counts), S, int(K), int(n_folds), int(n_sq), ptr(w), ptr(skip), _stream()))
    return w, skip


def pair_penalty_loss_normal(y, z, loc, scale, sync=True):
    """``tlbo_pair_penalty_loss_normal`` (OBTL's sampled ranking loss): y [n], z [S, Mr, n] device or host,
    loc / scale [Mr, n] device -> loss float64[S, Mr]."""
    dev = _dev()

    def up(a):
        return a if torch.is_tensor(a) and a.is_cuda else h2d(np.ascontiguousarray(a, dtype=np.float64))
    y, z, loc, scale = up(y), up(z), up(loc), up(scale)
    S, Mr, n = (int(v) for v in z.shape)
    assert loc.shape == (Mr, n) and scale.shape == (Mr, n) and y.numel() == n
    loss = torch.empty(S, Mr, dtype=F64, device=dev)
    with _call('pair_penalty_loss', 1):
        check(lib.tlbo_pair_penalty_loss_normal(ptr(y), ptr(z.contiguous()), ptr(loc.contiguous()),
                                                ptr(scale.contiguous()), n, S, Mr, ptr(loss), _stream()))
    return d2h(loss) if sync else loss


class DeviceRandomState(object):
    """A legacy ``np.random.RandomState`` whose MT19937 state lives on the device: ``rand_into(out)`` fills a
    device tensor with the doubles ``rng.rand(out.numel())`` would return, bit for bit, and advances the
    state (``tlbo_mt19937_random_sample``).  ``to_host(rng)`` / ``from_host(rng)`` move the state.

    ``slices > 1`` (chosen from ``n_hint``, the length of the draws to come): the stream is produced by that many
    thread blocks at once -- slice c starts from the state jumped ahead by c * S doubles (``tlbo_mt19937_jump``,
    polynomials from ``tlbo_b200.utils.mt_jump``, computed once) -- and the jumped states of the NEXT draw are
    prepared right behind the current one, on the same stream."""

    SLICE_MIN = 65536            # doubles per slice below which more slices do not pay
    SLICE_MAX_COUNT = 64

    def __init__(self, rng, n_hint=0):
        self.state = torch.empty(625, dtype=torch.int32, device=_dev())
        self._slices, self._S, self._n = 1, 0, 0
        n_hint = int(n_hint)
        C = min(int(os.environ.get('TLBO_KEY_SLICES_MAX', self.SLICE_MAX_COUNT)), n_hint // self.SLICE_MIN)
        if C >= 2:
            from tlbo_b200.utils import mt_jump
            S = round_up((n_hint + C - 1) // C, 312)
            C = (n_hint + S - 1) // S
            if C >= 2:
                g1 = mt_jump.jump_poly(2 * S)                      # 2 words per double
                polys, g = [1], 1
                phi = mt_jump.charpoly()
                for _ in range(1, C):
                    g = mt_jump._mulmod(g, g1, phi)
                    polys.append(g)
                pos = [mt_jump.bit_positions(p) for p in polys]
                stride = max(len(p) for p in pos)
                tab = np.zeros((C, stride), dtype=np.int32)
                for c, p in enumerate(pos):
                    tab[c, :len(p)] = p
                self._bitpos = torch.from_numpy(tab).to(_dev())
                self._nbits = torch.from_numpy(np.array([len(p) for p in pos], dtype=np.int32)).to(_dev())
                self._stride = int(stride)
                self._states = torch.empty(C, 625, dtype=torch.int32, device=_dev())
                self._slices, self._S, self._n = int(C), int(S), n_hint
        self.from_host(rng)

    def _jump(self):
        with _call('mt19937_jump', 1):
            check(lib.tlbo_mt19937_jump(ptr(self.state), ptr(self._states), self._slices, ptr(self._bitpos),
                                        ptr(self._nbits), self._stride, _stream()))

    def from_host(self, rng):
        st = rng.get_state()
        assert st[0] == 'MT19937'
        h = np.empty(625, dtype=np.uint32)
        h[:624] = st[1]
        h[624] = st[2]
        self.state.copy_(torch.from_numpy(h.view(np.int32)))
        if self._slices > 1:
            self._jump()
        torch.cuda.current_stream().synchronize()      # the generator runs on a side stream
        self._aux = (st[3], st[4])

    def to_host(self, rng):
        h = self.state.cpu().numpy().view(np.uint32)
        rng.set_state(('MT19937', h[:624].copy(), int(h[624]), self._aux[0], self._aux[1]))

    def rand_into(self, out):
        assert out.is_cuda and out.dtype == F64 and out.is_contiguous()
        N = int(out.numel())
        if self._slices > 1 and N == self._n:
            with _call('mt19937_random_sample', 1):
                check(lib.tlbo_mt19937_random_sample_sliced(ptr(self._states), self._slices, N, self._S, ptr(out),
                                                            ptr(self.state), _stream()))
            self._jump()                               # the next draw's slice states, behind this one
            return out
        with _call('mt19937_random_sample', 1):
            check(lib.tlbo_mt19937_random_sample(ptr(self.state), N, ptr(out), _stream()))
        if self._slices > 1:
            self._jump()
        return out


class PairLogistic(object):
    """Pairwise logistic ranking loss + gradient with A and y resident on the device across
    the SLSQP callbacks of one ``scipy_solve``.  Per callback: ONE launch with the weights as
    a kernel argument, the result written straight into mapped pinned host memory, one stream
    synchronisation (no host<->device memcpy)."""

    def __init__(self, A, y):
        _dev()
        A = np.ascontiguousarray(A, dtype=np.float64)
        self.n, self.m = A.shape
        self.dA = h2d(A)
        self.dy = h2d(np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1)))
        self.host_w = self.m <= 64
        self.dw = None if self.host_w else torch.zeros(self.m, dtype=F64, device=self.dA.device)
        self.out = torch.zeros(2 + self.m, dtype=F64).pin_memory()      # device-visible (UVA)
        self.out_np = self.out.numpy()
        self._cache_w = None
        self._cache = None
        # the launching stream is fixed for the lifetime of one scipy_solve (looking it up costs
        # ~20 us per SLSQP callback)
        self._stream = torch.cuda.current_stream()
        self._sptr = ctypes.c_void_p(self._stream.cuda_stream)
        self._pA, self._py, self._pout = ptr(self.dA), ptr(self.dy), ptr(self.out)

    def __call__(self, w):
        w = np.ascontiguousarray(np.asarray(w, dtype=np.float64).reshape(-1))
        if self._cache_w is not None and np.array_equal(w, self._cache_w):
            return self._cache
        stream, sptr = self._stream, self._sptr
        with _call('pairwise_logistic', 1):
            if self.host_w:
                # launch + spin on the mapped result words inside ONE C call (no stream synchronisation)
                check(lib.tlbo_pairwise_logistic_loss_grad_hw_wait(self._pA, self._py, ptr(w), self.n, self.m,
                                                                   self._pout, sptr))
            else:
                self.dw.copy_(torch.from_numpy(w))
                check(lib.tlbo_pairwise_logistic_loss_grad(ptr(self.dA), ptr(self.dy), ptr(self.dw), self.n, self.m,
                                                           ptr(self.out), sptr))
                stream.synchronize()
        STATS.h2d += w.nbytes
        STATS.d2h += self.out_np.nbytes
        o = self.out_np
        self._cache_w = w.copy()
        self._cache = (float(o[0]), int(o[1]), o[2:].copy())
        return self._cache


# ---- MCMC-marginalised GP (BASELINE config 5) -------------------------------------------

class McmcTables(object):
    """Device copy of the stretch-move random tables (``stretch_move_tables``) of one or more
    samplers, stacked ``[n_tables, 2 * nsteps, W / 2]``.  Uploads are cached by identity of
    the host tables: every model a facade builds shares one table."""
    _cache = {}

    def __init__(self, tables):
        key = tuple(id(t) for t in tables) + (torch.cuda.current_device(),)
        hit = McmcTables._cache.get(key)
        if hit is None:
            if len(McmcTables._cache) > 16:
                McmcTables._cache.clear()
            dev = [h2d(np.ascontiguousarray(np.stack([t[i] for t in tables]))) for i in range(5)]
            hit = (list(tables), dev)          # the host tables stay alive, so the ids stay unique
            McmcTables._cache[key] = hit
        self.sidx, self.cidx, self.zz, self.fac, self.logu = hit[1]
        self.n_tables = len(tables)
        self.nhalf, self.Ns = int(self.sidx.shape[1]), int(self.sidx.shape[2])


class McmcHandle(object):
    """Buffers of an enqueued ``tlbo_gp_mcmc_chain``; ``finish`` reads them back."""

    def __init__(self, pos, lp, naccept, status, batch, keep):
        self.pos, self.lp, self.naccept, self.status, self.batch, self._keep = pos, lp, naccept, status, batch, keep

    def finish(self):
        st = d2h(self.status)
        self._keep = None
        return st, d2h(self.pos), d2h(self.lp), d2h(self.naccept)


def upload_batch(X_list, y_list):
    """Padded device batch ``(dX [B, ncap, d], dy [B, ncap], dn [B], ncap)`` (one upload)."""
    d = int(X_list[0].shape[1])
    ncap = batch_cap(max(x.shape[0] for x in X_list))
    dX, dy, dn = _pad_batch(X_list, y_list, ncap, d)
    return dX, dy, dn, ncap


def gp_mcmc_chain(batch, types, W, nsteps, pos0, tables, table_of, prior_bounds_raw, theta0):
    """``tlbo_gp_mcmc_chain``: ``nsteps`` emcee stretch-move steps of ``W`` walkers for each of
    the B surrogates of ``batch`` (``upload_batch``) in one persistent launch (per group of
    co-resident surrogates).  ``pos0`` host ``[B, W, H]`` initial ensembles; ``tables`` a
    ``McmcTables`` (or None when ``nsteps == 0``: one batched evaluation of ``_ll``).
    Returns a ``McmcHandle`` right after the launch."""
    dX, dy, dn, ncap = batch
    B, d = int(dX.shape[0]), int(dX.shape[2])
    H = d + 2
    dev = dX.device
    pos = h2d(np.ascontiguousarray(np.asarray(pos0, dtype=np.float64).reshape(B, W, H)))
    lp = torch.empty(B, W, dtype=F64, device=dev)
    nacc = torch.zeros(B, W, dtype=torch.int32, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    pb = np.asarray(prior_bounds_raw, dtype=np.float64).reshape(H, 2)
    dlo = _cached_const(pb[:, 0])
    dhi = _cached_const(pb[:, 1])
    dth0 = None if theta0 is None else _cached_const(theta0)
    dtab = None
    if nsteps > 0:
        assert tables is not None and tables.nhalf == 2 * nsteps and tables.Ns == W // 2
        if table_of is not None and tables.n_tables > 1:
            dtab = h2d(np.ascontiguousarray(table_of, dtype=np.int32))
    work = torch.empty(int(lib.tlbo_gp_mcmc_workspace_bytes(B, W, ncap, d)), dtype=torch.uint8, device=dev)
    t = _types_arr(types)
    tb = tables
    with _call('gp_mcmc_chain', 1):
        check(lib.tlbo_gp_mcmc_chain(ptr(dX), ptr(dy), ptr(dn), B, ncap, d, ptr(t), int(W), int(nsteps), ptr(pos),
                                     ptr(lp), ptr(tb.sidx) if nsteps else None, ptr(tb.cidx) if nsteps else None,
                                     ptr(tb.zz) if nsteps else None, ptr(tb.fac) if nsteps else None,
                                     ptr(tb.logu) if nsteps else None, ptr(dtab), ptr(dlo), ptr(dhi), ptr(dth0),
                                     ptr(nacc), ptr(status), ptr(work), _stream()))
    return McmcHandle(pos, lp, nacc, status, batch, (work, dtab, tb))


def gp_factorize_walkers(batch, types, theta_dev, ymean, ystd, pack, slot0, W):
    """Factorise ``W`` walker GPs per surrogate of ``batch`` at the DEVICE thetas
    ``theta_dev [B * W, H]`` into pack slots ``slot0 ...`` (``tlbo_gp_factorize_batched`` over
    the replicated batch).  Returns the device status tensor ``int32[B * W]`` (no sync)."""
    dX, dy, dn, ncap = batch
    B, d = int(dX.shape[0]), int(dX.shape[2])
    if ncap > pack.ncap:
        raise ValueError('pack row capacity %d < %d' % (pack.ncap, ncap))
    rX = dX.repeat_interleave(W, dim=0).contiguous()
    ry = dy.repeat_interleave(W, dim=0).contiguous()
    rn = dn.repeat_interleave(W, dim=0).contiguous()
    hym = np.ascontiguousarray(np.repeat(np.asarray(ymean, dtype=np.float64), W))
    hys = np.ascontiguousarray(np.repeat(np.asarray(ystd, dtype=np.float64), W))
    status = torch.zeros(B * W, dtype=torch.int32, device=dX.device)
    work = torch.empty(int(lib.tlbo_gp_fit_workspace_bytes(B * W, 1, ncap, d)), dtype=torch.uint8, device=dX.device)
    t = _types_arr(types)
    theta_dev = theta_dev.contiguous()
    with _call('gp_factorize', 1):
        check(lib.tlbo_gp_factorize_batched(ptr(rX), ptr(ry), ptr(rn), B * W, ncap, d, ptr(t), ptr(theta_dev), ptr(hym),
                                            ptr(hys), pack.ref(), int(slot0), None, None, ptr(status), ptr(work),
                                            _stream()))
    status._keep = (rX, ry, rn, work, hym, hys, theta_dev)
    return status


def mcmc_marginalize(mu, var, G, W, out=None):
    """``tlbo_mcmc_marginalize``: device ``mu, var [G * W, N]`` -> ``gmu, gvar [G, N]`` (written
    into ``out = (gmu, gvar)`` when given)."""
    assert mu.shape == var.shape and mu.shape[0] == G * W and mu.is_contiguous() and var.is_contiguous()
    N = int(mu.shape[1])
    if out is not None:
        gmu, gvar = out
        assert gmu.shape == (G, N) and gvar.shape == (G, N) and gmu.is_contiguous() and gvar.is_contiguous()
    else:
        gmu = torch.empty(G, N, dtype=F64, device=mu.device)
        gvar = torch.empty(G, N, dtype=F64, device=mu.device)
    if N:
        with _call('mcmc_marginalize', 1):
            check(lib.tlbo_mcmc_marginalize(ptr(mu), ptr(var), int(G), int(W), N, ptr(gmu), ptr(gvar), _stream()))
    return gmu, gvar


def fp64_peak_probe():
    """(DFMA TFLOP/s, DMMA TFLOP/s) measured on the current device."""
    _dev()
    out = (ctypes.c_double * 2)()
    check(lib.tlbo_fp64_peak_probe(ctypes.cast(out, ctypes.c_void_p), _stream()))
    return float(out[0]), float(out[1])


def set_fit_variant(variant):
    """``tlbo_set_fit_variant``: 1 = warp-cooperative L-BFGS-B (default), 0 = single-thread
    statement of the same optimiser.  Returns the previous value."""
    return int(lib.tlbo_set_fit_variant(int(variant)))
