"""ctypes binding of ``include/tlbo_b200.h`` (the C ABI of the CUDA hot path).

There is NO CPU fallback: importing this module without the built library raises,
and every compute entry point needs a CUDA device.  Build with
``python -m tlbo_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('TLBO_B200_LIB') or os.path.join(os.path.dirname(_HERE), 'lib', 'libtlbo_b200.so')

HYP_STRIDE = 32
MAX_D = 22


def ldl(ncap):
    """TLBO_LDL: row stride (doubles) of the packed L^-1 matrices."""
    return ncap + 4


class TlboPack(ctypes.Structure):
    """``struct tlbo_pack`` (include/tlbo_b200.h)."""
    _fields_ = [
        ('M', c_int32), ('ncap', c_int32), ('d', c_int32), ('nmax', c_int32),
        ('d_n', c_void_p), ('d_Xs', c_void_p), ('d_hyp', c_void_p), ('d_alpha', c_void_p),
        ('d_Linv', c_void_p), ('d_theta', c_void_p),
    ]


# every symbol the header declares: name -> (restype, argtypes)
_P = c_void_p
SYMBOLS = {
    'tlbo_last_error': (c_char_p, []),
    'tlbo_version': (c_int, []),
    'tlbo_stream_synchronize': (c_int, [_P]),
    'tlbo_gp_fit_batched': (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P, _P, _P, c_int,
                                    POINTER(TlboPack), c_int, _P, _P, _P, _P, _P, _P]),
    'tlbo_gp_fit_workspace_bytes': (c_int64, [c_int, c_int, c_int, c_int]),
    'tlbo_set_fit_variant': (c_int, [c_int]),
    'tlbo_gp_factorize_batched': (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, POINTER(TlboPack),
                                          c_int, _P, _P, _P, _P, _P]),
    'tlbo_gp_nll_grad_batched': (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    'tlbo_gp_predict': (c_int, [POINTER(TlboPack), _P, _P, c_int, _P, _P, _P]),
    'tlbo_predict_ei_fused': (c_int, [POINTER(TlboPack), _P, _P, _P, _P, c_int, _P, c_int64, c_double, c_double,
                                      c_double, _P, _P, c_int, c_int64, _P, _P, _P, _P, _P, _P]),
    'tlbo_predict_ei_workspace_bytes': (c_int64, [c_int64]),
    'tlbo_merge_best': (c_int, [_P, c_int, _P, _P]),
    'tlbo_predict_pool_posteriors': (c_int, [POINTER(TlboPack), _P, _P, c_int64, _P, _P, _P, _P]),
    'tlbo_predict_ei_fused_cached': (c_int, [POINTER(TlboPack), _P, _P, _P, _P, c_int, _P, c_int64, c_double, c_double,
                                             c_double, _P, _P, c_int, c_int64, _P, _P, _P, _P, _P, _P, c_int, _P, _P]),
    'tlbo_rank_loss_counts': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_int, _P, _P]),
    'tlbo_rank_loss_counts_normal': (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    'tlbo_pairwise_logistic_loss_grad': (c_int, [_P, _P, _P, c_int, c_int, _P, _P]),
    'tlbo_pairwise_logistic_loss_grad_hw': (c_int, [_P, _P, _P, c_int, c_int, _P, _P]),
    'tlbo_pairwise_logistic_loss_grad_hw_wait': (c_int, [_P, _P, _P, c_int, c_int, _P, _P]),
    'tlbo_gp_mcmc_chain': (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P,
                                   _P, _P, _P, _P, _P, _P, _P]),
    'tlbo_gp_mcmc_workspace_bytes': (c_int64, [c_int, c_int, c_int, c_int]),
    'tlbo_mcmc_marginalize': (c_int, [_P, _P, c_int, c_int, c_int64, _P, _P, _P]),
    'tlbo_one_exchange_neighbourhood': (c_int, [_P, c_int, _P, _P, ctypes.c_uint32, c_int, c_double, _P, c_int, _P]),
    'tlbo_local_search_climb': (c_int, [POINTER(TlboPack), _P, _P, _P, _P, _P, c_int, c_double, c_double, c_double,
                                        c_int, c_double, _P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P]),
    'tlbo_ensemble_ei_small': (c_int, [POINTER(TlboPack), _P, _P, _P, _P, c_int, _P, _P, c_int, c_double, c_double,
                                       c_double, _P, _P, _P, _P, _P]),
    'tlbo_rgpe_weights': (c_int, [_P, c_int, c_int, c_int, c_int64, _P, _P, _P]),
    'tlbo_pair_penalty_loss_normal': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    'tlbo_mt19937_random_sample': (c_int, [_P, c_int64, _P, _P]),
    'tlbo_mt19937_jump': (c_int, [_P, _P, c_int, _P, _P, c_int, _P]),
    'tlbo_mt19937_random_sample_sliced': (c_int, [_P, c_int, c_int64, c_int64, _P, _P, _P]),
    'tlbo_host_prepare_jobs': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, c_int, _P, _P, _P, _P,
                                       _P]),
    'tlbo_fp64_peak_probe': (c_int, [_P, _P]),
}


class TlboError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            'tlbo_b200: the CUDA library %s is missing. Build it with `python -m tlbo_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        msg = lib.tlbo_last_error()
        raise TlboError('tlbo_b200 C-ABI call failed (rc=%d): %s' % (rc, msg.decode() if msg else ''))


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return c_void_p(t.data_ptr())
    return c_void_p(t.ctypes.data)
