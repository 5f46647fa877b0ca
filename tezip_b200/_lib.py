"""ctypes binding of libtezip_b200.so (include/tezip_b200.h).  There is no CPU fallback: if the library is
missing or a call fails, TezipError is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtezip_b200.so")

TZ_MAX_LAYERS = 8
TZ_HIST_BINS = 4096
TZ_SYMBOL_OFFSET = 1600
TZ_PREDNET_FP32_DIRECT = 1
TZ_ABI_VERSION = 2
TZ_WIDE_OFFSET = 400000
TZ_WIDE_SYM_MIN = TZ_WIDE_OFFSET - 131071
TZ_WIDE_BINS = 262144
MODES = {"abs": 0, "rel": 1, "absrel": 2, "pwrel": 3}

c_vp, c_int, c_ll, c_dbl, c_flt = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_float
c_ull, c_uint = ctypes.c_ulonglong, ctypes.c_uint


class TezipError(RuntimeError):
    pass


class PrednetConfig(ctypes.Structure):
    _fields_ = [("n_layers", c_int), ("stack_sizes", c_int * TZ_MAX_LAYERS), ("r_stack_sizes", c_int * TZ_MAX_LAYERS),
                ("Hp", c_int), ("Wp", c_int), ("pixel_max", c_flt), ("max_batch", c_int), ("device", c_int),
                ("flags", c_int)]


# every symbol include/tezip_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tz_abi_version": (c_int, []),
    "tz_last_error": (ctypes.c_char_p, []),
    "tz_device_count": (c_int, []),
    "tz_launch_count": (c_ll, []),
    "tz_prednet_create": (c_int, [ctypes.POINTER(PrednetConfig), ctypes.POINTER(c_vp), ctypes.POINTER(c_ll), c_int,
                                  ctypes.POINTER(c_vp)]),
    "tz_prednet_destroy": (c_int, [c_vp]),
    "tz_prednet_p0": (c_int, [c_vp, c_vp, c_vp]),
    "tz_prednet_next": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "tz_prednet_next_chained": (c_int, [c_vp, c_vp, c_int, c_vp]),
    "tz_prednet_kernel_count": (c_int, [c_vp]),
    "tz_prednet_kernel_info": (c_int, [c_vp, c_int, ctypes.c_char_p, c_int, ctypes.POINTER(c_dbl)]),
    "tz_prednet_next_timed": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, ctypes.POINTER(c_flt), c_int]),
    "tz_prednet_device_bytes": (c_ll, [c_vp]),
    "tz_prednet_flops_per_frame": (c_dbl, [c_vp]),
    "tz_pad_normalize": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "tz_residual": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "tz_error_bound": (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_vp]),
    "tz_delta_hist": (c_int, [c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_vp]),
    "tz_last_residual": (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "tz_build_table": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_delta_rank": (c_int, [c_vp, c_ll, c_int, c_vp, c_vp, c_vp, c_vp]),
    "tz_encode_lossless": (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int,
                                   c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_encode_lossy_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "tz_encode_lossy": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_dbl,
                                c_dbl, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_pad_normalize16": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "tz_residual16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "tz_last_residual16": (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "tz_error_bound16": (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_vp]),
    "tz_encode16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int,
                            c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_build_table16_workspace_bytes": (c_ll, []),
    "tz_build_table16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_reconstruct16_workspace_bytes": (c_ll, [c_ll]),
    "tz_reconstruct16": (c_int, [c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp,
                                 c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_window_sse16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "tz_reconstruct_workspace_bytes": (c_ll, [c_ll]),
    "tz_reconstruct": (c_int, [c_vp, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp,
                               c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_window_sse": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "tz_dwp_gather": (c_int, [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp]),
    "tz_dwp_update": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_dbl, c_int, c_dbl,
                              c_int, c_int, c_vp]),
    "tz_key_plane": (c_int, [c_vp, c_vp, c_vp, c_ll, c_ll, c_vp]),
    "tz_frames_nonzero": (c_int, [c_vp, c_vp, c_ll, c_ll, c_vp]),
    "tz_memcpy2d_async": (c_int, [c_vp, c_ll, c_vp, c_ll, c_ll, c_ll, c_vp]),
    "tz_zstd_bound": (c_ull, [c_ull]),
    "tz_zstd_workspace_bytes": (c_ull, [c_ull]),
    "tz_zstd_hist": (c_int, [c_vp, c_ull, c_vp, c_vp, c_vp]),
    "tz_zstd_encode": (c_int, [c_vp, c_ull, c_vp, c_vp, c_uint, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tz_zstd_decode": (c_int, [c_vp, c_vp, c_ull, c_vp, c_vp, c_vp, c_vp]),
}

_lib = None


def load():
    """Loads the CUDA extension; raises TezipError (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TezipError("%s is missing: build it with `python -m tezip_b200.build` "
                         "(hand-written sm_100a CUDA; there is no CPU fallback)" % LIB_PATH)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as e:
        raise TezipError("cannot load %s: %s" % (LIB_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.tz_abi_version() != TZ_ABI_VERSION:
        raise TezipError("ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().tz_last_error()
        raise TezipError("%s failed (%d): %s" % (what or "tezip_b200 call", rc, msg.decode("utf-8", "replace")))


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())
