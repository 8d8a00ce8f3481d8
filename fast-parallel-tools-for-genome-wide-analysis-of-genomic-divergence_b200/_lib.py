"""ctypes binding of libfpt_b200.so (C ABI: include/fpt_b200.h). Fails loudly — there is no fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfpt_b200.so")

FPT_OK = 0
FPT_ERR_CUDA, FPT_ERR_ARG, FPT_ERR_POSITIONS, FPT_ERR_WINDOW_TOO_LARGE, FPT_ERR_NO_DEVICE = -1, -2, -3, -4, -5
FPT_SCAN_SERIAL, FPT_SCAN_THREADED = 0, 1
FPT_WIN_EMPTY, FPT_WIN_DISCARDED, FPT_WIN_SCORED = 0, 1, 2


class FptError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libfpt_b200 error %d: %s" % (code, message))
        self.code = code


class Genotypes(C.Structure):
    _fields_ = [("avals", C.c_void_p), ("bvals", C.c_void_p), ("acodes", C.c_void_p), ("bcodes", C.c_void_p),
                ("pos", C.c_void_p), ("nsnp", C.c_int64), ("asize", C.c_int), ("bsize", C.c_int)]


class ScanRange(C.Structure):
    _fields_ = [("regend", C.c_int), ("wsize", C.c_int), ("wstep", C.c_int), ("semantics", C.c_int),
                ("window_begin", C.c_int64), ("window_end", C.c_int64), ("seed", C.c_uint64),
                ("states_resample", C.c_void_p), ("states_init", C.c_void_p)]


class ChromRun(C.Structure):
    _fields_ = [("name_off", C.c_int64), ("name_len", C.c_int32), ("reserved", C.c_int32), ("first_record", C.c_int64)]


class CssProbes(C.Structure):
    _fields_ = [("status", C.c_void_p), ("X", C.c_void_p), ("evals", C.c_void_p), ("hits", C.c_void_p),
                ("nperm", C.c_void_p), ("smacof_iters", C.c_void_p), ("smacof_sigma", C.c_void_p)]


# every symbol include/fpt_b200.h declares: (restype, argtypes)
_P = C.c_void_p
_I = C.c_int
_DROPIN_FET = [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, C.c_double, _P, _P]
_DROPIN_CSS = [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]
SYMBOLS = {
    "fpt_last_error": (C.c_char_p, []),
    "fpt_device_count": (_I, []),
    "fpt_set_device": (_I, [_I]),
    "fpt_set_seed": (None, [C.c_uint64]),
    "fpt_get_seed": (C.c_uint64, []),
    "fpt_set_perm_mode": (None, [_I]),
    "fpt_get_perm_mode": (_I, []),
    "fpt_set_perm_large_kernel": (None, [_I]),
    "fpt_set_perm_small_kernel": (None, [_I]),
    "fpt_set_mds_small_kernel": (None, [_I]),
    "fpt_set_lanczos_threads": (None, [_I]),
    "fpt_debug_umma_phases": (_I, [_P]),
    "fpt_debug_lanczos_phases": (_I, [_P]),
    "fpt_set_lanczos_form": (None, [_I]),
    "fpt_set_k4_mode": (None, [_I]),
    "fpt_debug_k4_phases": (_I, [_P]),
    "fpt_debug_k4_counts": (_I, [C.POINTER(Genotypes), C.POINTER(ScanRange), _I, C.c_int64, _P]),
    "fpt_css_perm_rechecks": (C.c_longlong, []),
    "fpt_release": (None, []),
    "fpt_window_state": (C.c_uint64, [C.c_uint64, C.c_int64, _I]),
    "fpt_profile_enable": (_I, [_I]),
    "fpt_profile_summary": (_I, [C.c_char_p, C.c_size_t]),
    "fpt_fet_threadcompute": (_I, _DROPIN_FET),
    "fpt_fet_compute": (_I, _DROPIN_FET),
    "fpt_css_threadcompute": (_I, _DROPIN_CSS),
    "fpt_css_compute": (_I, _DROPIN_CSS),
    "fpt_fet_scan": (_I, [C.POINTER(Genotypes), C.POINTER(ScanRange), C.c_double, _P, _P, _P]),
    "fpt_css_scan": (_I, [C.POINTER(Genotypes), C.POINTER(ScanRange), _I, _I, _I, _I, _P, _P, _P, C.POINTER(CssProbes)]),
    "fpt_fet_per_snp": (_I, [C.POINTER(Genotypes), _P, _P]),
    "fpt_fet_tables": (_I, [_P, C.c_int64, _I, _P]),
    "fpt_dev_fet_count_f64": (_I, [_P, _P, C.c_int64, _I, _I, _P, _P]),
    "fpt_dev_fet_count_i8": (_I, [_P, _P, C.c_int64, _I, _I, _P, _P]),
    "fpt_dev_fet_score": (_I, [_P, C.c_int64, _I, _I, _P, _P]),
    "fpt_dev_window_table": (_I, [_P, C.c_int64, C.POINTER(ScanRange), _P, _P, _P, _P]),
    "fpt_dev_fet_windows": (_I, [_P, _P, _P, C.POINTER(ScanRange), _I, C.c_double, _P, _P, _P, _P, _P]),
    "fpt_dev_css_planes_bytes": (C.c_size_t, [C.c_int64, _I]),
    "fpt_dev_css_pack_f64": (_I, [_P, _P, C.c_int64, _I, _I, _P, _P]),
    "fpt_dev_css_pack_i8": (_I, [_P, _P, C.c_int64, _I, _I, _P, _P]),
    "fpt_dev_css_absdiff": (_I, [_P, _P, C.c_int64, _P, _P]),
    "fpt_dev_css_workspace_bytes": (C.c_size_t, [_I, C.c_int64, _I]),
    "fpt_dev_css_windows": (_I, [_P, _P, _I, _I, _P, _P, C.POINTER(ScanRange), _I, _I, _I, _P, C.c_size_t, _P, _P, _P,
                                 C.POINTER(CssProbes), _P]),
    "fpt_vcf_scan": (_I, [C.c_char_p, C.c_size_t, _P, _P, _P]),
    "fpt_vcf_parse": (_I, [C.c_char_p, C.c_size_t, C.c_int64, _I, _I, _I, _P, _I, C.c_int64, _P, _P, _P, C.c_int64, _P]),
    "fpt_gtrack_scan": (_I, [C.c_char_p, C.c_size_t, _P]),
    "fpt_gtrack_parse": (_I, [C.c_char_p, C.c_size_t, _I, _I, _I, C.c_int64, _P, _P, _P, C.c_int64, _P]),
    "fpt_compact_codes": (_I, [_P, C.c_int64, _P]),
}

_lib = None


def load():
    """Load libfpt_b200.so (once). Raises FptError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FptError(FPT_ERR_NO_DEVICE, "%s is missing: build it with `make -C %s` (or __graft_entry__.build()); "
                       "there is no CPU fallback" % (LIB_PATH, os.path.join(_HERE, "csrc")))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != FPT_OK:
        raise FptError(rc, load().fpt_last_error().decode("utf-8", "replace"))
