"""ctypes binding of libnmfk_b200.so (the C ABI declared in include/nmfk_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present
when a context is created, this raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(os.path.dirname(_HERE))  # nmfk.jl_b200/
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libnmfk_b200.so")

F32, F64 = 0, 1


class NMFkError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("nmfk_b200 status %d: %s" % (status, msg))
        self.status = status


class NegativeEntriesError(NMFkError, ValueError):
    """ErrorException("All matrix entries must be nonnegative!") NMFkMultiplicative.jl:4-7"""


class Params(C.Structure):
    _fields_ = [("tol", C.c_double), ("tolOF", C.c_double), ("eps_clamp", C.c_double), ("weight", C.c_double),
                ("maxiter", C.c_int32), ("maxbaditers", C.c_int32), ("maxreattempts", C.c_int32),
                ("stopconv", C.c_int32), ("check_every", C.c_int32), ("Wfixed", C.c_int32), ("Hfixed", C.c_int32),
                ("normalize", C.c_int32), ("iter_limit", C.c_int32), ("engine", C.c_int32),
                ("clusterWmatrix", C.c_int32), ("stop_rule", C.c_int32), ("variant", C.c_int32),
                ("reserved", C.c_int32 * 1)]


class XInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("nnan", C.c_int64), ("nzero", C.c_int64),
                ("zero_row", C.c_int32), ("zero_col", C.c_int32), ("xmin", C.c_double), ("dtype", C.c_int32),
                ("reserved", C.c_int32)]


_P = C.c_void_p
_i32, _i64, _u64, _dbl = C.c_int32, C.c_int64, C.c_uint64, C.c_double
_pi32, _pi64, _pdbl = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

# name -> (restype, argtypes): one entry per declaration in include/nmfk_b200.h
SIGNATURES = {
    "nmfk_abi_version": (_i32, []),
    "nmfk_last_error": (C.c_char_p, [_P]),
    "nmfk_default_params": (None, [C.POINTER(Params)]),
    "nmfk_ctx_create": (_i32, [_i32, C.POINTER(_P)]),
    "nmfk_ctx_destroy": (_i32, [_P]),
    "nmfk_ctx_sync": (_i32, [_P]),
    "nmfk_set_X": (_i32, [_P, _P, _i64, _i64, _i32, _dbl, _P, _i32]),
    "nmfk_get_xinfo": (_i32, [_P, C.POINTER(XInfo)]),
    "nmfk_set_weight": (_i32, [_P, _P, _i64, _i64]),
    "nmfk_set_sparsity_options": (_i32, [_P, _dbl, _dbl, _dbl]),
    "nmfk_batch_create": (_i32, [_P, _i32, _i32, C.POINTER(_P)]),
    "nmfk_batch_destroy": (_i32, [_P]),
    "nmfk_batch_set_init": (_i32, [_P, _P, _P]),
    "nmfk_batch_init_random": (_i32, [_P, _u64]),
    "nmfk_batch_set_init_partial": (_i32, [_P, _P, _P, _u64]),
    "nmfk_batch_select": (_i32, [_P, _dbl, _dbl, _i32, _pi32, _pi32]),
    "nmfk_batch_create_hstack": (_i32, [_P, _i32, _i32, C.POINTER(_P)]),
    "nmfk_batch_device_ptrs": (_i32, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "nmfk_batch_import": (_i32, [_P, _P, _P, _pdbl, _pi32, _i32]),
    "nmfk_fit": (_i32, [_P, _i32, _P, _P, _pdbl]),
    "nmfk_comm_unique_id": (_i32, [_P]),
    "nmfk_ctx_comm_init": (_i32, [_P, _i32, _i32, _P, _i64, _i64]),
    "nmfk_ctx_comm_destroy": (_i32, [_P]),
    "nmfk_solve": (_i32, [_P, C.POINTER(_P), _i32, C.POINTER(Params)]),
    "nmfk_batch_get": (_i32, [_P, _P, _P, _pdbl, _pdbl, _pi32, _pi32]),
    "nmfk_batch_objective": (_i32, [_P, _dbl, _pdbl]),
    "nmfk_batch_cluster": (_i32, [_P, _i32, _pi32, _pi32, _pdbl, _pdbl, _pdbl, _P, _pi32]),
    "nmfk_batch_cluster_means": (_i32, [_P, _pi32, _pi32, _P, _P, _P, _P]),
    "nmfk_run_batch": (_i32, [_P, _i32, _i32, _P, _P, C.POINTER(Params), _P, _P, _pdbl, _pdbl, _pi32, _pi32]),
    "nmfk_trace": (_i32, [_P, _i32, _P, _P, C.POINTER(Params), _i32, _P, _P, _pdbl]),
    "nmfk_execute_run": (_i32, [_P, _i32, _i32, _P, _P, _u64, C.POINTER(Params), _P, _P, _pdbl, _pdbl, _pdbl, _pi64]),
    "nmfk_execute": (_i32, [_P, _pi32, _i32, _i32, C.POINTER(_P), C.POINTER(_P), _u64, C.POINTER(Params), _dbl,
                            C.POINTER(_P), C.POINTER(_P), _pdbl, _pdbl, _pdbl, _pi32, _pi64]),
    "nmfk_ctx_sweep_comm_init": (_i32, [_P, _i32, _i32, _P]),
    "nmfk_sweep": (_i32, [_P, _pi32, _i32, _i32, C.POINTER(_P), C.POINTER(_P), _u64, C.POINTER(Params), _dbl,
                          C.POINTER(_P), C.POINTER(_P), _pdbl, _pdbl, _pdbl, _pi32, _pi64, _pi64]),
    "nmfk_robustkmeans": (_i32, [_P, _pdbl, _i32, _i32, _i32, _i32, _pi32, _i32, _dbl, _i32, _pi32, _pdbl, _pdbl, _pi32, _pdbl,
                                 _pi32, _pi32, _pdbl, _pi32, _pi32]),
    "nmfk_getk": (_i32, [_pi32, _pdbl, _i32, _dbl, _i32]),
    "nmfk_signalorder": (_i32, [_P, _P, _i64, _i32, _i64, _i32, _pi32]),
    "nmfk_launch_count": (_i64, [_P]),
    "nmfk_last_solve_ms": (_dbl, [_P]),
    "nmfk_profile_enable": (_i32, [_P, _i32]),
    "nmfk_profile_get": (_i32, [_P, _pdbl, _pi64]),
    "nmfk_measure_peak": (_i32, [_P, _i32, _pdbl]),
    "nmfk_gemm_nt": (_i32, [_P, _i32, _P, _P, _i32, _i32, _i32, _P, _i32, _pdbl]),
    "nmfk_umma_timing": (_i32, [_P, _P, _P, _i32, _P, _P]),
    "nmfk_umma_selftest": (_i32, [_P, _P, _P, _i32, _P, _P, _P, _P, _pi32]),
    "nmfk_philox_host": (_i32, [_u64, _i64, _pdbl]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("nmfk_b200: %s is missing - build it with `python nmfk.jl_b200/build.py` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, ctx=None):
    if status == 0:
        return
    msg = load().nmfk_last_error(ctx)
    msg = msg.decode() if msg else ""
    if status == -2:
        raise NegativeEntriesError(status, msg)
    raise NMFkError(status, msg)
