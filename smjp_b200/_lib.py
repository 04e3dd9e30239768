"""ctypes binding of libsmjp_b200.so (include/smjp_b200.h).  No CPU fallback: importing this module
works anywhere (so that host-side logic is testable), but every compute entry point needs the
built CUDA library and a B200; failures are loud."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsmjp_b200.so")
CSRC = os.path.join(_HERE, "csrc")

SMJP_F64, SMJP_F32 = 0, 1
ST_GRID_NOT_INCREASING, ST_DEGENERATE, ST_OVERFLOW, ST_BAD_INPUT = 1, 2, 4, 8
CAND_REFERENCE, CAND_EXACT = 0, 1
RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC = 0, 1

_lib = None

c_i64p = ctypes.POINTER(ctypes.c_int64)
c_i32p = ctypes.POINTER(ctypes.c_int32)
c_f64p = ctypes.POINTER(ctypes.c_double)
P = ctypes.c_void_p  # every pointer argument is passed as an address (host or device)

# name -> (restype, argtypes); mirrors include/smjp_b200.h one to one
SIGNATURES = {
    "smjp_abi_version": (ctypes.c_int, []),
    "smjp_last_error": (ctypes.c_char_p, []),
    "smjp_ctx_create": (P, [ctypes.c_int, P]),
    "smjp_ctx_destroy": (None, [P]),
    "smjp_ctx_sync": (ctypes.c_int, [P]),
    "smjp_ctx_launch_count": (ctypes.c_int64, [P]),
    "smjp_ctx_get_stat": (ctypes.c_int64, [P, ctypes.c_char_p]),
    "smjp_ctx_set_option": (ctypes.c_int, [P, ctypes.c_char_p, ctypes.c_int64]),
    "smjp_ctx_kernel_time": (ctypes.c_int, [P, ctypes.c_char_p, c_f64p, c_i64p]),
    "smjp_set_model": (ctypes.c_int, [P, ctypes.c_int, ctypes.c_int, P, P, ctypes.c_double, P,
                                      ctypes.c_double, ctypes.c_double]),
    "smjp_ffbs_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                     ctypes.c_int, P, P, P, P, P, P, P, P, P, P, ctypes.c_int64, ctypes.c_int64]),
    "smjp_ffbs_host": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                      ctypes.c_int, P, P, P, P, P, P, P, P, P, P]),
    "smjp_prior_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, ctypes.c_int, ctypes.c_uint64, P, P, P, P]),
    "smjp_simulate_obs_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, ctypes.c_uint64, P]),
    "smjp_candidates_count_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                                 P, P, c_i64p, c_i64p]),
    "smjp_candidates_fill_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                                ctypes.c_int, P, P, P]),
    "smjp_candidates_host": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                            ctypes.c_int, P, P, ctypes.c_int64, c_i64p]),
    "smjp_sweep_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                      ctypes.c_int, ctypes.c_int, ctypes.c_int64, P, P, P, P, P, P, P, P, P,
                                      c_i64p, c_i64p]),
    "smjp_sweep_host": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int64, P, P, P, P, P, P, P, P,
                                       c_i64p]),
    "smjp_sweep_host_begin": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int64, P, P, P, P, P, P, P, P,
                                             c_i64p]),
    "smjp_sweep_host_end": (ctypes.c_int, [P]),
    "smjp_hmm_dense_host": (ctypes.c_int, [P, ctypes.c_int, ctypes.c_int, P, P, ctypes.c_int, P, P, ctypes.c_int, P,
                                           P, P, ctypes.c_uint64, ctypes.c_int, P, P, P, P, P]),
    "smjp_ffbs_dense_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                           ctypes.c_int, P, P, P, P, P, P, P, ctypes.c_int64]),
    "smjp_sweep_dense_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, P, P, P, ctypes.c_uint64, ctypes.c_uint32,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int64, P, P, P, P, P, P, P, P, P,
                                            c_i64p, c_i64p]),
    "smjp_shape_mle_dev": (ctypes.c_int, [P, ctypes.c_int, ctypes.c_int, P, P, P, P, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                          P, P, P, P, P]),
    "smjp_pf_dev": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_int64, P, P,
                                   P, P, P, P, P]),
    "smjp_hmm_dense_dev": (ctypes.c_int, [P, ctypes.c_int, ctypes.c_int, P, P, ctypes.c_int, P, P, ctypes.c_int, P, P, P,
                                          ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, P, P, P, P, P]),
    "smjp_pf_host": (ctypes.c_int, [P, ctypes.c_int, P, P, P, P, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, P, P,
                                    P, P, P, P, P]),
}


class EngineError(RuntimeError):
    pass


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode:
        raise EngineError("building libsmjp_b200.so failed")
    return LIB_PATH


def load():
    """dlopen the library and attach the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EngineError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().smjp_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        if rc == -3:
            raise MemoryError(msg)
        raise EngineError("smjp_b200 error %d: %s" % (rc, msg))
