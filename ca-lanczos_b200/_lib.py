"""ctypes binding of libcalz.so (the C ABI declared in include/calz.h).

There is no CPU fallback: if the library is missing it is built with nvcc (build.py); if that fails, or if
no CUDA device is present when a context is created, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcalz.so")

OK = 0
ERR_CHOL = 4
LAYOUT = {"auto": 0, "csr": 1, "sell": 2, "selld": 3}
LAYOUT_NAME = {v: k for k, v in LAYOUT.items()}
QR = {"tsqr": 0, "cholqr": 1, "cholqr2": 2}

c_i64 = C.c_int64
c_dp = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_ip = C.POINTER(C.c_int)
c_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/calz.h declares must be listed here (tests check it)
SIGNATURES = {
    "calz_version": (C.c_int, []),
    "calz_init": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "calz_finalize": (C.c_int, [c_vp]),
    "calz_last_error": (C.c_char_p, [c_vp]),
    "calz_set_stream": (C.c_int, [c_vp, c_vp]),
    "calz_get_stream": (c_vp, [c_vp]),
    "calz_sync": (C.c_int, [c_vp]),
    "calz_launch_count": (c_i64, [c_vp, C.c_int]),
    "calz_set_option": (C.c_int, [c_vp, C.c_char_p, c_i64]),
    "calz_shared_context": (C.c_int, [C.POINTER(c_vp)]),
    "calz_shared_release": (C.c_int, []),
    "calz_mat_cache_get_csc64": (C.c_int, [c_vp, c_i64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), c_dp, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "calz_mat_cache_clear": (C.c_int, [c_vp]),
    "calz_vec_create": (C.c_int, [c_vp, c_i64, C.c_int, C.POINTER(c_vp)]),
    "calz_vec_destroy": (C.c_int, [c_vp]),
    "calz_vec_info": (C.c_int, [c_vp, C.POINTER(c_vp), c_i64p, c_ip, c_i64p]),
    "calz_vec_upload": (C.c_int, [c_vp, C.c_int, C.c_int, c_dp, c_i64]),
    "calz_vec_download": (C.c_int, [c_vp, C.c_int, C.c_int, c_dp, c_i64]),
    "calz_vec_copy": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, C.c_int]),
    "calz_comm_unique_id": (C.c_int, [C.c_char_p, C.c_char_p]),
    "calz_comm_init": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_char_p, C.c_char_p]),
    "calz_comm_rank": (C.c_int, [c_vp, c_ip, c_ip]),
    "calz_partition_bounds": (C.c_int, [c_i64, C.c_int, c_i64p]),
    "calz_level_sets": (C.c_int, [c_i64, c_i64, c_i64, c_i64p, c_i32p, c_i64, c_i64, C.c_int, c_i32p]),
    "calz_mat_create_csr": (C.c_int, [c_vp, c_i64, c_i64, c_i64, c_i64p, c_i32p, c_dp, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "calz_mat_create_csc64": (C.c_int, [c_vp, c_i64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), c_dp, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "calz_mat_destroy": (C.c_int, [c_vp]),
    "calz_mat_info": (C.c_int, [c_vp, C.c_char_p, c_i64p]),
    "calz_mat_ghost_indices": (C.c_int, [c_vp, c_i64p, c_i64p]),
    "calz_mat_recv_list": (C.c_int, [c_vp, C.c_int, c_i64p, c_i64p]),
    "calz_mat_send_list": (C.c_int, [c_vp, C.c_int, c_i64p, c_i64p]),
    "calz_spmv": (C.c_int, [c_vp, c_vp, c_vp]),
    "calz_mpk_monomial": (C.c_int, [c_vp, c_vp, C.c_int, c_vp, c_i64]),
    "calz_mpk_newton": (C.c_int, [c_vp, c_vp, C.c_int, c_dp, c_dp, C.c_int, c_vp, c_i64]),
    "calz_mpk_inplace": (C.c_int, [c_vp, c_vp, C.c_int, c_dp, c_dp, C.c_int, C.c_int, C.POINTER(c_vp), c_i64p]),
    "calz_spmv_host": (C.c_int, [c_vp, c_dp, c_dp]),
    "calz_mpk_monomial_host": (C.c_int, [c_vp, c_dp, C.c_int, c_dp, c_i64]),
    "calz_mpk_newton_host": (C.c_int, [c_vp, c_dp, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_i64]),
    "calz_tsqr": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, c_i64, c_dp]),
    "calz_cholqr": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, c_i64, c_dp, c_ip]),
    "calz_normalize": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, C.c_int, C.c_double, c_vp, c_i64, c_dp, c_ip]),
    "calz_project": (C.c_int, [c_vp, c_i64, C.c_int, C.POINTER(c_vp), c_i64p, c_ip, C.c_int, c_vp, c_i64, C.c_int, C.POINTER(c_dp)]),
    "calz_project_and_normalize": (C.c_int, [c_vp, c_i64, C.c_int, C.POINTER(c_vp), c_i64p, c_ip, C.c_int, c_vp, c_i64,
                                             C.c_int, C.c_int, c_vp, c_i64, C.POINTER(c_dp), c_dp, c_ip, c_ip]),
    "calz_project_and_normalize_async": (C.c_int, [c_vp, c_i64, C.c_int, C.POINTER(c_vp), c_i64p, c_ip, C.c_int, c_vp, c_i64,
                                                   C.c_int, C.c_int, c_vp, c_i64, c_ip]),
    "calz_pan_collect": (C.c_int, [c_vp, C.c_int, C.POINTER(c_dp), c_dp, c_ip, c_ip, c_ip]),
    "calz_tsqr_host": (C.c_int, [c_vp, c_i64, C.c_int, c_dp, c_i64, c_dp, c_i64, c_dp]),
    "calz_cholqr_host": (C.c_int, [c_vp, c_i64, C.c_int, c_dp, c_i64, c_dp, c_i64, c_dp, c_ip]),
    "calz_normalize_host": (C.c_int, [c_vp, c_i64, C.c_int, c_dp, c_i64, C.c_int, C.c_double, c_dp, c_i64, c_dp, c_ip]),
    "calz_project_host": (C.c_int, [c_vp, c_i64, C.c_int, C.POINTER(c_dp), c_i64p, c_ip, C.c_int, c_dp, c_i64, C.c_int, C.POINTER(c_dp)]),
    "calz_project_and_normalize_host": (C.c_int, [c_vp, c_i64, C.c_int, C.POINTER(c_dp), c_i64p, c_ip, C.c_int, c_dp, c_i64,
                                                  C.c_int, C.c_int, c_dp, c_i64, C.POINTER(c_dp), c_dp, c_ip, c_ip]),
    "calz_dmma_peak": (C.c_int, [c_vp, c_dp]),
    "calz_block_axpy": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, C.c_int, c_dp, c_vp, c_i64, c_vp, c_i64]),
    "calz_orth_error": (C.c_int, [c_vp, c_i64, C.c_int, C.POINTER(c_vp), c_i64p, c_ip, C.c_int, C.c_int, c_dp]),
    "calz_gram": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp]),
}

_lib = None


class CalzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libcalz error %d: %s" % (code, msg))
        self.code = code


def load(build_if_missing: bool = True):
    """Load libcalz.so (building it in-tree first if it does not exist).  Never falls back to the CPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise FileNotFoundError(LIB_PATH + " is missing: run `python ca-lanczos_b200/build.py`")
        import importlib.util
        spec = importlib.util.spec_from_file_location("_calz_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, ctx=None):
    if status != OK:
        msg = load().calz_last_error(ctx)
        raise CalzError(status, msg.decode() if msg else "?")


def nccl_library_path():
    """Path of the libnccl.so.2 bundled with torch (what torch.distributed itself uses), or None."""
    try:
        import nvidia.nccl  # type: ignore
        for base in list(getattr(nvidia.nccl, "__path__", [])):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None
