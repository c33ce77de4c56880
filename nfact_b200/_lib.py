"""ctypes binding of libnfact_b200.so (the C ABI declared in include/nfact_b200.h).

There is no CPU path: if the shared library is missing or cannot be loaded this module raises, and
every compute entry point of the package raises with it.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnfact_b200.so")


class NfbError(RuntimeError):
    """A libnfact_b200 call failed; the message is nfb_last_error()."""


class NmfParams(ctypes.Structure):
    _fields_ = [
        ("solver", c_int),
        ("max_iter", c_int),
        ("tol", c_double),
        ("l1_reg_W", c_double),
        ("l1_reg_H", c_double),
        ("l2_reg_W", c_double),
        ("l2_reg_H", c_double),
        ("verbose", c_int),
    ]


# name -> (restype, argtypes); every name here must be declared in include/nfact_b200.h
SIGNATURES = {
    "nfb_last_error": (c_char_p, []),
    "nfb_version": (c_char_p, []),
    "nfb_create": (c_int, [c_int, c_void_p, POINTER(c_void_p)]),
    "nfb_destroy": (c_int, [c_void_p]),
    "nfb_synchronize": (c_int, [c_void_p]),
    "nfb_release_cached": (c_int, [c_void_p]),
    "nfb_set_comm": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "nfb_nccl_unique_id": (c_int, [c_char_p, c_void_p]),
    "nfb_nccl_init_rank": (c_int, [c_void_p, c_char_p, c_void_p, c_int, c_int]),
    "nfb_ingest_coo": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                               POINTER(c_int64), POINTER(c_double), c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "nfb_ingest_coo_matrix": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                      POINTER(c_int64), POINTER(c_double), c_int64, c_int64, c_int,
                                      POINTER(c_void_p)]),
    "nfb_ingest_dot_matrix": (c_int, [c_void_p, c_void_p, c_int64, c_double, POINTER(c_void_p), POINTER(c_int64),
                                      POINTER(c_int)]),
    "nfb_dot_parse_coo_gpu": (c_int, [c_void_p, c_void_p, c_int64, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p),
                                      POINTER(c_int64), POINTER(c_int64), POINTER(c_int)]),
    "nfb_dot_count_lines": (c_int, [c_void_p, c_int64, c_int, POINTER(c_int64)]),
    "nfb_dot_parse": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "nfb_dot_parse_coo": (c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p,
                                  POINTER(c_int64)]),
    "nfb_lz4_frame_size": (c_int, [c_void_p, c_int64, c_int, POINTER(c_int64)]),
    "nfb_lz4_frame_decompress": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, POINTER(c_int64)]),
    "nfb_lz4_last_chunks": (c_int, [POINTER(c_int64), POINTER(c_int64)]),
    "nfb_matrix_from_host": (c_int, [c_void_p, c_void_p, c_int64, c_int64, POINTER(c_void_p)]),
    "nfb_matrix_from_device": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, POINTER(c_void_p)]),
    "nfb_matrix_free": (c_int, [c_void_p]),
    "nfb_matrix_shape": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int64)]),
    "nfb_matrix_sumsq": (c_int, [c_void_p, POINTER(c_double)]),
    "nfb_matrix_sum": (c_int, [c_void_p, POINTER(c_double)]),
    "nfb_matrix_has_negative": (c_int, [c_void_p, POINTER(c_int)]),
    "nfb_matrix_matmul": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64]),
    "nfb_nmf_fit_host": (c_int, [c_void_p, c_void_p, c_int, POINTER(NmfParams), c_void_p, c_void_p,
                                 POINTER(c_int), POINTER(c_double)]),
    "nfb_nmf_create": (c_int, [c_void_p, c_void_p, c_int, POINTER(NmfParams), POINTER(c_void_p)]),
    "nfb_nmf_free": (c_int, [c_void_p]),
    "nfb_nmf_set_factors_host": (c_int, [c_void_p, c_void_p, c_void_p]),
    "nfb_nmf_set_factors_dev": (c_int, [c_void_p, c_void_p, c_void_p]),
    "nfb_nmf_get_factors_host": (c_int, [c_void_p, c_void_p, c_void_p]),
    "nfb_nmf_get_factors_dev": (c_int, [c_void_p, c_void_p, c_void_p]),
    "nfb_nmf_iterate": (c_int, [c_void_p, c_int, POINTER(c_double)]),
    "nfb_nmf_run": (c_int, [c_void_p, POINTER(c_int), POINTER(c_double)]),
    "nfb_nmf_error": (c_int, [c_void_p, POINTER(c_double)]),
    "nfb_nmf_exchange_mode": (c_int, [c_void_p, POINTER(c_int)]),
    "nfb_launch_count": (c_int64, [c_void_p]),
    "nfb_profile_gemms": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_int64)]),
    "nfb_dual_regression_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                         POINTER(c_int), POINTER(c_int)]),
    "nfb_dual_regression_post_host": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_double, c_void_p, c_int,
                                              c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_int),
                                              POINTER(c_int)]),
    "nfb_dr_components_create": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, POINTER(c_void_p), POINTER(c_int)]),
    "nfb_dr_components_free": (c_int, [c_void_p]),
    "nfb_dual_regression_resident_host": (c_int, [c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_int, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, POINTER(c_int)]),
    "nfb_component_postops_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_void_p, c_int, c_int,
                                           c_double, c_void_p]),
    "nfb_component_wta_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_double, c_void_p]),
    "nfb_component_loadings_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_int64, c_void_p]),
    "nfb_row_correlation_host": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "nfb_sso_create": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, POINTER(c_void_p)]),
    "nfb_sso_free": (c_int, [c_void_p]),
    "nfb_nmf_init_random": (c_int, [c_void_p, c_double, ctypes.c_uint64]),
    "nfb_sso_add_run": (c_int, [c_void_p, c_void_p, c_int, c_double]),
    "nfb_sso_similarity_host": (c_int, [c_void_p, c_void_p, c_void_p]),
    "nfb_sso_seed_final": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "nfb_nmf_factors_all_zero": (c_int, [c_void_p, c_void_p]),
    "nfb_host_alloc": (c_int, [c_int64, POINTER(c_void_p)]),
    "nfb_host_free": (c_int, [c_void_p]),
    "nfb_test_gemm": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64,
                              c_int, c_int]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C nfact_b200/csrc`).  nfact_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def pinned_empty(shape, dtype):
    """``np.empty(shape, dtype)`` on page-locked memory from the library's pool (nfb_host_alloc): device-to-host
    copies into it run at PCIe speed.  The block goes back to the pool when the array (and every view of it)
    is garbage-collected."""
    import weakref

    import numpy as np
    lib = load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    ptr = c_void_p()
    check(lib.nfb_host_alloc(nbytes, ctypes.byref(ptr)))
    raw = (ctypes.c_byte * max(nbytes, 1)).from_address(ptr.value)
    weakref.finalize(raw, lib.nfb_host_free, c_void_p(ptr.value))
    return np.frombuffer(raw, dtype=dtype, count=nbytes // dtype.itemsize).reshape(shape)


def pinned_wrap(ptr: int, shape, dtype):
    """numpy view of a page-locked block handed out by the library (nfb_host_alloc inside a C call); the block goes
    back to the pool when the array is garbage-collected."""
    import weakref

    import numpy as np
    lib = load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    raw = (ctypes.c_byte * max(nbytes, 1)).from_address(ptr)
    weakref.finalize(raw, lib.nfb_host_free, c_void_p(ptr))
    return np.frombuffer(raw, dtype=dtype, count=nbytes // dtype.itemsize).reshape(shape)


def check(rc: int) -> None:
    if rc != 0:
        msg = load().nfb_last_error()
        raise NfbError(msg.decode() if msg else f"libnfact_b200 call failed with code {rc}")
