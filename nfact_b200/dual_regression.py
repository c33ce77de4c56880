"""Non-negative dual regression -- host-side mirror of the reference's DR seam.

Same names, arguments and error behaviour as
  NFACT/dual_reg/dual_regression/dual_regression_methods.py   run_decomp :8-42, nmf_dual_regression :92-116
Every scipy.optimize.nnls call of the reference (one per target column, then one per seed row) is
replaced by two tensor-core passes over X (Gram right-hand sides) plus the batched float64 NNLS kernel.
"""
from __future__ import annotations

from ctypes import byref, c_int

import numpy as np

from ._lib import check
from .device import DeviceMatrix
from .utils import error_and_exit


def nmf_dual_regression(components: dict, connectivity_matrix, n_jobs: int = 1) -> dict:
    """NFACT/dual_reg/dual_regression/dual_regression_methods.py:92-116.

    ``n_jobs`` is accepted for signature compatibility; the device solves all columns at once.
    Returns float64 arrays {"grey_components": [m,k], "white_components": [k,n]} like the reference.
    """
    grey = np.ascontiguousarray(components["grey_components"], dtype=np.float64)
    if grey.ndim != 2:
        raise ValueError(f"Expected a 2D array, but the shape of A is {grey.shape}")
    if not np.isfinite(grey).all():
        raise ValueError("array must not contain infs or NaNs")
    if isinstance(connectivity_matrix, DeviceMatrix):
        xdev, owns = connectivity_matrix, False
    else:
        xhost = np.asarray(connectivity_matrix)
        if not np.isfinite(xhost).all():
            raise ValueError("array must not contain infs or NaNs")
        xdev, owns = DeviceMatrix.from_host(xhost), True
    try:
        m, n = xdev.shape
        if grey.shape[0] != m:
            raise ValueError("Incompatible dimensions. The first dimension of "
                             f"A is {grey.shape[0]}, while the shape of b is {(m,)}")
        k = grey.shape[1]
        gs = np.empty((m, k), dtype=np.float64)
        hs = np.empty((k, n), dtype=np.float64)
        bad = c_int()
        h = xdev.handle
        check(h.lib.nfb_dual_regression_host(h.ptr, xdev.ptr, k, grey.ctypes.data, gs.ctypes.data, hs.ctypes.data,
                                             byref(bad)))
        if bad.value:
            raise RuntimeError("Maximum number of iterations reached.")
    finally:
        if owns:
            xdev.free()
    return {"grey_components": gs, "white_components": hs}


def run_decomp(decomp, components: dict, connectivity_matrix, parallel: int = None) -> dict:
    """NFACT/dual_reg/dual_regression/dual_regression_methods.py:8-42."""
    try:
        components = decomp(components, connectivity_matrix, parallel)
    except ValueError as e:
        error_and_exit(False, f"Components have incompatable size with connectivity Matrix {e}")
    except Exception as e:
        error_and_exit(False, f"Unable to perform dual regression due to {e}")
    return components
