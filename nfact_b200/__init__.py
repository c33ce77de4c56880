"""nfact_b200 -- B200-native (sm_100a) decomposition hot path of NFACT.

Drop-in replacements for the three seams where NFACT hands NumPy arrays to sklearn / scipy:
ingest (``load_fdt_matrix`` / ``avg_fdt`` / ``process_fdt_matrix2``), the NMF solve (``nmf_decomp``)
and the non-negative dual regression (``nmf_dual_regression``).  All arithmetic runs in
``libnfact_b200.so`` (hand-written CUDA, C ABI in ``include/nfact_b200.h``); there is no CPU path.
"""
from ._lib import NfbError, load as load_library  # noqa: F401
from .device import DeviceComponents, DeviceMatrix, NmfState, get_handle  # noqa: F401
from .dual_regression import (dual_regression_subject, nmf_dual_regression, run_decomp, run_locally,  # noqa: F401
                              stream_subjects)
from .matrix_handling import (avg_fdt, component_loadings, create_wta_map, decompress_lz4_frame, load_fdt_matrix,  # noqa: F401
                              load_previous_matrix, load_single_matrix, normalise_components, process_fdt_matrix2, read_in_matrix, save_matrix,
                              threshold_grey_components, thresholding, thresholding_components)
from .nmf import NMF, compute_regularization, get_parameters, nmf_decomp  # noqa: F401
from .nmfsso import NMFsso, nmf_sso  # noqa: F401

__version__ = "0.1.0"
