"""Matrix ingest -- host-side mirror of the reference's loaders, arithmetic on the GPU.

Same names, arguments and error behaviour as
  NFACT/base/matrix_handling.py            read_in_matrix :93, load_single_matrix :116, load_fdt_matrix :143,
                                           normalise_components :45, thresholding :218, save_matrix :161
  NFACT/decomp/decomposition/matrix_handling.py   process_fdt_matrix2 :12, avg_fdt :73, load_previous_matrix :47
LZ4 decoding and the text parse stay on the host (they are I/O, native and multi-threaded in libnfact_b200); duplicate summation, waytotal scaling, group averaging,
densification and the float32 rounding run in libnfact_b200 (ingest.cu) and are bit-identical to the reference.
"""
from __future__ import annotations

import os
from collections import namedtuple

import numpy as np

from . import device
from .utils import colours, error_and_exit, nprint

CooMatrix = namedtuple("CooMatrix", "rows cols data shape scale")


def read_in_matrix(matfile: str) -> np.ndarray:
    """NFACT/base/matrix_handling.py:93-113.  float64 [T+1, 3], from ``.dot`` text or an ``.lz4`` frame."""
    return np.ascontiguousarray(parse_dot_bytes(_read_bytes(matfile), matfile).T)


def _read_bytes(matfile: str):
    """The file's text as a byte buffer; ``.lz4`` frames are decoded natively (``decompress_lz4_frame``)."""
    with open(matfile, "rb") as fil:
        raw = fil.read()
    return decompress_lz4_frame(raw, matfile) if matfile.endswith(".lz4") else raw


def decompress_lz4_frame(raw, name: str = "<buffer>") -> np.ndarray:
    """``lz4.frame.open(matfile, "rb").read()`` (NFACT/base/matrix_handling.py:108-110) without the lz4 module:
    libnfact_b200's multi-threaded LZ4 frame decoder (``nfb_lz4_frame_decompress``, checksums verified).
    Returns the decoded bytes as a uint8 array; a corrupt file raises ``RuntimeError`` like python-lz4 does."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    src = np.frombuffer(raw, dtype=np.uint8)
    size, written = ctypes.c_int64(), ctypes.c_int64()
    try:
        _lib.check(lib.nfb_lz4_frame_size(src.ctypes.data, src.size, 0, ctypes.byref(size)))
        out = np.empty(size.value, dtype=np.uint8)
        _lib.check(lib.nfb_lz4_frame_decompress(src.ctypes.data, src.size, out.ctypes.data, out.size, 0,
                                                ctypes.byref(written)))
    except _lib.NfbError as e:
        raise RuntimeError(f"{name}: {e}") from None
    return out[:written.value]


def _as_buffer(raw):
    """(address, length, keep-alive) of a bytes-like object or uint8 array"""
    arr = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw)
    return arr.ctypes.data, arr.size, arr


def parse_dot_bytes(raw: bytes, name: str = "<buffer>") -> np.ndarray:
    """Native multi-threaded parse (libnfact_b200 ``nfb_dot_parse``): float64 [3, lines] = the row, column
    and value columns of the file (``np.loadtxt(...).T``)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    buf, nbytes, _keep = _as_buffer(raw)
    n_lines = ctypes.c_int64()
    _lib.check(lib.nfb_dot_count_lines(buf, nbytes, 0, ctypes.byref(n_lines)))
    if n_lines.value == 0:
        raise ValueError(f"{name}: file is empty")
    cols3 = np.empty((3, n_lines.value), dtype=np.float64)
    try:
        _lib.check(lib.nfb_dot_parse(buf, nbytes, 0, n_lines.value, cols3[0].ctypes.data, cols3[1].ctypes.data,
                                     cols3[2].ctypes.data))
    except _lib.NfbError as e:
        raise ValueError(f"{name}: {e}") from None
    return cols3


def parse_dot_coo(raw: bytes, name: str = "<buffer>", pinned: bool = False):
    """One native pass from the file's bytes to the ingest kernels' input (``nfb_dot_parse_coo``): 0-based int32
    rows / cols and float64 values of all lines but the last, and the (nrows, ncols) the last line states.
    ``pinned`` puts the three arrays in page-locked memory so their upload runs at PCIe speed."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    buf, nbytes, _keep = _as_buffer(raw)
    n_lines = ctypes.c_int64()
    _lib.check(lib.nfb_dot_count_lines(buf, nbytes, 0, ctypes.byref(n_lines)))
    if n_lines.value == 0:
        raise ValueError(f"{name}: file is empty")
    nnz = n_lines.value - 1
    alloc = _lib.pinned_empty if pinned else (lambda shape, dtype: np.empty(shape, dtype=dtype))
    rows, cols, vals = alloc((nnz,), np.int32), alloc((nnz,), np.int32), alloc((nnz,), np.float64)
    shape = (ctypes.c_int64 * 2)()
    try:
        _lib.check(lib.nfb_dot_parse_coo(buf, nbytes, 0, n_lines.value, rows.ctypes.data, cols.ctypes.data,
                                         vals.ctypes.data, shape))
    except _lib.NfbError as e:
        raise ValueError(f"{e}") from None
    return rows, cols, vals, (int(shape[0]), int(shape[1]))


def load_single_matrix(matfile: str, pinned: bool = False) -> CooMatrix:
    """NFACT/base/matrix_handling.py:116-140.  Returns the COO triplets (0-based) plus the
    ``1e8 / waytotal`` scale; the sparse-matrix arithmetic itself happens on the device."""
    raw = _read_bytes(matfile)
    parsed = None
    if pinned and len(raw) and os.environ.get("NFB_DOT_ON_DEVICE", "1") != "0":
        # a CUDA context is up (pinned staging was asked for): parse on the device, host parser as the fallback
        try:
            parsed = device.parse_dot_coo_device(raw)
        except Exception as e:
            raise ValueError(f"{e}") from None
    rows, cols, data, shape = parsed if parsed is not None else parse_dot_coo(raw, matfile, pinned)
    waytotal = float(np.loadtxt(os.path.join(os.path.dirname(matfile), "waytotal")))
    return CooMatrix(rows, cols, data, shape, 1e8 / waytotal)


def _cuda_ready() -> bool:
    """Page-locked staging needs a CUDA context; the parser alone also runs on a box without a GPU."""
    try:
        device.get_handle()
        return True
    except Exception:
        return False


def load_fdt_matrix(matfile: str, resident: bool = False, handle=None, raw=None):
    """NFACT/base/matrix_handling.py:143-158: dense float32 [m, n].

    ``resident=True`` keeps the matrix in HBM and returns a ``DeviceMatrix`` (what ``nmf_decomp`` /
    ``nmf_dual_regression`` take directly) instead of copying 4mn bytes to the host and back; the text itself is
    then parsed on the device too (``nfb_ingest_dot_matrix``), the host parser only takes over for files with
    numbers outside the device parser's exact fast path or with duplicate coordinates."""
    if resident and os.environ.get("NFB_DOT_ON_DEVICE", "1") != "0":
        if raw is None:           # ``raw``: the file's (decompressed) bytes when a reader thread already has them
            raw = _read_bytes(matfile)
        if len(raw) == 0:
            raise ValueError(f"{matfile}: file is empty")
        waytotal = float(np.loadtxt(os.path.join(os.path.dirname(matfile), "waytotal")))
        try:
            xdev = device.ingest_dot_resident(raw, 1e8 / waytotal, handle)
        except Exception as e:
            raise ValueError(f"{e}") from None
        if xdev is not None:
            return xdev
        rows, cols, data, shape = parse_dot_coo(raw, matfile, pinned=True)
        coo = CooMatrix(rows, cols, data, shape, 1e8 / waytotal)
    else:
        coo = load_single_matrix(matfile, pinned=_cuda_ready())
    ingest = device.ingest_coo_resident if resident else device.ingest_coo
    return ingest([(coo.rows, coo.cols, coo.data, coo.scale)], coo.shape[0], coo.shape[1], False, handle)


def avg_fdt(list_of_matfiles: list) -> np.ndarray:
    """NFACT/decomp/decomposition/matrix_handling.py:73-100: group average, dense float32."""
    if len(list_of_matfiles) == 0:
        raise ValueError("no fdt_matrix2 files given")
    subjects, shape = [], None
    for matrix in list_of_matfiles:
        coo = load_single_matrix(matrix)
        if shape is None:
            shape = coo.shape
        elif coo.shape != shape:
            raise ValueError(f"inconsistent shapes {shape} and {coo.shape}")
        subjects.append((coo.rows, coo.cols, coo.data, coo.scale))
    return device.ingest_coo(subjects, shape[0], shape[1], True)


def process_fdt_matrix2(list_of_ptx_folds: list, group_mode: bool) -> np.ndarray:
    """NFACT/decomp/decomposition/matrix_handling.py:12-44."""
    list_of_fdt = [
        os.path.join(sub_folder, fname)
        for sub_folder in list_of_ptx_folds
        for fname in ("fdt_matrix2.dot", "fdt_matrix2.dot.lz4")
        if os.path.exists(os.path.join(sub_folder, fname))
    ]
    fdt_matrix2 = None
    if group_mode:
        try:
            fdt_matrix2 = avg_fdt(list_of_fdt)
        except Exception as e:
            error_and_exit(False, f"Unable to load fdt_matrix2 due to {e}")
    if not group_mode:
        try:
            fdt_matrix2 = load_fdt_matrix(list_of_fdt[0])
        except Exception as e:
            error_and_exit(False, f"Unable to load fdt_matrix2 due to {e}")
    return fdt_matrix2


def load_previous_matrix(path: str):
    """NFACT/decomp/decomposition/matrix_handling.py:47-70."""
    try:
        return np.load(os.path.join(path))
    except Exception:
        col = colours()
        nprint(f"{col['pink']}Error:{col['reset']} Unable to read previous matrix. Averaging")
        return None


def save_matrix(matrix: np.ndarray, directory: str, matrix_name: str) -> None:
    """NFACT/base/matrix_handling.py:161-181."""
    try:
        np.save(os.path.join(directory, matrix_name), matrix)
    except Exception as e:
        error_and_exit(False, f"Unable to save matrix due to {e}")


def _component_postops(data: np.ndarray, comp_major: bool, group_id=None, n_groups: int = 0, threshold: bool = False,
                       zscore: float = 0.0, normalise: bool = False):
    """``nfb_component_postops_host``: z-threshold ``data`` in place and / or return its z-scored copy.
    ``data`` must be a C-contiguous float32 / float64 2-D array ([k, len] when ``comp_major`` else [len, k])."""
    from ctypes import c_void_p
    from . import _lib
    handle = device.get_handle()
    k, length = data.shape if comp_major else data.shape[::-1]
    gid = None if group_id is None else np.ascontiguousarray(group_id, dtype=np.int32)
    if gid is not None and gid.shape != (length,):
        raise ValueError(f"group ids have shape {gid.shape}, expected {(length,)}")
    norm = np.empty_like(data) if normalise else None
    _lib.check(handle.lib.nfb_component_postops_host(
        handle.ptr, c_void_p(data.ctypes.data), int(data.dtype == np.float64), int(comp_major), k, length,
        None if gid is None else c_void_p(gid.ctypes.data), int(n_groups), int(threshold), float(zscore),
        None if norm is None else c_void_p(norm.ctypes.data)))
    return norm


def _device_ready(arr: np.ndarray) -> bool:
    """The kernels take float32 / float64 C-contiguous 2-D arrays; other inputs are converted by the callers."""
    return arr.ndim == 2 and arr.dtype in (np.float32, np.float64) and arr.flags.c_contiguous and arr.flags.writeable


def thresholding(component: np.ndarray, zscore_val) -> np.ndarray:
    """NFACT/base/matrix_handling.py:218-238: entries below ``mean + zscore_val * std`` of their component (row)
    are zeroed IN PLACE, on the device (float64 statistics)."""
    if component.size == 0:
        return component
    if _device_ready(component):
        _component_postops(component, True, threshold=True, zscore=zscore_val)
        return component
    work = np.ascontiguousarray(component, dtype=np.float64 if component.dtype != np.float32 else np.float32)
    _component_postops(work, True, threshold=True, zscore=zscore_val)
    component[...] = work           # non-contiguous view (e.g. per_seed = comps[rows].T in the reference) or other dtype
    return component


def normalise_components(grey_matter: np.ndarray, white_matter: np.ndarray) -> dict:
    """NFACT/base/matrix_handling.py:45-68 (StandardScaler semantics: population std, constant component ->
    scale 1), on the device; returns new arrays of the inputs' dtype."""
    def zscore(arr, comp_major):
        arr = np.ascontiguousarray(arr, dtype=np.float32 if arr.dtype == np.float32 else np.float64)
        if arr.size == 0:
            return arr.copy()
        return _component_postops(arr, comp_major, normalise=True)
    col = colours()
    nprint(f"{col['pink']}Normalising:{col['reset']} Components")
    return {"grey_matter": zscore(grey_matter, False), "white_matter": zscore(white_matter, True)}


def threshold_grey_components(components: np.ndarray, coord_path: str, seeds: list, zscore_val) -> np.ndarray:
    """NFACT/base/matrix_handling.py:184-215: z-threshold the grey components separately for every seed
    (the seed index of each row is the second-to-last column of ``coords_for_fdt_matrix2``); one device pass
    with per-(component, seed) statistics, in place."""
    seeds_id = np.loadtxt(coord_path, dtype=int)[:, -2]
    if components.size == 0 or len(seeds) == 0:
        return components
    if _device_ready(components):
        _component_postops(components, False, seeds_id, len(seeds), threshold=True, zscore=zscore_val)
        return components
    work = np.ascontiguousarray(components, dtype=np.float64 if components.dtype != np.float32 else np.float32)
    _component_postops(work, False, seeds_id, len(seeds), threshold=True, zscore=zscore_val)
    components[...] = work
    return components


def component_loadings(subject_data: np.ndarray, group_data: np.ndarray) -> np.ndarray:
    """``__correlating`` of NFACT/stats/stats_component_loadings.py:199-231: per-component correlation between a
    subject's dual-regression maps and the group maps, both ``[samples, k]`` like the reference; float64 ``[k]``.
    Computed on the device (``nfb_component_loadings_host``, float64 statistics)."""
    from ctypes import c_void_p
    from . import _lib
    subject = np.asarray(subject_data)
    group = np.asarray(group_data)
    if subject.shape != group.shape or subject.ndim != 2:
        raise ValueError(f"operands could not be broadcast together with shapes {subject.shape} {group.shape}")
    dtype = np.float32 if subject.dtype == np.float32 and group.dtype == np.float32 else np.float64
    subject = np.ascontiguousarray(subject, dtype=dtype)
    group = np.ascontiguousarray(group, dtype=dtype)
    length, k = subject.shape
    out = np.empty(k, dtype=np.float64)
    if k == 0 or length == 0:
        return np.full(k, np.nan)
    handle = device.get_handle()
    _lib.check(handle.lib.nfb_component_loadings_host(handle.ptr, c_void_p(subject.ctypes.data),
                                                      c_void_p(group.ctypes.data), int(dtype == np.float64), 0, k,
                                                      length, c_void_p(out.ctypes.data)))
    return out


def create_wta_map(component: np.ndarray, axis: int, z_thr: float) -> np.ndarray:
    """NFACT/decomp/pipes/image_handling.py:216-245: winner-takes-all map.  The columns of ``component`` are
    z-scored (StandardScaler), then every entry along ``axis`` is replaced by the 1-based index of its first maximum,
    0 where that maximum is below ``z_thr``: ``create_wta_map(white [k, n], 0, z)`` -> int [1, n],
    ``create_wta_map(grey [m, k], 1, z)`` -> int [m, 1].  On the device (``nfb_component_wta_host``, float64
    statistics, z-scores rounded to the array's dtype like sklearn's in-place transform)."""
    from ctypes import c_void_p
    from . import _lib
    arr = np.asarray(component)
    if arr.ndim != 2 or axis not in (0, 1):
        raise ValueError(f"expected a 2-D array and axis 0 or 1, got shape {arr.shape}, axis {axis}")
    arr = np.ascontiguousarray(arr, dtype=np.float32 if arr.dtype == np.float32 else np.float64)
    rows, cols = arr.shape
    if rows == 0:
        raise ValueError(f"Found array with 0 sample(s) (shape={arr.shape}) while a minimum of 1 is required by "
                         "StandardScaler.")
    if cols == 0:
        raise ValueError(f"Found array with 0 feature(s) (shape={arr.shape}) while a minimum of 1 is required by "
                         "StandardScaler.")
    k, length = (rows, cols) if axis == 0 else (cols, rows)
    out = np.empty(length, dtype=np.int64)
    handle = device.get_handle()
    _lib.check(handle.lib.nfb_component_wta_host(handle.ptr, c_void_p(arr.ctypes.data), int(arr.dtype == np.float64),
                                                 int(axis == 0), k, length, float(z_thr), c_void_p(out.ctypes.data)))
    return np.array(out.reshape((1, cols) if axis == 0 else (rows, 1)), dtype=int)


def thresholding_components(threshold_val: int, path_to_coords: str, seeds: list, components: dict) -> dict:
    """NFACT/base/matrix_handling.py:241-277."""
    col = colours()
    shown = threshold_val if threshold_val != 0 else "Not thresholding"
    nprint(f"{col['pink']}Thresholding value:{col['reset']} {shown}\n")
    if threshold_val != 0:
        components["white_components"] = thresholding(components["white_components"], threshold_val)
        components["grey_components"] = threshold_grey_components(components["grey_components"], path_to_coords,
                                                                  seeds, threshold_val)
    return components
