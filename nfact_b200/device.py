"""Thin object layer over the C ABI: device handle, resident connectivity matrix, NMF state.

PyTorch appears only as plumbing (device tensors handed over as raw pointers, the current CUDA
stream, torch.distributed for rendezvous); all arithmetic happens inside libnfact_b200.so.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int, c_int64, c_void_p

import numpy as np

from . import _lib
from ._lib import NmfParams, check

_handles: dict = {}


class Handle:
    """One per (process, device): CUDA stream + scratch owned by the library."""

    def __init__(self, device: int = 0, stream: int | None = None):
        """``stream``: a cudaStream_t as an integer (e.g. ``torch.cuda.Stream().cuda_stream``) so that
        the caller's CUDA events see the library's kernels; None -> the library creates its own."""
        self.lib = _lib.load()
        self.device = int(device)
        self._ptr = c_void_p()
        check(self.lib.nfb_create(self.device, c_void_p(stream) if stream else None, byref(self._ptr)))
        self.rank, self.world, self.group = 0, 1, None

    @property
    def ptr(self) -> c_void_p:
        return self._ptr

    def synchronize(self) -> None:
        check(self.lib.nfb_synchronize(self._ptr))

    def launch_count(self) -> int:
        return int(self.lib.nfb_launch_count(self._ptr))

    def release_cached(self) -> None:
        """Hand back the device memory the library keeps cached between calls (the subject loop's per-subject
        buffers, the stream-ordered pool)."""
        check(self.lib.nfb_release_cached(self._ptr))

    def profile_gemms(self, enable: bool):
        """(total ms, launches) of the X-streaming GEMMs since the last call; then set the enable flag."""
        total, count = c_double(), c_int64()
        check(self.lib.nfb_profile_gemms(self._ptr, int(enable), byref(total), byref(count)))
        return total.value, count.value

    def init_nccl(self, rank: int, world: int, broadcast_bytes) -> None:
        """Create a NCCL communicator inside the library.  ``broadcast_bytes(b: bytes|None) -> bytes``
        must return rank 0's 128-byte unique id on every rank (e.g. via torch.distributed)."""
        path = find_libnccl().encode()
        uid = ctypes.create_string_buffer(128)
        if rank == 0:
            check(self.lib.nfb_nccl_unique_id(path, uid))
        payload = broadcast_bytes(bytes(uid.raw) if rank == 0 else None)
        uid = ctypes.create_string_buffer(payload, 128)
        check(self.lib.nfb_nccl_init_rank(self._ptr, path, uid, rank, world))
        self.rank, self.world = rank, world

    def close(self) -> None:
        if self._ptr:
            self.lib.nfb_destroy(self._ptr)
            self._ptr = c_void_p()


def find_libnccl() -> str:
    """The NCCL that ships with torch (no system dependency)."""
    import glob
    import os
    import sys
    for base in sys.path:
        hits = glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
        if hits:
            return sorted(hits)[0]
    return "libnccl.so.2"


_default_device = 0


def set_default_handle(handle: Handle) -> None:
    """Make ``handle`` (its device, stream and - in row-sharded runs - NCCL communicator) the one the
    reference-facing entry points (``nmf_decomp``, ``nmf_dual_regression``, the loaders) use in this process."""
    global _default_device
    _handles[handle.device] = handle
    _default_device = handle.device


def get_handle(device: int | None = None) -> Handle:
    device = _default_device if device is None else device
    if device not in _handles:
        _handles[device] = Handle(device)
    return _handles[device]


class DeviceMatrix:
    """Connectivity matrix X resident in HBM as fp16 hi/lo planes (4 B/element)."""

    def __init__(self, handle: Handle, ptr: c_void_p, shape):
        self.handle, self._ptr, self.shape = handle, ptr, tuple(shape)

    @classmethod
    def from_host(cls, x: np.ndarray, handle: Handle | None = None) -> "DeviceMatrix":
        handle = handle or get_handle()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise ValueError("expected a 2-D matrix")
        ptr = c_void_p()
        check(handle.lib.nfb_matrix_from_host(handle.ptr, x.ctypes.data, x.shape[0], x.shape[1], byref(ptr)))
        return cls(handle, ptr, x.shape)

    @classmethod
    def from_device_ptr(cls, data_ptr: int, m: int, n: int, ld: int | None = None,
                        handle: Handle | None = None) -> "DeviceMatrix":
        handle = handle or get_handle()
        ptr = c_void_p()
        check(handle.lib.nfb_matrix_from_device(handle.ptr, c_void_p(data_ptr), m, n, ld or n, byref(ptr)))
        return cls(handle, ptr, (m, n))

    @property
    def ptr(self) -> c_void_p:
        return self._ptr

    @property
    def sumsq(self) -> float:
        out = c_double()
        check(self.handle.lib.nfb_matrix_sumsq(self._ptr, byref(out)))
        return out.value

    @property
    def mean(self) -> float:
        out = c_double()
        check(self.handle.lib.nfb_matrix_sum(self._ptr, byref(out)))
        return out.value / (self.shape[0] * self.shape[1])

    @property
    def has_negative(self) -> bool:
        out = c_int()
        check(self.handle.lib.nfb_matrix_has_negative(self._ptr, byref(out)))
        return bool(out.value)

    @property
    def has_nonfinite(self) -> bool:
        """inf / NaN anywhere in X (the float64 sum of squares accumulated during the split pass overflows
        or turns NaN exactly then) -- sklearn's / scipy's finiteness check without a host pass over X."""
        return not np.isfinite(self.sumsq)

    def matmul_dev(self, trans: bool, b_ptr: int, n_b: int, out_ptr: int, ld_out: int) -> None:
        check(self.handle.lib.nfb_matrix_matmul(self.handle.ptr, self._ptr, int(trans), c_void_p(b_ptr), n_b,
                                                c_void_p(out_ptr), ld_out))

    def free(self) -> None:
        if self._ptr:
            self.handle.lib.nfb_matrix_free(self._ptr)
            self._ptr = c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceComponents:
    """The group grey components ``[m, k]`` prepared once for many dual regressions on one handle
    (``nfb_dr_components_create``): the nfact_dr subject loop regresses every subject onto the same maps."""

    def __init__(self, grey: np.ndarray, handle: Handle | None = None):
        grey = np.asarray(grey)
        if grey.dtype != np.float32:          # scipy's nnls casts everything else to float64
            grey = grey.astype(np.float64, copy=False)
        grey = np.ascontiguousarray(grey)
        if grey.ndim != 2:
            raise ValueError(f"Expected a 2D array, but the shape of A is {grey.shape}")
        self.handle = handle or get_handle()
        self.shape = grey.shape
        self._ptr = c_void_p()
        nonfinite = c_int()
        check(self.handle.lib.nfb_dr_components_create(self.handle.ptr, c_void_p(grey.ctypes.data),
                                                       int(grey.dtype == np.float32), grey.shape[0], grey.shape[1],
                                                       byref(self._ptr), byref(nonfinite)))
        self.has_nonfinite = bool(nonfinite.value)

    @property
    def ptr(self) -> c_void_p:
        if not self._ptr:
            raise ValueError("the components have been freed")
        return self._ptr

    def free(self) -> None:
        if self._ptr:
            self.handle.lib.nfb_dr_components_free(self._ptr)
            self._ptr = c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def make_params(solver: str, max_iter: int, tol: float, regs, verbose: int = 0) -> NmfParams:
    if solver not in ("cd", "mu"):
        raise ValueError(f"Invalid solver parameter '{solver}'.")
    l1w, l1h, l2w, l2h = regs
    return NmfParams(0 if solver == "cd" else 1, int(max_iter), float(tol), float(l1w), float(l1h), float(l2w),
                     float(l2h), int(verbose))


class NmfState:
    """Stepwise NMF solver on a resident matrix (benchmarks and per-iteration parity tests)."""

    def __init__(self, x: DeviceMatrix, k: int, params: NmfParams):
        self.x, self.k, self.params = x, int(k), params
        self.handle = x.handle
        self._ptr = c_void_p()
        check(self.handle.lib.nfb_nmf_create(self.handle.ptr, x.ptr, self.k, byref(params), byref(self._ptr)))

    def set_factors(self, w: np.ndarray, h: np.ndarray) -> None:
        m, n = self.x.shape
        w = np.ascontiguousarray(w, dtype=np.float32)
        h = np.ascontiguousarray(h, dtype=np.float32)
        if w.shape != (m, self.k) or h.shape != (self.k, n):
            raise ValueError(f"Array with wrong shape passed to NMF: W {w.shape}, H {h.shape}, "
                             f"expected {(m, self.k)} and {(self.k, n)}")
        check(self.handle.lib.nfb_nmf_set_factors_host(self._ptr, w.ctypes.data, h.ctypes.data))

    def set_factors_dev(self, w_ptr: int, h_ptr: int) -> None:
        check(self.handle.lib.nfb_nmf_set_factors_dev(self._ptr, c_void_p(w_ptr), c_void_p(h_ptr)))

    def get_factors(self):
        m, n = self.x.shape
        w = _lib.pinned_empty((m, self.k), np.float32)
        h = _lib.pinned_empty((self.k, n), np.float32)
        check(self.handle.lib.nfb_nmf_get_factors_host(self._ptr, w.ctypes.data, h.ctypes.data))
        return w, h

    def iterate(self, n_iters: int = 1, want_violation: bool = False):
        viol = c_double()
        check(self.handle.lib.nfb_nmf_iterate(self._ptr, int(n_iters), byref(viol) if want_violation else None))
        return viol.value if want_violation else None

    def run(self):
        n_iter, err = c_int(), c_double()
        check(self.handle.lib.nfb_nmf_run(self._ptr, byref(n_iter), byref(err)))
        return n_iter.value, err.value

    def init_random(self, avg: float, seed: int) -> None:
        """|avg N(0, 1)| into both factors on the device (the distribution of sklearn's init='random')."""
        check(self.handle.lib.nfb_nmf_init_random(self._ptr, float(avg), int(seed) & 0xFFFFFFFFFFFFFFFF))

    def factors_all_zero(self):
        """(W is all zeros, H is all zeros) for the factors currently in the state."""
        flags = (c_int * 2)()
        check(self.handle.lib.nfb_nmf_factors_all_zero(self._ptr, flags))
        return bool(flags[0]), bool(flags[1])

    @property
    def exchange_mode(self) -> str:
        """'single', 'nccl' or 'p2p': how the H half-step of a row-sharded solve exchanges data."""
        out = c_int()
        check(self.handle.lib.nfb_nmf_exchange_mode(self._ptr, byref(out)))
        return ("single", "nccl", "p2p")[out.value]

    def error(self) -> float:
        err = c_double()
        check(self.handle.lib.nfb_nmf_error(self._ptr, byref(err)))
        return err.value

    def free(self) -> None:
        if self._ptr:
            self.handle.lib.nfb_nmf_free(self._ptr)
            self._ptr = c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SsoStacks:
    """The components of all NMF-sso runs, resident in HBM (nfb_sso_*): every finished run's factors are copied
    into the stacks on the device (white copy z-thresholded there), the similarity matrix of the stack is the only
    thing that goes to the host for clustering, and the centrotype rows are gathered on the device into the state
    of the final solve."""

    def __init__(self, x: DeviceMatrix, k: int, runs: int):
        self.handle, self.k, self.runs = x.handle, int(k), int(runs)
        self.m, self.n = x.shape
        self._ptr = c_void_p()
        check(self.handle.lib.nfb_sso_create(self.handle.ptr, self.m, self.n, self.k, self.runs, byref(self._ptr)))

    def add_run(self, state: "NmfState", run: int, zscore: float) -> None:
        check(self.handle.lib.nfb_sso_add_run(self._ptr, state._ptr, int(run), float(zscore)))

    def similarity(self):
        """(Pearson correlations [R, R] float32 of the stacked white components, dead-row flags [R])."""
        rows = self.runs * self.k
        sim = np.empty((rows, rows), dtype=np.float32)
        dead = np.zeros(rows, dtype=np.int32)
        check(self.handle.lib.nfb_sso_similarity_host(self._ptr, sim.ctypes.data, dead.ctypes.data))
        return sim, dead

    def seed_final(self, state: "NmfState", centroids) -> None:
        idx = np.ascontiguousarray(centroids, dtype=np.int32)
        check(self.handle.lib.nfb_sso_seed_final(self._ptr, state._ptr, idx.ctypes.data, int(idx.size)))

    def free(self) -> None:
        if self._ptr:
            self.handle.lib.nfb_sso_free(self._ptr)
            self._ptr = c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _coo_args(subjects):
    n = len(subjects)
    rows = [np.ascontiguousarray(s[0], dtype=np.int32) for s in subjects]
    cols = [np.ascontiguousarray(s[1], dtype=np.int32) for s in subjects]
    vals = [np.ascontiguousarray(s[2], dtype=np.float64) for s in subjects]
    arr_t = c_void_p * n
    keep = (rows, cols, vals)
    return (n, arr_t(*[r.ctypes.data for r in rows]), arr_t(*[c.ctypes.data for c in cols]),
            arr_t(*[v.ctypes.data for v in vals]), (c_int64 * n)(*[len(r) for r in rows]),
            (c_double * n)(*[float(s[3]) for s in subjects]), keep)


def ingest_coo(subjects, nrows: int, ncols: int, group_mode: bool, handle: Handle | None = None) -> np.ndarray:
    """subjects: list of (rows int32 0-based, cols int32 0-based, vals float64, scale float)."""
    handle = handle or get_handle()
    n, rows_p, cols_p, vals_p, nnz, scale, _keep = _coo_args(subjects)
    # page-locked when it is cheap to be (pinning tens of GB costs seconds and starves the host)
    big = 4 * nrows * ncols > (2 << 30)
    out = np.empty((nrows, ncols), dtype=np.float32) if big else _lib.pinned_empty((nrows, ncols), np.float32)
    check(handle.lib.nfb_ingest_coo(handle.ptr, n, rows_p, cols_p, vals_p, nnz, scale, nrows, ncols,
                                    1 if group_mode else 0, out.ctypes.data, None))
    return out


def ingest_coo_resident(subjects, nrows: int, ncols: int, group_mode: bool,
                        handle: Handle | None = None) -> DeviceMatrix:
    """Same ingest, but the matrix stays in HBM as a ``DeviceMatrix`` (no dense round trip through the host)."""
    handle = handle or get_handle()
    n, rows_p, cols_p, vals_p, nnz, scale, _keep = _coo_args(subjects)
    ptr = c_void_p()
    check(handle.lib.nfb_ingest_coo_matrix(handle.ptr, n, rows_p, cols_p, vals_p, nnz, scale, nrows, ncols,
                                           1 if group_mode else 0, byref(ptr)))
    return DeviceMatrix(handle, ptr, (nrows, ncols))


def ingest_dot_resident(raw, scale: float, handle: Handle | None = None):
    """The bytes of one subject's ``fdt_matrix2.dot`` -> resident ``DeviceMatrix``, parsed on the device
    (``nfb_ingest_dot_matrix``).  Returns ``None`` when the file needs the host parser (a number outside the exact
    fast path, or duplicate coordinates that the reference sums in float64)."""
    handle = handle or get_handle()
    buf = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw)
    ptr, fallback = c_void_p(), c_int()
    shape = (c_int64 * 2)()
    check(handle.lib.nfb_ingest_dot_matrix(handle.ptr, c_void_p(buf.ctypes.data), buf.size, float(scale), byref(ptr),
                                           shape, byref(fallback)))
    if fallback.value:
        return None
    return DeviceMatrix(handle, ptr, (int(shape[0]), int(shape[1])))


def parse_dot_coo_device(raw, handle: Handle | None = None):
    """``parse_dot_coo`` on the device (``nfb_dot_parse_coo_gpu``): ``(rows0, cols0, vals, (nrows, ncols))`` as
    page-locked arrays, or ``None`` when the file needs the host parser."""
    handle = handle or get_handle()
    buf = np.frombuffer(raw, dtype=np.uint8) if not isinstance(raw, np.ndarray) else np.ascontiguousarray(raw)
    rows, cols, vals = c_void_p(), c_void_p(), c_void_p()
    nnz, fallback = c_int64(), c_int()
    shape = (c_int64 * 2)()
    check(handle.lib.nfb_dot_parse_coo_gpu(handle.ptr, c_void_p(buf.ctypes.data), buf.size, byref(rows), byref(cols),
                                           byref(vals), byref(nnz), shape, byref(fallback)))
    if fallback.value:
        return None
    n = nnz.value
    return (_lib.pinned_wrap(rows.value, (n,), np.int32), _lib.pinned_wrap(cols.value, (n,), np.int32),
            _lib.pinned_wrap(vals.value, (n,), np.float64), (int(shape[0]), int(shape[1])))

