"""Non-negative dual regression -- host-side mirror of the reference's DR seam.

Same names, arguments and error behaviour as
  NFACT/dual_reg/dual_regression/dual_regression_methods.py   run_decomp :8-42, nmf_dual_regression :92-116
Every scipy.optimize.nnls call of the reference (one per target column, then one per seed row) is
replaced by two tensor-core passes over X (Gram right-hand sides) plus the batched float64 NNLS kernel.
"""
from __future__ import annotations

from ctypes import byref, c_int

import numpy as np

from ._lib import check, pinned_empty
from .device import DeviceComponents, DeviceMatrix
from .utils import error_and_exit


def nmf_dual_regression(components: dict, connectivity_matrix, n_jobs: int = 1, *, threshold: float = 0,
                        seeds_id=None, n_seeds: int = 0, normalise: bool = False) -> dict:
    """NFACT/dual_reg/dual_regression/dual_regression_methods.py:92-116.

    ``n_jobs`` is accepted for signature compatibility; the device solves all columns at once.
    Returns float64 arrays {"grey_components": [m,k], "white_components": [k,n]} like the reference.

    Keyword extras (not in the reference's signature; ``dual_regression_subject`` uses them): the per-subject
    post-processing the pipeline runs next (run_dual_regression.py:166-186) applied on the device before the
    results are copied out -- ``threshold`` != 0: ``thresholding_components`` with ``seeds_id`` (the seed index of
    every row) and ``n_seeds``; ``normalise``: adds "normalised_grey" / "normalised_white"."""
    grey = components["grey_components"]
    prepared = grey if isinstance(grey, DeviceComponents) else None     # uploaded once by the subject loop
    if prepared is None:
        grey = np.asarray(grey)
        if grey.dtype != np.float32:          # scipy's nnls casts everything else to float64
            grey = grey.astype(np.float64, copy=False)
        grey = np.ascontiguousarray(grey)
        if grey.ndim != 2:
            raise ValueError(f"Expected a 2D array, but the shape of A is {grey.shape}")
    if isinstance(connectivity_matrix, DeviceMatrix):
        xdev, owns = connectivity_matrix, False
    else:
        xdev, owns = DeviceMatrix.from_host(np.asarray(connectivity_matrix)), True
    out = {}
    try:
        if xdev.has_nonfinite:
            raise ValueError("array must not contain infs or NaNs")
        m, n = xdev.shape
        if grey.shape[0] != m:
            raise ValueError("Incompatible dimensions. The first dimension of "
                             f"A is {grey.shape[0]}, while the shape of b is {(m,)}")
        k = grey.shape[1]
        gs = pinned_empty((m, k), np.float64)
        hs = pinned_empty((k, n), np.float64)
        gs_norm = pinned_empty((m, k), np.float64) if normalise else None
        hs_norm = pinned_empty((k, n), np.float64) if normalise else None
        seed_arr = None
        if threshold:
            seed_arr = np.ascontiguousarray(np.zeros(m) if seeds_id is None else seeds_id, dtype=np.int32)
            if seed_arr.shape != (m,):
                raise ValueError(f"seeds_id has shape {seed_arr.shape}, expected {(m,)}")
            n_seeds = int(n_seeds) if seeds_id is not None else 1
        bad, nonfinite = c_int(), c_int()
        h = xdev.handle
        post = (float(threshold), seed_arr.ctypes.data if seed_arr is not None else None, int(n_seeds),
                gs.ctypes.data, hs.ctypes.data, gs_norm.ctypes.data if normalise else None,
                hs_norm.ctypes.data if normalise else None)
        if prepared is not None:
            nonfinite.value = int(prepared.has_nonfinite)
            check(h.lib.nfb_dual_regression_resident_host(h.ptr, xdev.ptr, prepared.ptr, *post, byref(bad)))
        else:
            check(h.lib.nfb_dual_regression_post_host(h.ptr, xdev.ptr, k, grey.ctypes.data,
                                                      int(grey.dtype == np.float32), *post, byref(bad),
                                                      byref(nonfinite)))
        if nonfinite.value:               # checked on the device while G is transposed
            raise ValueError("array must not contain infs or NaNs")
        if bad.value:
            raise RuntimeError("Maximum number of iterations reached.")
        if normalise:
            out = {"normalised_grey": gs_norm, "normalised_white": hs_norm}
    finally:
        if owns:
            xdev.free()
    return {"grey_components": gs, "white_components": hs, **out}


def run_decomp(decomp, components: dict, connectivity_matrix, parallel: int = None) -> dict:
    """NFACT/dual_reg/dual_regression/dual_regression_methods.py:8-42."""
    try:
        components = decomp(components, connectivity_matrix, parallel)
    except ValueError as e:
        error_and_exit(False, f"Components have incompatable size with connectivity Matrix {e}")
    except Exception as e:
        error_and_exit(False, f"Unable to perform dual regression due to {e}")
    return components


def _subject_matrix_file(fdt_path: str):
    import os
    return next((os.path.join(fdt_path, fname) for fname in ("fdt_matrix2.dot.lz4", "fdt_matrix2.dot")
                 if os.path.exists(os.path.join(fdt_path, fname))), None)


def _postprocess_options(fdt_path: str, seeds: list, threshold: int, normalise: bool) -> dict:
    """Keyword arguments that make ``nmf_dual_regression`` run ``thresholding_components`` (seed index of every
    row = second-to-last column of ``coords_for_fdt_matrix2``, NFACT/base/matrix_handling.py:205) and
    ``normalise_components`` on the device."""
    import os
    from .utils import colours, nprint
    col = colours()
    shown = int(threshold) if int(threshold) != 0 else "Not thresholding"
    nprint(f"{col['pink']}Thresholding value:{col['reset']} {shown}\n")
    options = {"normalise": bool(normalise)}
    if int(threshold) != 0:
        seeds_id = np.loadtxt(os.path.join(fdt_path, "coords_for_fdt_matrix2"), dtype=int)[:, -2]
        options.update(threshold=int(threshold), seeds_id=seeds_id, n_seeds=len(seeds))
    if normalise:
        nprint(f"{col['pink']}Normalising:{col['reset']} Components")
    return options


def dual_regression_subject(fdt_path: str, components: dict, seeds: list, threshold: int = 3,
                            normalise: bool = False, parallel: int = 1, matrix=None) -> dict:
    """Compute part of NFACT/dual_reg/dual_regression/run_dual_regression.py:66-199 for one subject:
    load the subject's fdt_matrix2 -> dual regression -> z-threshold -> optional z-score normalisation.
    Saving the images stays with the reference's ``save_dual_regression_images``.  ``matrix``: the subject's
    matrix when it has already been loaded (a ``DeviceMatrix`` from the prefetching loop of ``run_locally``)."""
    import os
    from .matrix_handling import load_fdt_matrix
    owns = matrix is None
    if owns:
        try:
            matrix = load_fdt_matrix(_subject_matrix_file(fdt_path), resident=True)   # stays in HBM
        except Exception:
            error_and_exit(False, f"Unable to load {os.path.basename(os.path.normpath(fdt_path))} fdt matrix")
    options = _postprocess_options(fdt_path, seeds, threshold, normalise)
    try:
        return run_decomp(lambda comp, mat, par: nmf_dual_regression(comp, mat, par, **options), components, matrix,
                          parallel)
    finally:
        if owns:
            matrix.free()


def stream_subjects(load_matrix, n_subjects: int, components: dict, device: int | None = None, prefetch: bool = True,
                    options=None):
    """Generator over ``(index, dual-regression result)`` for subjects ``0 .. n_subjects-1``.

    ``load_matrix(index, handle) -> DeviceMatrix`` makes subject ``index`` resident using ``handle``.  With
    ``prefetch`` the next subject is loaded (file parse, triplet upload, scatter-add ingest, plane split) by a
    background thread on a second handle -- its own CUDA stream -- while the current subject's regression
    runs, so a subject costs max(load, solve) instead of their sum; two subjects are resident at a time.
    ``options(index) -> dict``: per-subject keyword arguments for ``nmf_dual_regression`` (post-processing)."""
    options = options or (lambda idx: {})
    from concurrent.futures import ThreadPoolExecutor
    from .device import Handle, get_handle
    first = get_handle(device)
    handles = [first, Handle(first.device)] if prefetch and n_subjects > 1 else [first]
    # the group maps are the same for every subject: uploaded once per handle, not once per subject
    grey = components["grey_components"]
    prepared = [grey if isinstance(grey, DeviceComponents) else DeviceComponents(grey, hd) for hd in handles] \
        if n_subjects > 0 else []
    per_handle = [dict(components, grey_components=prep) for prep in prepared]
    try:
        if len(handles) == 1:
            for idx in range(n_subjects):
                xdev = load_matrix(idx, first)
                try:
                    yield idx, nmf_dual_regression(per_handle[0], xdev, **options(idx))
                finally:
                    xdev.free()
            return
        with ThreadPoolExecutor(max_workers=1) as pool:
            pending = pool.submit(load_matrix, 0, handles[0])
            for idx in range(n_subjects):
                xdev = pending.result()
                if idx + 1 < n_subjects:
                    pending = pool.submit(load_matrix, idx + 1, handles[(idx + 1) % 2])
                try:
                    result = nmf_dual_regression(per_handle[idx % 2], xdev, **options(idx))   # on xdev's stream
                finally:
                    xdev.free()
                yield idx, result
                result = None     # let the page-locked result blocks go back to the pool before the next subject
    finally:
        for prep in prepared:
            if prep is not grey:
                prep.free()
        for extra in handles[1:]:
            extra.close()
        # the ingest keeps every subject's planes / staging buffers for the next subject: hand them back now
        check(first.lib.nfb_release_cached(first.ptr))


def run_locally(list_of_subject_dirs: list, components: dict, seeds: list, threshold: int = 3,
                normalise: bool = False, rank: int = 0, world: int = 1, save=None, prefetch: bool = True,
                read_ahead: int = 0) -> dict:
    """Subject loop of NFACT/dual_reg/dual_regression/set_up_dual_regression.py:186-235.  Subjects are
    independent: with ``world`` processes (one per GPU) rank r takes subjects r, r+world, ... and no
    collective is needed.  ``save(subject_dir, dr_results)`` receives each result (the reference's image
    writer); without ``save`` the results of this rank are returned keyed by subject directory (with it the
    dictionary only records which subjects were done).  The next subject's
    matrix is parsed and ingested while the current one is being regressed (``stream_subjects``).

    ``read_ahead`` > 0: that many host threads read -- and, for ``.dot.lz4``, decompress -- the files of the next
    subjects while the current ones are ingested and regressed.  An LZ4 frame with linked blocks (what python-lz4
    writes) decodes on one core, ~4 s for a 5 GB matrix, far longer than the 0.2 s the GPU needs per subject;
    different subjects' files are independent, so N readers give N times the decode throughput.  Every pending file
    holds its decompressed text in host memory (~5 GB at C3).  Off by default."""
    import os
    from .matrix_handling import _read_bytes, load_fdt_matrix
    mine = [d for idx, d in enumerate(list_of_subject_dirs) if idx % world == rank]
    readers, pending = None, {}
    if read_ahead > 0 and os.environ.get("NFB_DOT_ON_DEVICE", "1") != "0":
        from concurrent.futures import ThreadPoolExecutor
        readers = ThreadPoolExecutor(max_workers=int(read_ahead))

    def schedule_reads(upto):
        # files [0, upto) have been asked for: keep `read_ahead` of the following ones in flight
        for nxt in range(upto, min(upto + int(read_ahead), len(mine))):
            if nxt not in pending:
                pending[nxt] = readers.submit(_read_bytes, _subject_matrix_file(mine[nxt]))

    def load(idx, handle):
        try:
            raw = None
            if readers is not None:
                schedule_reads(idx)
                raw = pending.pop(idx).result()
                schedule_reads(idx + 1)
            return load_fdt_matrix(_subject_matrix_file(mine[idx]), resident=True, handle=handle, raw=raw)
        except Exception:
            error_and_exit(False, f"Unable to load {os.path.basename(os.path.normpath(mine[idx]))} fdt matrix")

    out = {}
    try:
        options = lambda idx: _postprocess_options(mine[idx], seeds, threshold, normalise)   # noqa: E731
        for idx, dr_results in stream_subjects(load, len(mine), components, prefetch=prefetch, options=options):
            if save is not None:
                save(mine[idx], dr_results)
            # results are only kept when nobody consumes them on the way (100 C3-sized subjects are 27 GB)
            out[mine[idx]] = dr_results if save is None else None
            del dr_results
    except ValueError as e:
        error_and_exit(False, f"Components have incompatable size with connectivity Matrix {e}")
    except SystemExit:
        raise
    except Exception as e:
        error_and_exit(False, f"Unable to perform dual regression due to {e}")
    finally:
        if readers is not None:
            readers.shutdown(wait=False, cancel_futures=True)
    return out
