"""NMF-sso orchestration on a resident matrix.

Mirrors NFACT/decomp/decomposition/sso/nmfsso.py (``NMFsso`` :62-238, ``nmf_sso`` :322-360) and the
clustering glue of sso_functions.py (``compute_similairty_matrix`` :49, ``sim2dis`` :71,
``clustering_components`` :359, ``idx2centrotype`` :426): N random-init solves -> per-run z-threshold of H
-> similarity of all N*k white-matter components -> average-linkage clustering cut at the elbow -> one
centrotype per cluster -> final ``init='custom'`` solve.

Difference by design: the reference copies X into POSIX shared memory and forks ``n_cores`` joblib workers;
here X is uploaded to HBM once (``DeviceMatrix``) and the N+1 solves reuse it, so ``n_cores`` only keeps its
meaning as "run the sso loop" and the single-run path does not inherit the reference's ``range(colours)``
TypeError (nmfsso.py:207).  ``nmf_sso`` on one GPU keeps the runs' components on the device as well
(``nmf_sso_resident``: device random starts, resident stacks, similarity from the stack, centrotypes gathered
into the final solve; ``NFB_SSO_RESIDENT=0`` or the ``NMFsso`` class give the host-stack flow).  Plots / csv
(sso_plotting.py) are out of scope.
"""
from __future__ import annotations

import numpy as np

from .device import DeviceMatrix
from .matrix_handling import thresholding
from .nmf import nmf_decomp
from .utils import colours, error_and_exit, nprint


def rownorm(nmf_mat: np.ndarray) -> np.ndarray:
    norms = np.linalg.norm(nmf_mat, axis=1, keepdims=True)
    norms[norms == 0] = 1
    return nmf_mat / norms


def compute_similairty_matrix(components: np.ndarray, handle=None) -> np.ndarray:
    """sso_functions.py:49-68: correlation between the row-normalised components, clipped to [0, 1].

    The (N k)^2 n product runs on the device (``nfb_row_correlation_host``: rows standardised with float64
    statistics, then one split-precision GEMM pass); Pearson correlation is invariant to the row
    normalisation the reference applies first, so ``rownorm`` is not needed.  float64 out like np.corrcoef."""
    from ctypes import c_void_p
    from . import _lib
    from .device import get_handle
    handle = handle or get_handle()
    comps = np.ascontiguousarray(components, dtype=np.float32)
    if comps.ndim != 2:
        raise ValueError("expected a 2-D stack of components")
    r, n = comps.shape
    sim32 = np.empty((r, r), dtype=np.float32)
    dead = np.zeros(r, dtype=np.int32)
    _lib.check(handle.lib.nfb_row_correlation_host(handle.ptr, c_void_p(comps.ctypes.data), r, n,
                                                   c_void_p(sim32.ctypes.data), c_void_p(dead.ctypes.data)))
    return _finish_similarity(sim32, dead)


def _finish_similarity(sim32: np.ndarray, dead: np.ndarray) -> np.ndarray:
    """Device correlations -> the matrix ``compute_similairty_matrix`` returns (float64, symmetric, clipped)."""
    sim = sim32.astype(np.float64)
    sim = 0.5 * (sim + sim.T)                 # the two triangles come from different tile passes
    # A component that a run drove to all-zero has no variance: corrcoef gives NaN there and scipy's
    # linkage rejects the matrix (the reference crashes).  Treat it as dissimilar to everything.
    if dead.any():
        sim[dead.astype(bool), :] = 0.0
        sim[:, dead.astype(bool)] = 0.0
    np.fill_diagonal(sim, 1.0)
    np.clip(sim, 0, 1, out=sim)
    return sim


def sim2dis(sim: np.ndarray) -> np.ndarray:
    return 1 - sim


def calculate_elbow(merge_dist: np.ndarray) -> float:
    """sso_functions.py:268-300: merge height farthest from the chord between the first and last merge."""
    xs = np.asarray(merge_dist, dtype=np.float64)
    ys = np.arange(xs.shape[0], dtype=np.float64)
    dx, dy = xs[-1] - xs[0], ys[-1] - ys[0]
    chord = np.hypot(dx, dy)
    dist = np.abs(dx * (ys - ys[0]) - dy * (xs - xs[0])) / chord
    return xs[int(np.argmax(dist))]


def cluster_valid(cluster_partition: np.ndarray, dim: int) -> bool:
    clusters, counts = np.unique(cluster_partition, return_counts=True)
    if np.any(counts == 1) or len(clusters) == 1:
        return False
    return clusters.max() <= dim


def clustering_components(dis: np.ndarray, dim: int):
    """sso_functions.py:249-377: average linkage, cut at the elbow, raised until the partition is valid."""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform
    zlink = linkage(squareform(dis, checks=False), method="average")
    elbow = calculate_elbow(zlink[:, 2])
    first = fcluster(zlink, t=elbow, criterion="distance")
    if cluster_valid(first, dim):
        return first
    heights = np.sort(np.unique(zlink))
    for height in heights[np.searchsorted(heights, elbow):]:
        part = fcluster(zlink, t=height, criterion="distance")
        if cluster_valid(part, dim):
            return part
    return None


def idx2centrotype(sim: np.ndarray, partition: np.ndarray) -> np.ndarray:
    """sso_functions.py:405-453: per cluster, the member with the largest summed similarity to the others."""
    n_cluster = int(partition.max())
    out = np.zeros(n_cluster, dtype=int)
    for cluster in range(1, n_cluster + 1):
        members = np.where(partition == cluster)[0]
        sub = sim[np.ix_(members, members)]
        out[cluster - 1] = members[int(np.argmax(sub.sum(axis=0)))]
    return out


class NMFsso:
    """est = NMFsso(fdt_mat, num_int, nmf_params, n_jobs); results = est.run()  (nmfsso.py:62-238)."""

    def __init__(self, fdt_mat, num_int: int, nmf_params: dict, n_jobs: int) -> None:
        self.num_int = num_int
        self.nmf_params = nmf_params.copy()
        self.n_jobs = n_jobs
        self.fdt_mat = fdt_mat
        self.nmf_params["init"] = "random"
        self.col = colours()

    def run(self) -> dict:
        nprint(f"{self.col['pink']}NMF iterations: {self.col['reset']}{self.num_int}")
        owns = not isinstance(self.fdt_mat, DeviceMatrix)
        xdev = DeviceMatrix.from_host(np.asarray(self.fdt_mat)) if owns else self.fdt_mat
        results = {"grey": [], "white": []}
        try:
            for iterat in range(self.num_int):
                nprint(f"{self.col['pink']}Run: {self.col['reset']}{iterat + 1}/{self.num_int}")
                self.nmf_params["random_state"] = None
                state = nmf_decomp(self.nmf_params, xdev)
                results["grey"].append(state["grey_components"])
                results["white"].append(thresholding(state["white_components"], 3))   # nmfsso.py:137
        finally:
            if owns:
                xdev.free()
        return results


def _run_seed(run: int) -> int:
    """64-bit seed of a run's random start: OS entropy like ``random_state=None`` in the reference (nmfsso.py:133);
    NFB_SSO_SEED pins it (base + run index) for reproducible tests."""
    import os
    pinned = os.environ.get("NFB_SSO_SEED")
    if pinned is not None:
        return (int(pinned) * 1000003 + run) & 0xFFFFFFFFFFFFFFFF
    return int.from_bytes(os.urandom(8), "little")


def nmf_sso_resident(xdev: DeviceMatrix, parameters: dict, args: dict) -> dict:
    """``nmf_sso`` with everything between the upload of X and the download of the final factors on the device
    (SURVEY section 8 f1): the N random starts are drawn on the device, every run's factors go from the solver
    state into the resident stacks (white copy z-thresholded there, nmfsso.py:137), the similarity of the stack is
    computed from the stack where it lies, the centrotype rows are gathered on the device into the final solve's
    state.  Host traffic: the [N k, N k] similarity matrix for scipy's linkage, and the final W, H.

    Returns (final_nmf, sim, dis, partitions, centroids)."""
    from .device import NmfState, SsoStacks, make_params
    from .nmf import _check_params, compute_regularization
    col = colours()
    runs = int(args["iterations"])
    p = _check_params(dict(parameters, init="random"))
    k = int(p["n_components"])
    m, n = xdev.shape
    regs = compute_regularization(m, n, p["alpha_W"], p["alpha_H"], p["l1_ratio"])
    nprint(f"{col['pink']}NMF iterations: {col['reset']}{runs}")
    stacks = None
    try:
        try:
            if xdev.has_negative:
                raise ValueError("Negative values in data passed to NMF (input X)")
            stacks = SsoStacks(xdev, k, runs)
            avg = float(np.sqrt(xdev.mean / k))          # sklearn _nmf.py:296
            for run in range(runs):
                nprint(f"{col['pink']}Run: {col['reset']}{run + 1}/{runs}")
                state = NmfState(xdev, k, make_params(p["solver"], p["max_iter"], p["tol"], regs, p["verbose"]))
                try:
                    state.init_random(avg, _run_seed(run))
                    state.run()
                    stacks.add_run(state, run, 3)       # thresholding(components, 3), nmfsso.py:137
                finally:
                    state.free()
        except Exception as e:
            error_and_exit(False, f"Unable to perform NMF due to {e}")
        sim = _finish_similarity(*stacks.similarity())
        dis = sim2dis(sim)
        partitions = clustering_components(dis, args["dim"])
        error_and_exit(partitions is not None and partitions.size != 0,
                       "Clustering Failed. Please check Num of dims")
        centroids = idx2centrotype(sim, partitions)
        parameters["random_state"] = None
        parameters["init"] = "custom"
        parameters["n_components"] = centroids.shape[0]
        kc = int(centroids.shape[0])
        final_nmf = None
        try:
            regs_c = compute_regularization(m, n, p["alpha_W"], p["alpha_H"], p["l1_ratio"])
            state = NmfState(xdev, kc, make_params(p["solver"], p["max_iter"], p["tol"], regs_c, p["verbose"]))
            try:
                stacks.seed_final(state, centroids)
                w_zero, h_zero = state.factors_all_zero()        # sklearn _check_init on a custom start
                if h_zero:
                    raise ValueError("Array passed to NMF (input H) is full of zeros.")
                if w_zero:
                    raise ValueError("Array passed to NMF (input W) is full of zeros.")
                state.run()
                w_fin, h_fin = state.get_factors()
            finally:
                state.free()
            final_nmf = {"grey_components": w_fin, "white_components": h_fin}
        except Exception as e:
            error_and_exit(False, f"Unable to perform NMF due to {e}")
        return final_nmf, sim, dis, partitions, centroids
    finally:
        if stacks is not None:
            stacks.free()


def _reference_sso_outputs(args: dict, sim, dis, partitions, centroids) -> None:
    """The reference ends ``nmf_sso`` with ``nmf_sso_output_wrapper(args['outdir'], ...)`` (nmfsso.py:264-320, 359):
    cluster_stats.csv and the similarity / cluster plots under nfact_decomp/sso_output.  Those writers are plotting
    code outside the hot path; when the reference is importable (the drop-in) its own function is called so the
    tool keeps its output files, otherwise there is nothing to write them with."""
    import sys
    if "outdir" not in args:
        return
    ref = sys.modules.get("NFACT.decomp.decomposition.sso.nmfsso")
    wrapper = getattr(ref, "nmf_sso_output_wrapper", None) if ref is not None else None
    if wrapper is None:
        return
    try:
        wrapper(args["outdir"], sim, dis, partitions, centroids)
    except Exception as e:          # the reference's wrapper reports and carries on as well
        nprint(f"Unable to save graphs due to: {e}")


def nmf_sso(fdt_matrix, parameters: dict, args: dict) -> dict:
    """NFACT/decomp/decomposition/sso/nmfsso.py:322-360.

    Under torchrun (``nfact_b200.multigpu``) the N random-initialised runs are spread over the GPUs -- every rank
    holds its own resident copy of X, the reference's joblib workers over a shared-memory X (nmfsso.py:140-190)
    turned into one process per GPU -- and rank 0 clusters the gathered components and runs the final solve."""
    if args["no_sso"]:
        return nmf_decomp(parameters, fdt_matrix)
    from . import multigpu
    owns = not isinstance(fdt_matrix, DeviceMatrix)
    xdev = None
    try:
        if owns and multigpu.active():
            nprint(f"{colours()['pink']}NMF iterations: {colours()['reset']}{args['iterations']} "
                   f"over {multigpu.context().world} GPUs")
            results, xdev = multigpu.sso_runs_distributed(parameters, np.asarray(fdt_matrix), args["iterations"])
        else:
            import os
            xdev = DeviceMatrix.from_host(np.asarray(fdt_matrix)) if owns else fdt_matrix
            if os.environ.get("NFB_SSO_RESIDENT", "1") != "0":
                # one H2D of X, one D2H of the result: the runs' components never leave the GPU
                final_nmf, sim, dis, partitions, centroids = nmf_sso_resident(xdev, parameters, args)
                _reference_sso_outputs(args, sim, dis, partitions, centroids)
                return final_nmf
            results = NMFsso(xdev, args["iterations"], parameters, args.get("n_cores", 1)).run()
        w_components = np.vstack(results["white"])
        g_components = np.hstack(results["grey"])
        sim = compute_similairty_matrix(w_components)
        dis = sim2dis(sim)
        partitions = clustering_components(dis, args["dim"])
        error_and_exit(partitions is not None and partitions.size != 0,
                       "Clustering Failed. Please check Num of dims")
        centroids = idx2centrotype(sim, partitions)
        parameters["random_state"] = None
        parameters["init"] = "custom"
        parameters["n_components"] = centroids.shape[0]
        w_mat = np.ascontiguousarray(g_components[:, centroids])
        h_mat = np.ascontiguousarray(w_components[centroids, :])
        final_nmf = nmf_decomp(parameters, xdev, W=w_mat, H=h_mat)
        _reference_sso_outputs(args, sim, dis, partitions, centroids)
        return final_nmf
    finally:
        if owns and xdev is not None:
            xdev.free()
