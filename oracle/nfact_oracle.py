"""CPU oracle for the NFACT decomposition hot path.

TEST INFRASTRUCTURE ONLY.  This module is a plain numpy restatement of the
reference's algorithm for the path named in BASELINE.json (ingest -> NMF solve ->
NNLS dual regression).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package ``nfact_b200`` never does (it fails loudly without its CUDA library).

Where the arithmetic lives.  NFACT is pure Python; every number on the hot path is
produced by third-party code it calls, none of which is vendored in /root/reference
(``pyproject.toml:20-31`` lists them without version pins):

  * scikit-learn (1.9.0 in this image) ``sklearn/decomposition/_nmf.py`` and
    ``_cdnmf_fast.pyx`` -- call site ``NFACT/decomp/decomposition/sso/nmfsso.py:54-56``
  * scipy (1.18.1 in this image) ``scipy/optimize/_nnls.py`` (Lawson-Hanson) -- call
    sites ``NFACT/dual_reg/dual_regression/dual_regression_methods.py:139,153,264``;
    ``scipy.sparse`` -- call site ``NFACT/base/matrix_handling.py:137``

Pinning.  The reference's own tests hold no golden vectors for this path
(``NFACT/testing/nfact_pipeline_test.py`` asserts ``isinstance(..., np.ndarray)`` only),
so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build
container: ``tests/golden/make_golden.py`` imports ``/root/reference`` (NFACT's own
``load_fdt_matrix``/``avg_fdt``/``nmf_decomp``/``nmf_dual_regression`` on top of the
installed sklearn/scipy) and commits the vectors under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against them, and
``tests/test_oracle_live_cpu.py`` runs the solvers, ``nnls`` and the random init against the
installed sklearn / scipy themselves on random problems (both boxes have them).

Each function cites the reference (or sklearn/scipy) file:line it follows.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

EPSILON = np.finfo(np.float32).eps  # sklearn/decomposition/_nmf.py:32

_HERE = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------------------
# a1-a5  ingest
# --------------------------------------------------------------------------------------
def read_dot(path: str) -> np.ndarray:
    """NFACT/base/matrix_handling.py:93-113 ``read_in_matrix`` (plain ``.dot`` branch).

    Whitespace separated text, 3 columns ``row col value`` (1-based); the last line holds
    ``nrows ncols <ignored>``.  Returns float64 [T+1, 3] like ``np.loadtxt``.
    """
    with open(path, "rb") as fh:
        flat = np.array(fh.read().split(), dtype=np.float64)
    return flat.reshape(-1, 3)


def read_waytotal(path: str) -> float:
    """NFACT/base/matrix_handling.py:89 ``float(np.loadtxt(waytotal))``."""
    with open(path, "r") as fh:
        return float(fh.read().split()[0])


def coo_sum_scale(triplets: np.ndarray, waytotal: float):
    """NFACT/base/matrix_handling.py:116-140 ``load_single_matrix`` + :71-90 ``normalise_matrix``.

    csc_matrix((data, (rows-1, cols-1)), shape=last line) sums duplicate coordinates in
    float64; the result is multiplied by ``1e8 / waytotal`` in float64.
    Returns (flat_index int64 [U] sorted, value float64 [U], nrows, ncols).
    The duplicate sum is order-free only for values whose partial sums are exact
    (integer streamline counts); probtrackx ``--pd`` output has no duplicates.
    """
    nrows = int(triplets[-1, 0])
    ncols = int(triplets[-1, 1])
    rows = np.array(triplets[:-1, 0] - 1, dtype=np.int64)
    cols = np.array(triplets[:-1, 1] - 1, dtype=np.int64)
    data = triplets[:-1, 2]
    if rows.size and (rows.min() < 0 or cols.min() < 0 or rows.max() >= nrows or cols.max() >= ncols):
        raise ValueError("row/column index exceeds matrix dimensions")
    flat = rows * ncols + cols
    order = np.argsort(flat, kind="stable")
    flat_sorted = flat[order]
    uniq, start = np.unique(flat_sorted, return_index=True)
    summed = np.add.reduceat(data[order], start) if flat.size else np.zeros(0)
    scale = 1e8 / waytotal
    return uniq, summed * scale, nrows, ncols


def load_fdt_matrix(matfile: str, waytotal: float | None = None) -> np.ndarray:
    """NFACT/base/matrix_handling.py:143-158: ``load_single_matrix(...).toarray().astype(float32)``."""
    if waytotal is None:
        waytotal = read_waytotal(os.path.join(os.path.dirname(matfile), "waytotal"))
    idx, val, nrows, ncols = coo_sum_scale(read_dot(matfile), waytotal)
    dense = np.zeros(nrows * ncols, dtype=np.float64)
    dense[idx] = val
    return dense.reshape(nrows, ncols).astype(np.float32)


def avg_fdt(matfiles: list, waytotals: list | None = None) -> np.ndarray:
    """NFACT/decomp/decomposition/matrix_handling.py:73-100 ``avg_fdt``.

    ``acc = mat_0; acc = acc + mat_s`` in list order (float64), ``acc.data *= 1.0/len``
    (scipy/sparse/_data.py:61-65: multiply by the reciprocal, not a division), densify,
    round once to float32.
    """
    acc = None
    for pos, matfile in enumerate(matfiles):
        way = waytotals[pos] if waytotals is not None else read_waytotal(
            os.path.join(os.path.dirname(matfile), "waytotal"))
        idx, val, nrows, ncols = coo_sum_scale(read_dot(matfile), way)
        dense = np.zeros(nrows * ncols, dtype=np.float64)
        dense[idx] = val
        acc = dense if acc is None else acc + dense
    acc *= 1.0 / len(matfiles)
    return acc.reshape(nrows, ncols).astype(np.float32)


# --------------------------------------------------------------------------------------
# a6, a9  hyper-parameters
# --------------------------------------------------------------------------------------
def get_parameters(parameters, algo: str, n_components: int) -> dict:
    """NFACT/decomp/decomposition/matrix_decomposition.py:74-107 (NMF branch) with the
    sklearn 1.9.0 ``NMF.__init__`` defaults that ``create_combined_algo_dict``
    (NFACT/config/nfact_config_functions.py:337-356) introspects."""
    if parameters:
        parameters["n_components"] = n_components
        return parameters
    if algo != "nmf":
        raise ValueError("oracle covers the NMF branch only")
    return {
        "init": "nndsvd", "solver": "cd", "beta_loss": "frobenius", "tol": 1e-4,
        "max_iter": 200, "random_state": None, "alpha_W": 0.1, "alpha_H": "same",
        "l1_ratio": 1, "verbose": 0, "shuffle": False, "n_components": n_components,
    }


def compute_regularization(n_samples, n_features, alpha_W, alpha_H, l1_ratio):
    """sklearn/decomposition/_nmf.py:1249-1259."""
    alpha_H = alpha_W if alpha_H == "same" else alpha_H
    l1_reg_W = n_features * alpha_W * l1_ratio
    l1_reg_H = n_samples * alpha_H * l1_ratio
    l2_reg_W = n_features * alpha_W * (1.0 - l1_ratio)
    l2_reg_H = n_samples * alpha_H * (1.0 - l1_ratio)
    return l1_reg_W, l1_reg_H, l2_reg_W, l2_reg_H


# --------------------------------------------------------------------------------------
# a8  initialisation
# --------------------------------------------------------------------------------------
def init_random(X: np.ndarray, k: int, seed: int):
    """sklearn/decomposition/_nmf.py:296-307 (``init='random'``)."""
    rng = np.random.RandomState(seed)
    avg = np.sqrt(X.mean() / k)
    H = avg * rng.standard_normal(size=(k, X.shape[1])).astype(X.dtype, copy=False)
    W = avg * rng.standard_normal(size=(X.shape[0], k)).astype(X.dtype, copy=False)
    return np.abs(W), np.abs(H)


def randomized_svd(M: np.ndarray, k: int, seed: int, n_oversamples: int = 10):
    """sklearn/utils/extmath.py:313-383 + :560-630 (``_randomized_svd`` with the defaults
    ``_initialize_nmf`` uses: n_iter='auto', LU-normalised power iterations, transpose='auto',
    flip_sign=True)."""
    from scipy import linalg

    rng = np.random.RandomState(seed)
    n_random = k + n_oversamples
    n_samples, n_features = M.shape
    n_iter = 7 if k < 0.1 * min(M.shape) else 4
    transpose = n_samples < n_features
    if transpose:
        M = M.T
    Q = rng.normal(size=(M.shape[1], n_random))
    if M.dtype == np.float32:
        Q = Q.astype(np.float32, copy=False)
    if n_iter <= 2:
        normalizer = lambda x: (x, None)  # noqa: E731
    else:
        normalizer = lambda x: linalg.lu(x, permute_l=True, check_finite=False)  # noqa: E731
    for _ in range(n_iter):
        Q, _ = normalizer(M @ Q)
        Q, _ = normalizer(M.T @ Q)
    Q, _ = linalg.qr(M @ Q, mode="economic", check_finite=False)
    B = Q.T @ M
    Uhat, s, Vt = linalg.svd(B, full_matrices=False, lapack_driver="gesdd")
    U = Q @ Uhat
    if not transpose:
        piv = np.argmax(np.abs(U), axis=0)
        signs = np.sign(U[piv, np.arange(U.shape[1])])
    else:
        piv = np.argmax(np.abs(Vt), axis=1)
        signs = np.sign(Vt[np.arange(Vt.shape[0]), piv])
    U = U * signs[np.newaxis, :]
    Vt = Vt * signs[:, np.newaxis]
    if transpose:
        return Vt[:k, :].T, s[:k], U[:, :k].T
    return U[:, :k], s[:k], Vt[:k, :]


def nndsvd_from_svd(U, S, V, eps: float = 1e-6):
    """sklearn/decomposition/_nmf.py:309-347: Boutsidis-Gallopoulos sign split, then values
    below ``eps`` set to zero (``init='nndsvd'``)."""
    k = U.shape[1]
    W = np.zeros_like(U)
    H = np.zeros_like(V)
    W[:, 0] = np.sqrt(S[0]) * np.abs(U[:, 0])
    H[0, :] = np.sqrt(S[0]) * np.abs(V[0, :])

    def nrm(x):
        return math.sqrt(float(np.dot(x, x)))

    for j in range(1, k):
        x, y = U[:, j], V[j, :]
        x_p, y_p = np.maximum(x, 0), np.maximum(y, 0)
        x_n, y_n = np.abs(np.minimum(x, 0)), np.abs(np.minimum(y, 0))
        x_p_nrm, y_p_nrm = nrm(x_p), nrm(y_p)
        x_n_nrm, y_n_nrm = nrm(x_n), nrm(y_n)
        m_p, m_n = x_p_nrm * y_p_nrm, x_n_nrm * y_n_nrm
        if m_p > m_n:
            u, v, sigma = x_p / x_p_nrm, y_p / y_p_nrm, m_p
        else:
            u, v, sigma = x_n / x_n_nrm, y_n / y_n_nrm, m_n
        lbd = np.sqrt(S[j] * sigma)
        W[:, j] = lbd * u
        H[j, :] = lbd * v
    W[W < eps] = 0
    H[H < eps] = 0
    return W, H


def init_nndsvd(X: np.ndarray, k: int, seed: int):
    """sklearn/decomposition/_nmf.py:309-347 end to end."""
    U, S, V = randomized_svd(X, k, seed)
    return nndsvd_from_svd(U, S, V)


# --------------------------------------------------------------------------------------
# a11, a12  multiplicative update (beta_loss='frobenius')
# --------------------------------------------------------------------------------------
def frob_error(X, W, H) -> float:
    """sklearn/decomposition/_nmf.py:78-182 with beta=2, square_root=True: ||X - WH||_F."""
    diff = X - np.dot(W, H)
    return float(np.sqrt(np.dot(diff.ravel(), diff.ravel()) / 2.0 * 2))


def mu_update_w(X, W, H, l1_reg_W, l2_reg_W):
    """sklearn/decomposition/_nmf.py:521-626 (beta_loss == 2 branch, gamma == 1)."""
    numerator = np.dot(X, H.T)
    HHt = np.dot(H, H.T)
    denominator = np.dot(W, HHt)
    if l1_reg_W > 0:
        denominator += l1_reg_W
    if l2_reg_W > 0:
        denominator = denominator + l2_reg_W * W
    denominator[denominator == 0] = EPSILON
    numerator /= denominator
    W *= numerator
    return W


def mu_update_h(X, W, H, l1_reg_H, l2_reg_H):
    """sklearn/decomposition/_nmf.py:629-723 (beta_loss == 2 branch, gamma == 1)."""
    numerator = np.dot(W.T, X)
    denominator = np.linalg.multi_dot([W.T, W, H])
    if l1_reg_H > 0:
        denominator += l1_reg_H
    if l2_reg_H > 0:
        denominator = denominator + l2_reg_H * H
    denominator[denominator == 0] = EPSILON
    numerator /= denominator
    H *= numerator
    return H


def fit_mu(X, W0, H0, l1W, l1H, l2W, l2H, tol=1e-4, max_iter=200, record=None):
    """sklearn/decomposition/_nmf.py:726-897 ``_fit_multiplicative_update`` (beta=2).

    ``record`` (optional list) receives (n_iter, W.copy(), H.copy()) after every iteration.
    """
    W, H = W0.copy(), H0.copy()
    error_at_init = frob_error(X, W, H)
    previous_error = error_at_init
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        W = mu_update_w(X, W, H, l1W, l2W)
        H = mu_update_h(X, W, H, l1H, l2H)
        if record is not None:
            record.append((n_iter, W.copy(), H.copy()))
        if tol > 0 and n_iter % 10 == 0:
            error = frob_error(X, W, H)
            if (previous_error - error) / error_at_init < tol:
                break
            previous_error = error
    return W, H, n_iter


# --------------------------------------------------------------------------------------
# a10  coordinate descent (HALS)
# --------------------------------------------------------------------------------------
def _clib():
    """Build (once) and load the plain-C restatement of the two sequential kernels."""
    so = os.path.join(_HERE, "liboracle_c.so")
    src = os.path.join(_HERE, "oracle_c.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src, "-lm"])
    lib = ctypes.CDLL(so)
    lib.oracle_cd_sweep_f32.restype = ctypes.c_float
    lib.oracle_cd_sweep_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_long, ctypes.c_long]
    lib.oracle_nnls_f64.restype = ctypes.c_int
    lib.oracle_nnls_f64.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_long,
                                    ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p]
    return lib


def cd_half_update(X, W, Ht, l1_reg, l2_reg) -> float:
    """sklearn/decomposition/_nmf.py:369-396 ``_update_coordinate_descent`` (shuffle=False)
    followed by ``_cdnmf_fast.pyx:8-38``.  Updates ``W`` in place, returns the violation."""
    k = Ht.shape[1]
    HHt = np.ascontiguousarray(np.dot(Ht.T, Ht))
    XHt = np.ascontiguousarray(np.dot(X, Ht))
    if l2_reg != 0.0:
        HHt.flat[:: k + 1] += l2_reg
    if l1_reg != 0.0:
        XHt -= l1_reg
    assert W.dtype == np.float32 and W.flags.c_contiguous
    lib = _clib()
    return float(lib.oracle_cd_sweep_f32(W.ctypes.data, HHt.ctypes.data, XHt.ctypes.data,
                                         W.shape[0], k))


def fit_cd(X, W0, H0, l1W, l1H, l2W, l2H, tol=1e-4, max_iter=200, record=None):
    """sklearn/decomposition/_nmf.py:399-518 ``_fit_coordinate_descent`` (shuffle=False)."""
    W = np.ascontiguousarray(W0.copy())
    Ht = np.ascontiguousarray(H0.T.copy())
    Xt = np.ascontiguousarray(X.T)
    violation_init = 0.0
    n_iter = 0
    for n_iter in range(1, max_iter + 1):
        violation = 0.0
        violation += cd_half_update(X, W, Ht, l1W, l2W)
        violation += cd_half_update(Xt, Ht, W, l1H, l2H)
        if record is not None:
            record.append((n_iter, W.copy(), Ht.T.copy(), violation))
        if n_iter == 1:
            violation_init = violation
        if violation_init == 0:
            break
        if violation / violation_init <= tol:
            break
    return W, Ht.T, n_iter


def nmf_objective(X, W, H, l1W, l1H, l2W, l2H) -> float:
    """The function both solvers minimise (sklearn NMF docstring): 0.5||X-WH||^2 + l1 + 0.5 l2."""
    diff = X.astype(np.float64) - W.astype(np.float64) @ H.astype(np.float64)
    obj = 0.5 * float((diff * diff).sum())
    obj += l1W * float(np.abs(W).sum()) + l1H * float(np.abs(H).sum())
    obj += 0.5 * l2W * float((W.astype(np.float64) ** 2).sum())
    obj += 0.5 * l2H * float((H.astype(np.float64) ** 2).sum())
    return obj


def nmf_decomp(parameters: dict, X: np.ndarray, W=None, H=None) -> dict:
    """NFACT/decomp/decomposition/sso/nmfsso.py:32-59 on top of the restated solver."""
    p = dict(parameters)
    if W is not None and H is not None:
        p["init"] = "custom"
    k = p["n_components"]
    if p.get("beta_loss", "frobenius") not in ("frobenius", 2):
        raise ValueError("oracle restates beta_loss='frobenius' only")
    if p["init"] == "custom":
        W0, H0 = W, H
    elif p["init"] == "random":
        W0, H0 = init_random(X, k, p["random_state"])
    elif p["init"] == "nndsvd":
        W0, H0 = init_nndsvd(X, k, p["random_state"])
    else:
        raise ValueError(f"oracle does not restate init={p['init']!r}")
    regs = compute_regularization(X.shape[0], X.shape[1], p["alpha_W"], p["alpha_H"], p["l1_ratio"])
    fit = fit_cd if p["solver"] == "cd" else fit_mu
    Wf, Hf, n_iter = fit(X, W0, H0, *regs, tol=p["tol"], max_iter=p["max_iter"])
    return {"grey_components": Wf, "white_components": Hf, "n_iter": n_iter}


# --------------------------------------------------------------------------------------
# a14  post-ops
# --------------------------------------------------------------------------------------
def thresholding(component: np.ndarray, zscore_val) -> np.ndarray:
    """NFACT/base/matrix_handling.py:218-238 (in place, per row)."""
    for comp in range(component.shape[0]):
        tract = component[comp, :]
        threshold = tract.mean() + (zscore_val * tract.std())
        tract[tract < threshold] = 0.0
    return component


def normalise_components(grey: np.ndarray, white: np.ndarray) -> dict:
    """NFACT/base/matrix_handling.py:45-68: StandardScaler (population std, zero std -> 1)
    on the columns of grey [m,k] and on the rows of white [k,n]."""
    def zscore_cols(a):
        a64 = a.astype(np.float64)
        mean = a64.mean(axis=0)
        std = a64.std(axis=0)
        std[std < 10 * np.finfo(np.float64).eps] = 1.0
        return ((a64 - mean) / std).astype(a.dtype)
    return {"grey_matter": zscore_cols(grey), "white_matter": zscore_cols(white.T).T}


def create_wta_map(component: np.ndarray, axis: int, z_thr: float) -> np.ndarray:
    """NFACT/decomp/pipes/image_handling.py:216-245: winner-takes-all map.  StandardScaler z-scores every COLUMN of
    ``component`` (float64 statistics; ``X -= mean_; X /= scale_`` each rounded to the array's dtype,
    sklearn/preprocessing/_data.py StandardScaler.transform; a constant column gets scale 1,
    ``_is_constant_feature`` + ``_handle_zeros_in_scale``), then the 1-based index of the first maximum along
    ``axis`` (0: white components [k, n]; 1: grey components [m, k]), 0 where that maximum is below ``z_thr``
    (compared in the array's dtype)."""
    a = np.asarray(component)
    dtype = a.dtype if a.dtype in (np.float32, np.float64) else np.dtype(np.float64)
    a64 = a.astype(np.float64)
    n = a.shape[0]
    mean = a64.sum(axis=0) / n
    var = ((a64 - mean) ** 2).sum(axis=0) / n
    eps = np.finfo(np.float64).eps
    scale = np.sqrt(var)
    constant = var <= n * eps * var + (n * mean * eps) ** 2
    scale[constant | (scale < 10 * eps)] = 1.0
    z = (a.astype(dtype) - mean).astype(dtype)
    z = (z / scale).astype(dtype)
    top = np.max(z, axis=axis, keepdims=True)
    wta = np.argmax(z, axis=axis, keepdims=True) + 1
    wta[top < dtype.type(z_thr)] = 0
    return np.array(wta, dtype=int)


def threshold_grey_components(components: np.ndarray, seeds_id: np.ndarray, n_seeds: int, zscore_val) -> np.ndarray:
    """NFACT/base/matrix_handling.py:184-215: ``thresholding`` of the rows of every seed separately, in place
    (``seeds_id`` = second-to-last column of coords_for_fdt_matrix2; the reference loops ``enumerate(seeds)``)."""
    for idx in range(n_seeds):
        rows = seeds_id == idx
        per_seed = components[rows, :].T
        thresholding(per_seed, zscore_val)
        components[rows, :] = per_seed.T
    return components


def thresholding_components(threshold_val: int, seeds_id: np.ndarray, n_seeds: int, components: dict) -> dict:
    """NFACT/base/matrix_handling.py:241-277."""
    if threshold_val != 0:
        components["white_components"] = thresholding(components["white_components"], threshold_val)
        components["grey_components"] = threshold_grey_components(components["grey_components"], seeds_id, n_seeds,
                                                                  threshold_val)
    return components


# --------------------------------------------------------------------------------------
# a15  non-negative dual regression
# --------------------------------------------------------------------------------------
def nnls(A: np.ndarray, b: np.ndarray, maxiter: int | None = None) -> np.ndarray:
    """scipy/optimize/_nnls.py ``nnls`` -> compiled Lawson-Hanson active-set solver
    (Lawson & Hanson, "Solving Least Squares Problems", 1995, ch. 23), maxiter = 3 n.
    A and b are converted to float64 on every call exactly like scipy does.
    Raises RuntimeError when the iteration cap is hit (scipy's ``info == 3``)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    m, n = A.shape
    if not maxiter:
        maxiter = 3 * n
    x = np.zeros(n, dtype=np.float64)
    rnorm = ctypes.c_double(0.0)
    info = _clib().oracle_nnls_f64(A.ctypes.data, b.ctypes.data, m, n, maxiter,
                                   x.ctypes.data, ctypes.byref(rnorm))
    if info == 3:
        raise RuntimeError("Maximum number of iterations reached.")
    return x


def nmf_dual_regression(components: dict, X: np.ndarray) -> dict:
    """NFACT/dual_reg/dual_regression/dual_regression_methods.py:119-167 ``nnls_non_parallel``."""
    G = components["grey_components"]
    Hs = np.array([nnls(G, X[:, j]) for j in range(X.shape[1])])
    Gs = np.array([nnls(Hs, X.T[:, i]) for i in range(X.shape[0])])
    return {"grey_components": Gs, "white_components": Hs.T}
