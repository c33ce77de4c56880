"""GPU: solver parity at the sizes sklearn can run (BASELINE.json configs[0] in full, and a long-accumulation
k = 200 case), against sklearn ITSELF run inside the test -- the reference arithmetic of
NFACT/decomp/decomposition/sso/nmfsso.py:54-56 -- from sklearn's own NNDSVD(random_state=0) start.

  C1: 5 000 x 10 000, k = 50, generated as SURVEY.md section 8d says (planted Gamma(0.3) factors, 15 % Bernoulli
      mask, integer streamline counts x 1e8 / waytotal), written as fdt_matrix2.dot text and ingested through
      load_fdt_matrix so the matrix the solver sees went through the product's ingest (bit-exact vs scipy);
      * 20 multiplicative-update iterations, each within 1e-4 relative Frobenius error of sklearn's private
        _multiplicative_update_w / _multiplicative_update_h (sklearn/decomposition/_nmf.py:521-723);
      * the full default coordinate-descent fit (NFACT's get_parameters defaults: alpha_W = 0.1, l1_ratio = 1,
        tol = 1e-4, max_iter = 200): objective within 0.5 %, Hungarian-matched component correlation >= 0.99,
        n_iter within 5 of sklearn.decomposition.NMF's.
  K-length case: 8 209 x 65 551, k = 200 (ragged sizes; accumulation lengths 65 551 and 8 209 -- the tensor core's
      truncating accumulator is what the TMEM chain promotion is there for), 3 iterations of both solvers against
      the CPU oracle, per iteration within 1e-4.
"""
import os
import warnings

import numpy as np
import pytest

from conftest import rel_fro
from test_nmf_gpu import CD_CORR, CD_OBJ_TOL, MU_TOL, matched_correlation

pytestmark = pytest.mark.gpu

M1, N1, K1 = 5000, 10000, 50
WAYTOTAL = 1.5e7


@pytest.fixture(scope="module")
def c1(tmp_path_factory):
    """(X float32 [5000, 10000] from the product's ingest of the .dot text, W0, H0 from sklearn's NNDSVD)."""
    import scipy.sparse as sps
    import nfact_b200
    from sklearn.decomposition._nmf import _initialize_nmf
    rng = np.random.default_rng(0)
    w_star = rng.gamma(0.3, 1.0, (M1, K1))
    h_star = rng.gamma(0.3, 1.0, (K1, N1))
    counts = np.floor(300.0 * (w_star @ h_star)) * (rng.random((M1, N1)) < 0.15)
    rows, cols = np.nonzero(counts)
    vals = counts[rows, cols]
    sub = tmp_path_factory.mktemp("c1") / "sub-01"
    sub.mkdir()
    # fdt_matrix2.dot: 1-based "row col value" lines, last line = shape (NFACT/base/matrix_handling.py:93-113)
    body = np.column_stack([rows + 1, cols + 1, vals.astype(np.int64)])
    with open(sub / "fdt_matrix2.dot", "wb") as fh:
        np.savetxt(fh, body, fmt="%d %d %d")
        fh.write(f"{M1} {N1} 0\n".encode())
    (sub / "waytotal").write_text(f"{int(WAYTOTAL)}\n")
    X = nfact_b200.load_fdt_matrix(str(sub / "fdt_matrix2.dot"))
    # the reference's arithmetic for the same file: csc build, x (1e8 / waytotal) in float64, one rounding
    ref = sps.csc_matrix((vals, (rows, cols)), shape=(M1, N1)).multiply(1e8 / WAYTOTAL).toarray().astype(np.float32)
    assert X.dtype == np.float32 and np.array_equal(X, ref), "C1 ingest is not bit-exact"
    W0, H0 = _initialize_nmf(X, K1, init="nndsvd", random_state=0)
    return X, np.ascontiguousarray(W0, dtype=np.float32), np.ascontiguousarray(H0, dtype=np.float32)


def test_c1_mu_20_iterations_vs_sklearn_private_updates(c1):
    from sklearn.decomposition._nmf import _multiplicative_update_h, _multiplicative_update_w
    from nfact_b200.device import DeviceMatrix, NmfState, make_params
    X, W0, H0 = c1
    regs = (0.1 * N1, 0.1 * M1, 0.0, 0.0)                 # NFACT defaults: alpha_W = 0.1, alpha_H = same, l1_ratio = 1
    xdev = DeviceMatrix.from_host(X)
    st = NmfState(xdev, K1, make_params("mu", 200, 0.0, regs))
    st.set_factors(W0, H0)
    W, H = W0.copy(), H0.copy()
    worst = 0.0
    for it in range(1, 21):
        W, *_ = _multiplicative_update_w(X, W, H, 2, regs[0], regs[2], 1.0)
        H = _multiplicative_update_h(X, W, H, 2, regs[1], regs[3], 1.0)
        st.iterate(1)
        Wd, Hd = st.get_factors()
        ew, eh = rel_fro(Wd, W), rel_fro(Hd, H)
        worst = max(worst, ew, eh)
        assert ew < MU_TOL and eh < MU_TOL, (it, ew, eh)
    err = st.error()
    ref_err = np.linalg.norm(X.astype(np.float64) - W.astype(np.float64) @ H.astype(np.float64))
    assert abs(err - ref_err) <= 1e-5 * ref_err
    st.free()
    xdev.free()
    print(f"C1 MU: worst per-iteration relative error over 20 iterations {worst:.2e}")


def test_c1_cd_default_fit_vs_sklearn_nmf(c1):
    import nfact_b200
    from sklearn.decomposition import NMF
    from oracle import nfact_oracle as orc
    X, W0, H0 = c1
    p = nfact_b200.get_parameters(None, "nmf", K1)                         # NFACT's defaults, verbatim
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = NMF(**dict(p, init="custom"))
        Wr = ref.fit_transform(X, W=W0.copy(), H=H0.copy())
    Hr = ref.components_
    est = nfact_b200.NMF(**dict(p, init="custom"))
    W = est.fit_transform(X, W=W0.copy(), H=H0.copy())
    H = est.components_
    regs = orc.compute_regularization(M1, N1, p["alpha_W"], p["alpha_H"], p["l1_ratio"])
    obj, obj_ref = orc.nmf_objective(X, W, H, *regs), orc.nmf_objective(X, Wr, Hr, *regs)
    corr_h, live_h = matched_correlation(H, Hr)
    corr_w, live_w = matched_correlation(W.T, Wr.T)
    print(f"C1 CD: n_iter {est.n_iter_} (sklearn {ref.n_iter_}), objective rel diff {abs(obj - obj_ref) / obj_ref:.2e}, "
          f"min matched corr H {corr_h[live_h].min():.5f} W {corr_w[live_w].min():.5f}, "
          f"err {est.reconstruction_err_:.6g} (sklearn {ref.reconstruction_err_:.6g})")
    assert abs(obj - obj_ref) <= CD_OBJ_TOL * obj_ref
    assert corr_h[live_h].min() >= CD_CORR and corr_w[live_w].min() >= CD_CORR
    assert abs(est.n_iter_ - ref.n_iter_) <= 5
    assert abs(est.reconstruction_err_ - ref.reconstruction_err_) <= 1e-3 * ref.reconstruction_err_


def test_c1_nndsvd_on_device_vs_sklearn(c1):
    """NFACT's default init on the device (15 passes over X through the split GEMM) against sklearn's, same seed."""
    from nfact_b200.device import DeviceMatrix
    from nfact_b200.nndsvd import nndsvd_device
    X, W0, H0 = c1
    xdev = DeviceMatrix.from_host(X)
    W, H = nndsvd_device(xdev, K1, np.random.RandomState(0))
    xdev.free()
    assert rel_fro(W, W0) < 1e-3 and rel_fro(H, H0) < 1e-3


@pytest.mark.parametrize("solver", ["mu", "cd"])
def test_k200_long_accumulation_vs_oracle(solver):
    from nfact_b200.device import DeviceMatrix, NmfState, make_params
    from oracle import nfact_oracle as orc
    m, n, k = 8209, 65551, 200
    rng = np.random.default_rng(21)
    w_star = rng.gamma(0.3, 1.0, (m, 40)).astype(np.float32)
    h_star = rng.gamma(0.3, 1.0, (40, n)).astype(np.float32)
    X = np.floor(300.0 * (w_star @ h_star))
    X *= rng.random((m, n), dtype=np.float32) < 0.05
    X *= np.float32(1e8 / WAYTOTAL)
    W0, H0 = orc.init_random(X, k, 1)
    regs = orc.compute_regularization(m, n, 0.1, "same", 1.0) if solver == "cd" else (0.0, 0.0, 0.0, 0.0)
    rec = []
    (orc.fit_mu if solver == "mu" else orc.fit_cd)(X, W0, H0, *regs, tol=0, max_iter=3, record=rec)
    xdev = DeviceMatrix.from_host(X)
    st = NmfState(xdev, k, make_params(solver, 200, 0.0, regs))
    st.set_factors(W0, H0)
    for it in range(1, 4):
        st.iterate(1)
        W, H = st.get_factors()
        ew, eh = rel_fro(W, rec[it - 1][1]), rel_fro(H, rec[it - 1][2])
        print(f"k=200 {solver} iteration {it}: relW {ew:.2e} relH {eh:.2e}")
        assert ew < MU_TOL and eh < MU_TOL, (solver, it, ew, eh)
    st.free()
    xdev.free()
