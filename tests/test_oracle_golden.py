"""Pin the CPU oracle (oracle/) against the golden vectors produced by the reference itself
(tests/golden/make_golden.py, run in the build container).  CPU only."""
import os

import numpy as np
import pytest

from oracle import nfact_oracle as orc
from conftest import GOLDEN_DIR, rel_fro


def test_ingest_single_bit_exact(golden, golden_subjects):
    for s, path in enumerate(golden_subjects):
        X = orc.load_fdt_matrix(path)
        assert X.dtype == np.float32
        assert np.array_equal(X, golden[f"ingest_single_{s + 1}"])
    Xpd = orc.load_fdt_matrix(os.path.join(GOLDEN_DIR, "data", "sub-pd", "fdt_matrix2.dot"))
    assert np.array_equal(Xpd, golden["ingest_single_pd"])


def test_ingest_group_bit_exact(golden, golden_subjects):
    assert np.array_equal(orc.avg_fdt(golden_subjects), golden["ingest_group"])
    assert np.array_equal(orc.avg_fdt(golden_subjects[:2]), golden["ingest_group2"])


def test_fixture_properties(golden):
    """The inputs keep what the reference's own test data exercises: duplicate coordinates,
    leading all-zero target columns, shape from the last line."""
    X = golden["ingest_group"]
    assert X.shape == (99, 199)
    assert not X[:, :100].any() and X[:, 100:].any()


def test_parameters_and_regularization(golden):
    p = orc.get_parameters(None, "nmf", 10)
    assert sorted(p.keys()) == list(golden["params_keys"])
    assert repr(sorted(p.items(), key=lambda kv: kv[0])) == str(golden["params_repr"])
    regs = orc.compute_regularization(99, 199, p["alpha_W"], p["alpha_H"], p["l1_ratio"])
    assert np.array_equal(np.array(regs, dtype=np.float64), golden["regularization"])


def test_init_random_bit_exact(golden):
    W, H = orc.init_random(golden["ingest_group"], 10, 7)
    assert np.array_equal(W, golden["init_random_W"]) and np.array_equal(H, golden["init_random_H"])


def test_init_nndsvd(golden):
    W, H = orc.init_nndsvd(golden["ingest_group"], 10, 0)
    assert rel_fro(W, golden["init_nndsvd_W"]) < 1e-5
    assert rel_fro(H, golden["init_nndsvd_H"]) < 1e-5


def test_mu_trajectory(golden):
    X = golden["ingest_group"]
    regs = golden["regularization"]
    rec = []
    orc.fit_mu(X, golden["init_nndsvd_W"], golden["init_nndsvd_H"], *regs, tol=0, max_iter=40, record=rec)
    assert abs(orc.frob_error(X, golden["init_nndsvd_W"], golden["init_nndsvd_H"]) - golden["mu_err0"]) \
        <= 1e-6 * golden["mu_err0"]
    for it, W, H in rec:
        if f"mu_W_{it}" in golden:
            assert rel_fro(W, golden[f"mu_W_{it}"]) < 1e-6, it
            assert rel_fro(H, golden[f"mu_H_{it}"]) < 1e-6, it


def test_mu_l2(golden):
    X = golden["ingest_group"]
    l1W, l1H, l2W, l2H = golden["mu_l2_regs"]
    rec = []
    orc.fit_mu(X, golden["init_random_W"], golden["init_random_H"], l1W, l1H, l2W, l2H,
               tol=0, max_iter=5, record=rec)
    assert rel_fro(rec[-1][1], golden["mu_l2_W_5"]) < 1e-6
    assert rel_fro(rec[-1][2], golden["mu_l2_H_5"]) < 1e-6


def test_mu_full_fit(golden):
    X = golden["ingest_group"]
    p = dict(orc.get_parameters(None, "nmf", 10), solver="mu")
    res = orc.nmf_decomp(p, X, W=golden["init_nndsvd_W"], H=golden["init_nndsvd_H"])
    assert res["n_iter"] == int(golden["mu_fit_n_iter"])
    assert rel_fro(res["grey_components"], golden["mu_fit_W"]) < 1e-5
    assert rel_fro(res["white_components"], golden["mu_fit_H"]) < 1e-5


def test_cd_sweeps(golden):
    X = golden["ingest_group"]
    regs = golden["regularization"]
    rec = []
    orc.fit_cd(X, golden["init_nndsvd_W"], golden["init_nndsvd_H"], *regs, tol=0, max_iter=3, record=rec)
    for it, W, H, viol in rec:
        assert rel_fro(W, golden[f"cd_W_{it}"]) < 1e-5, it
        assert rel_fro(H, golden[f"cd_H_{it}"]) < 1e-5, it
        assert abs(viol - golden["cd_violation"][it - 1]) <= 1e-4 * golden["cd_violation"][it - 1]


def test_cd_full_fit(golden):
    X = golden["ingest_group"]
    p = orc.get_parameters(None, "nmf", 10)
    res = orc.nmf_decomp(p, X, W=golden["init_nndsvd_W"], H=golden["init_nndsvd_H"])
    assert abs(res["n_iter"] - int(golden["cd_fit_n_iter"])) <= 2
    regs = golden["regularization"]
    obj = orc.nmf_objective(X, res["grey_components"], res["white_components"], *regs)
    ref = orc.nmf_objective(X, golden["cd_fit_W"], golden["cd_fit_H"], *regs)
    assert abs(obj - ref) <= 1e-4 * ref
    assert rel_fro(res["white_components"], golden["cd_fit_H"]) < 1e-2
    assert abs(orc.frob_error(X, golden["cd_fit_W"], golden["cd_fit_H"]) - golden["cd_fit_err"]) \
        <= 1e-5 * golden["cd_fit_err"]


def test_postops(golden):
    norm = orc.normalise_components(golden["cd_fit_W"].copy(), golden["cd_fit_H"].copy())
    assert np.allclose(norm["grey_matter"], golden["norm_grey"], rtol=1e-5, atol=1e-6)
    assert np.allclose(norm["white_matter"], golden["norm_white"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(orc.thresholding(golden["cd_fit_H"].copy(), 3), golden["thresh_white_3"])
    assert np.array_equal(orc.thresholding(golden["mu_fit_H"].copy(), 1), golden["thresh_white_1"])


def test_nnls_known_answers(golden):
    # scipy/optimize/_nnls.py docstring examples
    A = np.array([[1, 0], [1, 0], [0, 1]], dtype=float)
    assert np.allclose(orc.nnls(A, np.array([2.0, 1.0, 1.0])), [1.5, 1.0])
    assert np.allclose(orc.nnls(A, np.array([-1.0, -1.0, -1.0])), [0.0, 0.0])
    A, B = golden["nnls_A"], golden["nnls_B"]
    for j in range(B.shape[1]):
        assert np.allclose(orc.nnls(A, B[:, j]), golden["nnls_X"][j], rtol=1e-9, atol=1e-11)


def test_dual_regression(golden):
    comps = {"grey_components": golden["cd_fit_W"].astype(np.float64)}
    dr = orc.nmf_dual_regression(comps, golden["ingest_single_1"])
    assert dr["grey_components"].shape == golden["dr_grey"].shape
    assert rel_fro(dr["grey_components"], golden["dr_grey"]) < 1e-8
    assert rel_fro(dr["white_components"], golden["dr_white"]) < 1e-8
    dr2 = orc.nmf_dual_regression({"grey_components": golden["mu_fit_W"]}, golden["ingest_single_2"])
    assert rel_fro(dr2["grey_components"], golden["dr2_grey"]) < 1e-8
    assert rel_fro(dr2["white_components"], golden["dr2_white"]) < 1e-8


def test_oracle_postops_match_reference():
    """a14: the oracle's thresholding / threshold_grey_components / thresholding_components / normalise_components
    against outputs of the reference's own functions (tests/golden/make_golden_postops.py)."""
    import os
    from conftest import GOLDEN_DIR
    from oracle import nfact_oracle as orc
    g = dict(np.load(os.path.join(GOLDEN_DIR, "golden_postops.npz")))
    seeds_id, n_seeds = g["seeds_id"], int(g["n_seeds"])
    for tag in ("f32", "f64"):
        grey, white = g[f"grey_{tag}"], g[f"white_{tag}"]
        for z in (1, 3):
            assert np.array_equal(orc.thresholding(white.copy(), z), g[f"white_thr{z}_{tag}"])
            assert np.array_equal(orc.threshold_grey_components(grey.copy(), seeds_id, n_seeds, z),
                                  g[f"grey_thr{z}_{tag}"])
        comps = orc.thresholding_components(2, seeds_id, n_seeds, {"grey_components": grey.copy(),
                                                                   "white_components": white.copy()})
        assert np.array_equal(comps["grey_components"], g[f"grey_thr2_{tag}"])
        assert np.array_equal(comps["white_components"], g[f"white_thr2_{tag}"])
        norm = orc.normalise_components(grey.copy(), white.copy())
        tol = dict(rtol=2e-6, atol=2e-6) if tag == "f32" else dict(rtol=1e-12, atol=1e-12)
        assert np.allclose(norm["grey_matter"], g[f"grey_norm_{tag}"], **tol)
        assert np.allclose(norm["white_matter"], g[f"white_norm_{tag}"], **tol)
        assert norm["grey_matter"].dtype == grey.dtype


def test_oracle_wta_matches_reference():
    """create_wta_map restatement against outputs of the reference's own function (make_golden_wta.py)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_wta.npz"))
    for tag in ("f32", "f64"):
        for key in ("0p0", "1p0", "2p5"):
            z = float(key.replace("p", "."))
            white = orc.create_wta_map(g[f"white_{tag}"].copy(), 0, z)
            grey = orc.create_wta_map(g[f"grey_{tag}"].copy(), 1, z)
            assert white.dtype == g[f"white_wta_{key}_{tag}"].dtype and white.shape == g[f"white_wta_{key}_{tag}"].shape
            assert np.array_equal(white, g[f"white_wta_{key}_{tag}"]), (tag, key)
            assert np.array_equal(grey, g[f"grey_wta_{key}_{tag}"]), (tag, key)


def test_oracle_wta_against_sklearn_formula():
    """The restatement against the reference's own three lines (StandardScaler + max / argmax, image_handling.py
    :240-245) evaluated with the installed sklearn, on shapes and value patterns the golden file does not hold."""
    from sklearn.preprocessing import StandardScaler

    def reference_formula(component, axis, z_thr):
        z = StandardScaler().fit_transform(component)
        top = np.max(z, axis=axis, keepdims=True)
        wta = np.argmax(z, axis=axis, keepdims=True) + 1
        wta[top < z_thr] = 0.0
        return np.array(wta, dtype=int)

    rng = np.random.default_rng(123)
    for (m, k) in ((1, 1), (1, 7), (5, 1), (257, 13), (77, 203)):
        for dtype in (np.float32, np.float64):
            grey = (rng.gamma(0.5, 1.0, (m, k)) * (rng.random((m, k)) < 0.7)).astype(dtype)
            grey[:, 0] = 1.25                                     # a constant column
            white = np.ascontiguousarray(grey.T)
            for z in (-1.0, 0.0, 0.5, 2.0):
                for arr, axis in ((grey, 1), (white, 0), (grey, 0), (white, 1)):
                    assert np.array_equal(orc.create_wta_map(arr, axis, z), reference_formula(arr.copy(), axis, z)), \
                        (m, k, dtype, z, axis)
    counts = rng.integers(0, 40, (300, 9))
    assert np.array_equal(orc.create_wta_map(counts, 1, 1.0), reference_formula(counts, 1, 1.0))
