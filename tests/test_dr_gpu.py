"""GPU: non-negative dual regression against scipy.optimize.nnls (golden) and the oracle."""
import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu

DR_TOL = 1e-4     # rel-Frobenius <= 1e-4 vs scipy float64


def test_dr_matches_scipy_golden(golden):
    import nfact_b200
    comps = {"grey_components": golden["cd_fit_W"].astype(np.float64),
             "white_components": golden["cd_fit_H"].astype(np.float64)}
    dr = nfact_b200.nmf_dual_regression(comps, golden["ingest_single_1"], 1)
    assert dr["grey_components"].dtype == np.float64 and dr["grey_components"].shape == golden["dr_grey"].shape
    assert dr["white_components"].shape == golden["dr_white"].shape
    assert rel_fro(dr["white_components"], golden["dr_white"]) < DR_TOL
    assert rel_fro(dr["grey_components"], golden["dr_grey"]) < DR_TOL


def test_dr_second_subject_parallel_flag(golden):
    import nfact_b200
    comps = {"grey_components": golden["mu_fit_W"], "white_components": golden["mu_fit_H"]}
    dr = nfact_b200.run_decomp(nfact_b200.nmf_dual_regression, comps, golden["ingest_single_2"], 4)
    assert rel_fro(dr["white_components"], golden["dr2_white"]) < DR_TOL
    assert rel_fro(dr["grey_components"], golden["dr2_grey"]) < DR_TOL


def test_dr_kkt_conditions():
    """Size-independent property: the solution satisfies the NNLS KKT conditions for both stages."""
    import nfact_b200
    rng = np.random.default_rng(9)
    m, n, k = 700, 1100, 37
    G = rng.gamma(0.3, 1.0, (m, k))
    X = (rng.gamma(0.3, 1.0, (m, k)) @ rng.gamma(0.3, 1.0, (k, n)) * (rng.random((m, n)) < 0.3)).astype(np.float32)
    dr = nfact_b200.nmf_dual_regression({"grey_components": G}, X)
    Hs, Gs = dr["white_components"], dr["grey_components"]
    assert (Hs >= 0).all() and (Gs >= 0).all()
    X64 = X.astype(np.float64)
    grad1 = G.T @ (G @ Hs - X64)                      # [k, n]
    scale1 = np.abs(G.T @ X64).max()
    assert grad1.min() > -1e-5 * scale1
    assert np.abs(grad1[Hs > 0]).max() < 1e-5 * scale1
    grad2 = (Gs @ Hs - X64) @ Hs.T                    # [m, k]
    scale2 = np.abs(X64 @ Hs.T).max()
    assert grad2.min() > -1e-5 * scale2
    assert np.abs(grad2[Gs > 0]).max() < 1e-5 * scale2


def test_dr_shape_mismatch_exits(golden):
    import nfact_b200
    comps = {"grey_components": np.ones((5, 3))}
    with pytest.raises(SystemExit):
        nfact_b200.run_decomp(nfact_b200.nmf_dual_regression, comps, golden["ingest_single_1"], 1)


def test_dr_subject_pipeline(golden, tmp_path):
    """dual_regression_pipeline compute part: ingest -> DR -> per-seed threshold -> normalise."""
    import os
    import shutil
    import nfact_b200
    from conftest import GOLDEN_DIR
    sub = tmp_path / "sub-1"
    shutil.copytree(os.path.join(GOLDEN_DIR, "data", "sub-1"), sub)
    coords = np.zeros((99, 5), dtype=int)
    coords[50:, -2] = 1                                   # two seeds: rows 0-49 and 50-98
    np.savetxt(sub / "coords_for_fdt_matrix2", coords, fmt="%d")
    comps = {"grey_components": golden["cd_fit_W"].astype(np.float64)}
    out = nfact_b200.run_locally([str(sub)], comps, ["L", "R"], threshold=1, normalise=True)
    res = out[str(sub)]
    assert res["grey_components"].shape == (99, 10) and res["white_components"].shape == (10, 199)
    assert res["normalised_white"].shape == (10, 199)
    ref_white = nfact_b200.thresholding(golden["dr_white"].copy(), 1)
    assert np.abs(res["white_components"] - ref_white).max() <= 1e-3 * np.abs(ref_white).max()


def test_dr_float32_components_and_nonfinite(golden):
    """GIFTI / CIFTI loaders hand float32 group components (nfact_dr_functions.py:296-353): same result as the
    float64 copy of the same values; inf / NaN anywhere is rejected like scipy does (checked on the device)."""
    import nfact_b200
    g32 = golden["cd_fit_W"].astype(np.float32)
    dr32 = nfact_b200.nmf_dual_regression({"grey_components": g32}, golden["ingest_single_1"])
    dr64 = nfact_b200.nmf_dual_regression({"grey_components": g32.astype(np.float64)}, golden["ingest_single_1"])
    assert dr32["grey_components"].dtype == np.float64
    assert np.array_equal(dr32["white_components"], dr64["white_components"])
    assert np.array_equal(dr32["grey_components"], dr64["grey_components"])
    bad = g32.copy()
    bad[7, 3] = np.nan
    with pytest.raises(ValueError):
        nfact_b200.nmf_dual_regression({"grey_components": bad}, golden["ingest_single_1"])
    xbad = golden["ingest_single_1"].copy()
    xbad[5, 60] = np.inf
    with pytest.raises(ValueError):
        nfact_b200.nmf_dual_regression({"grey_components": g32}, xbad)


def test_run_locally_prefetch_equals_sequential(golden, tmp_path):
    """The prefetching subject loop (next subject ingested on a second stream while the current one is solved)
    returns exactly what the sequential loop returns, subject by subject."""
    import os
    import shutil
    import nfact_b200
    from conftest import GOLDEN_DIR
    dirs = []
    for s in (1, 2, 3):
        sub = tmp_path / f"sub-{s}"
        shutil.copytree(os.path.join(GOLDEN_DIR, "data", f"sub-{s}"), sub)
        coords = np.zeros((99, 5), dtype=int)
        coords[40:, -2] = 1
        np.savetxt(sub / "coords_for_fdt_matrix2", coords, fmt="%d")
        dirs.append(str(sub))
    comps = {"grey_components": golden["cd_fit_W"].astype(np.float64)}
    seq = nfact_b200.run_locally(dirs, comps, ["L", "R"], threshold=2, prefetch=False)
    pre = nfact_b200.run_locally(dirs, comps, ["L", "R"], threshold=2, prefetch=True)
    assert list(seq) == list(pre) == dirs
    for d in dirs:
        for key in ("grey_components", "white_components"):
            assert np.array_equal(seq[d][key], pre[d][key])
    # rank r of `world` takes subjects r, r + world, ...
    part = nfact_b200.run_locally(dirs, comps, ["L", "R"], threshold=2, rank=1, world=2)
    assert list(part) == [dirs[1]] and np.array_equal(part[dirs[1]]["white_components"], seq[dirs[1]]["white_components"])


def test_prepared_components_equal_host_components(golden, tmp_path):
    """DeviceComponents (group maps uploaded once for the whole subject loop, nfb_dr_components_create +
    nfb_dual_regression_resident_host) give bit-identical results to passing the host array every time;
    float32 maps, non-finite maps and a shape mismatch behave as on the host path."""
    import os
    import shutil
    import nfact_b200
    from conftest import GOLDEN_DIR
    X = golden["ingest_single_1"]
    for dtype in (np.float64, np.float32):
        G = golden["cd_fit_W"].astype(dtype)
        host = nfact_b200.nmf_dual_regression({"grey_components": G}, X)
        prep = nfact_b200.DeviceComponents(G)
        assert prep.shape == G.shape and not prep.has_nonfinite
        for _ in range(2):                                    # re-used across subjects
            dev = nfact_b200.nmf_dual_regression({"grey_components": prep}, X, threshold=0)
            assert np.array_equal(dev["grey_components"], host["grey_components"])
            assert np.array_equal(dev["white_components"], host["white_components"])
        with pytest.raises(ValueError):
            nfact_b200.nmf_dual_regression({"grey_components": prep}, X[:50])
        prep.free()
        with pytest.raises(ValueError):
            nfact_b200.nmf_dual_regression({"grey_components": prep}, X)
    bad = golden["cd_fit_W"].astype(np.float64)
    bad[3, 2] = np.inf
    with pytest.raises(ValueError):
        nfact_b200.nmf_dual_regression({"grey_components": nfact_b200.DeviceComponents(bad)}, X)
    # run_locally with a save callback hands every result to it and keeps none
    sub = tmp_path / "sub-1"
    shutil.copytree(os.path.join(GOLDEN_DIR, "data", "sub-1"), sub)
    np.savetxt(sub / "coords_for_fdt_matrix2", np.zeros((99, 5), dtype=int), fmt="%d")
    saved = {}
    out = nfact_b200.run_locally([str(sub)], {"grey_components": golden["cd_fit_W"].astype(np.float64)}, ["L"],
                                 threshold=0, save=lambda d, r: saved.update({d: r["white_components"].copy()}))
    assert out == {str(sub): None} and saved[str(sub)].shape == (10, 199)
