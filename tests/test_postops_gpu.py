"""GPU: component post-processing on the device against outputs of the reference's own
thresholding / threshold_grey_components / thresholding_components / normalise_components
(tests/golden/make_golden_postops.py), standalone and fused into the dual regression."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR, rel_fro

pytestmark = pytest.mark.gpu

COORDS = os.path.join(GOLDEN_DIR, "data", "postops", "coords_for_fdt_matrix2")
SEEDS = ["left.surf.gii", "right.surf.gii"]


@pytest.fixture(scope="module")
def g():
    return dict(np.load(os.path.join(GOLDEN_DIR, "golden_postops.npz")))


def _same_zero_pattern(got, want, max_flips=0):
    """Thresholding only zeroes entries: survivors must be bit-identical, and the zero patterns equal (an entry
    sitting within rounding of its threshold could flip; none does on these vectors)."""
    flips = np.count_nonzero((got == 0) != (want == 0))
    assert flips <= max_flips, f"{flips} entries thresholded differently"
    keep = (got != 0) & (want != 0)
    assert np.array_equal(got[keep], want[keep])


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_thresholding_matches_reference(g, tag):
    import nfact_b200
    for z in (1, 3):
        white = g[f"white_{tag}"].copy()
        out = nfact_b200.thresholding(white, z)
        assert out is white and out.dtype == g[f"white_{tag}"].dtype          # in place, dtype kept
        _same_zero_pattern(white, g[f"white_thr{z}_{tag}"])
        grey = g[f"grey_{tag}"].copy()
        out = nfact_b200.threshold_grey_components(grey, COORDS, SEEDS, z)
        assert out is grey
        _same_zero_pattern(grey, g[f"grey_thr{z}_{tag}"])
        # rows that belong to no listed seed are untouched
        outside = g["seeds_id"] >= len(SEEDS)
        assert np.array_equal(grey[outside], g[f"grey_{tag}"][outside])
    comps = nfact_b200.thresholding_components(2, COORDS, SEEDS, {"grey_components": g[f"grey_{tag}"].copy(),
                                                                  "white_components": g[f"white_{tag}"].copy()})
    _same_zero_pattern(comps["grey_components"], g[f"grey_thr2_{tag}"])
    _same_zero_pattern(comps["white_components"], g[f"white_thr2_{tag}"])
    untouched = nfact_b200.thresholding_components(0, COORDS, SEEDS, {"grey_components": g[f"grey_{tag}"].copy(),
                                                                      "white_components": g[f"white_{tag}"].copy()})
    assert np.array_equal(untouched["grey_components"], g[f"grey_{tag}"])


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_normalise_matches_reference(g, tag, golden):
    import nfact_b200
    tol = dict(rtol=2e-6, atol=2e-6) if tag == "f32" else dict(rtol=1e-12, atol=1e-12)
    for src in ("", "_thr2"):
        grey, white = g[f"grey{src}_{tag}"], g[f"white{src}_{tag}"]
        norm = nfact_b200.normalise_components(grey.copy(), white.copy())
        assert norm["grey_matter"].dtype == grey.dtype and norm["grey_matter"].shape == grey.shape
        assert np.allclose(norm["grey_matter"], g[f"grey{src}_norm_{tag}"], **tol)
        assert np.allclose(norm["white_matter"], g[f"white{src}_norm_{tag}"], **tol)
    # the vectors the first golden set holds for the NMF output of the fixture subjects
    norm = nfact_b200.normalise_components(golden["cd_fit_W"].copy(), golden["cd_fit_H"].copy())
    assert np.allclose(norm["grey_matter"], golden["norm_grey"], rtol=1e-5, atol=1e-6)
    assert np.allclose(norm["white_matter"], golden["norm_white"], rtol=1e-5, atol=1e-6)
    _same_zero_pattern(nfact_b200.thresholding(golden["cd_fit_H"].copy(), 3), golden["thresh_white_3"])
    _same_zero_pattern(nfact_b200.thresholding(golden["mu_fit_H"].copy(), 1), golden["thresh_white_1"])


def test_views_other_dtypes_and_empty():
    import nfact_b200
    from oracle import nfact_oracle as orc
    rng = np.random.default_rng(3)
    base = rng.gamma(0.5, 1.0, (40, 70))
    view = base[::2, 5:60]                                   # non-contiguous view, thresholded in place
    want = orc.thresholding(view.copy(), 1)
    nfact_b200.thresholding(view, 1)
    _same_zero_pattern(view, want)
    assert np.array_equal(base[1::2], base[1::2]) and base[0, 0] != 0
    ints = (rng.integers(0, 50, (6, 90))).astype(np.int64)   # integer maps: statistics in float64
    want = orc.thresholding(ints.copy(), 1)
    _same_zero_pattern(nfact_b200.thresholding(ints, 1), want)
    empty = np.zeros((0, 10), np.float32)
    assert nfact_b200.thresholding(empty, 3).shape == (0, 10)
    one = np.full((1, 1), 2.0)
    assert nfact_b200.thresholding(one, 3)[0, 0] == 2.0      # x < mean + 3 * 0 is false


def test_large_maps_against_oracle():
    """Shapes that exercise the multi-CTA paths: ragged lengths, many groups."""
    import nfact_b200
    from oracle import nfact_oracle as orc
    rng = np.random.default_rng(8)
    k, m, n = 37, 4099, 10007
    white = (rng.gamma(0.3, 1.0, (k, n)) * (rng.random((k, n)) < 0.4))
    grey = (rng.gamma(0.3, 1.0, (m, k)) * (rng.random((m, k)) < 0.4)).astype(np.float32)
    seeds_id = rng.integers(0, 7, m)
    want_w = orc.thresholding(white.copy(), 2)
    want_g = orc.threshold_grey_components(grey.copy(), seeds_id, 5, 2)
    nfact_b200.thresholding(white, 2)
    from nfact_b200.matrix_handling import _component_postops
    _component_postops(grey, False, seeds_id, 5, threshold=True, zscore=2)
    _same_zero_pattern(white, want_w)
    _same_zero_pattern(grey, want_g, max_flips=2)            # float32 statistics in numpy vs float64 here
    norm = nfact_b200.normalise_components(grey, white)
    want = orc.normalise_components(grey, white)
    assert np.allclose(norm["grey_matter"], want["grey_matter"], rtol=1e-5, atol=1e-5)
    assert np.allclose(norm["white_matter"], want["white_matter"], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("normalise", [False, True])
def test_fused_dual_regression_postprocessing(tmp_path, normalise):
    """dual_regression_subject: load -> regression -> thresholding_components -> normalise_components, the last
    two on the device before the copy-out, equal to the oracle's pipeline on the regression result."""
    import shutil
    import nfact_b200
    from oracle import nfact_oracle as orc
    src = os.path.join(GOLDEN_DIR, "data", "sub-1")
    sub = tmp_path / "sub-1"
    shutil.copytree(src, sub)
    X = orc.load_fdt_matrix(str(sub / "fdt_matrix2.dot"))
    m, n = X.shape
    rng = np.random.default_rng(21)
    seeds_id = rng.integers(0, 2, m)
    np.savetxt(sub / "coords_for_fdt_matrix2",
               np.stack([np.zeros(m, int), np.zeros(m, int), np.zeros(m, int), seeds_id, np.arange(m)], axis=1), fmt="%d")
    G = rng.gamma(0.5, 1.0, (m, 6)) * (rng.random((m, 6)) < 0.7)
    ref = orc.nmf_dual_regression({"grey_components": G}, X)
    plain = nfact_b200.nmf_dual_regression({"grey_components": G}, X)
    assert rel_fro(plain["grey_components"], ref["grey_components"]) < 1e-4
    got = nfact_b200.dual_regression_subject(str(sub), {"grey_components": G}, SEEDS, threshold=2, normalise=normalise)
    # same thresholds as the un-fused path applied to the device's own regression result
    want = orc.thresholding_components(2, seeds_id, 2, {"grey_components": plain["grey_components"].copy(),
                                                        "white_components": plain["white_components"].copy()})
    _same_zero_pattern(got["grey_components"], want["grey_components"])
    _same_zero_pattern(got["white_components"], want["white_components"])
    assert ("normalised_grey" in got) == normalise
    if normalise:
        norm = orc.normalise_components(want["grey_components"], want["white_components"])
        assert np.allclose(got["normalised_grey"], norm["grey_matter"], rtol=1e-10, atol=1e-10)
        assert np.allclose(got["normalised_white"], norm["white_matter"], rtol=1e-10, atol=1e-10)
    no_thr = nfact_b200.dual_regression_subject(str(sub), {"grey_components": G}, SEEDS, threshold=0)
    assert np.array_equal(no_thr["grey_components"], plain["grey_components"])


def test_component_loadings_match_the_reference_formula(g):
    """__correlating (NFACT/stats/stats_component_loadings.py:199-231), restated here verbatim in numpy."""
    import nfact_b200

    def correlating(subject_data, group_data):
        group_mean, group_std = group_data.mean(axis=0), group_data.std(axis=0)
        subj_mean, subj_std = subject_data.mean(axis=0), subject_data.std(axis=0)
        cov = ((subject_data - subj_mean) * (group_data - group_mean)).sum(axis=0) / (subject_data.shape[0] - 1)
        return cov / (subj_std * group_std)

    rng = np.random.default_rng(12)
    group = g["grey_f64"].copy()
    group[:, 5] += rng.random(group.shape[0])                 # un-constant the constant component
    subject = group * rng.uniform(0.5, 1.5, group.shape) + 0.1 * rng.random(group.shape)
    with np.errstate(invalid="ignore", divide="ignore"):
        want = correlating(subject, group)
    got = nfact_b200.component_loadings(subject, group)
    assert got.shape == want.shape and got.dtype == np.float64
    assert np.isnan(got[3]) and np.isnan(want[3])             # the dead component: 0 / 0 like numpy
    ok = ~np.isnan(want)
    assert np.allclose(got[ok], want[ok], rtol=1e-12, atol=1e-14)
    got32 = nfact_b200.component_loadings(subject.astype(np.float32), group.astype(np.float32))
    assert np.allclose(got32[ok], want[ok], rtol=1e-5)
    big_s, big_g = rng.gamma(0.3, 1.0, (20011, 23)), rng.gamma(0.3, 1.0, (20011, 23))
    assert np.allclose(nfact_b200.component_loadings(big_s, big_g), correlating(big_s, big_g), rtol=1e-10, atol=1e-13)
    with pytest.raises(ValueError):
        nfact_b200.component_loadings(big_s, big_g[:, :5])


# ---- winner-takes-all maps (reference outputs from tests/golden/make_golden_wta.py) ------------------
@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_wta_maps_match_reference(tag):
    import nfact_b200
    w = np.load(os.path.join(GOLDEN_DIR, "golden_wta.npz"))
    for key in ("0p0", "1p0", "2p5"):
        z = float(key.replace("p", "."))
        white_in, grey_in = w[f"white_{tag}"].copy(), w[f"grey_{tag}"].copy()
        white = nfact_b200.create_wta_map(white_in, 0, z)
        grey = nfact_b200.create_wta_map(grey_in, 1, z)
        want_w, want_g = w[f"white_wta_{key}_{tag}"], w[f"grey_wta_{key}_{tag}"]
        assert white.shape == want_w.shape and white.dtype == want_w.dtype
        assert grey.shape == want_g.shape and grey.dtype == want_g.dtype
        assert np.array_equal(white, want_w), (tag, key, np.flatnonzero(white != want_w)[:8])
        assert np.array_equal(grey, want_g), (tag, key, np.flatnonzero(grey != want_g)[:8])
        assert np.array_equal(white_in, w[f"white_{tag}"]) and np.array_equal(grey_in, w[f"grey_{tag}"])  # inputs untouched


def test_wta_against_oracle_on_ragged_shapes():
    """Shapes that are no multiple of anything, k = 1, one sample, integer-valued and non-contiguous inputs."""
    import nfact_b200
    from oracle import nfact_oracle as orc
    rng = np.random.default_rng(99)
    for (m, k) in ((1, 1), (1, 7), (5, 1), (1031, 13), (77, 203), (4099, 50)):
        for dtype in (np.float32, np.float64):
            grey = (rng.gamma(0.5, 1.0, (m, k)) * (rng.random((m, k)) < 0.7)).astype(dtype)
            white = np.ascontiguousarray(grey.T)
            for z in (-1.0, 0.5, 2.0):
                assert np.array_equal(nfact_b200.create_wta_map(grey, 1, z), orc.create_wta_map(grey, 1, z)), (m, k, z)
                assert np.array_equal(nfact_b200.create_wta_map(white, 0, z), orc.create_wta_map(white, 0, z)), (m, k, z)
                # the scaler always works on the columns of what it is given: both axes on both layouts
                assert np.array_equal(nfact_b200.create_wta_map(grey, 0, z), orc.create_wta_map(grey, 0, z))
                assert np.array_equal(nfact_b200.create_wta_map(white, 1, z), orc.create_wta_map(white, 1, z))
    counts = rng.integers(0, 40, (300, 9))                                   # integer maps are scaled as float64
    assert np.array_equal(nfact_b200.create_wta_map(counts, 1, 1.0), orc.create_wta_map(counts, 1, 1.0))
    view = (rng.random((200, 24)) * 3).astype(np.float32)[:, ::2]            # a strided view
    assert np.array_equal(nfact_b200.create_wta_map(view, 1, 0.0), orc.create_wta_map(view, 1, 0.0))
    with pytest.raises(ValueError):
        nfact_b200.create_wta_map(np.zeros((0, 4), np.float32), 1, 1.0)
    with pytest.raises(ValueError):
        nfact_b200.create_wta_map(np.zeros(5, np.float32), 0, 1.0)
