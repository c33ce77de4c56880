"""GPU: NMF solver parity with sklearn as NFACT configures it (golden vectors + oracle)."""
import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu

MU_TOL = 1e-4     # BASELINE.json: MU trajectory within 1e-4 relative Frobenius error per iteration
CD_OBJ_TOL = 5e-3  # CD: final objective within 0.5 %
CD_CORR = 0.99     # CD: Hungarian-matched component correlation >= 0.99


def _state(X, k, solver, regs, max_iter=200, tol=1e-4):
    import nfact_b200
    from nfact_b200.device import DeviceMatrix, NmfState, make_params
    xdev = DeviceMatrix.from_host(X)
    return xdev, NmfState(xdev, k, make_params(solver, max_iter, tol, regs))


def matched_correlation(A, B):
    """Hungarian-matched Pearson correlation between the rows of A and B."""
    from scipy.optimize import linear_sum_assignment
    A = A - A.mean(axis=1, keepdims=True)
    B = B - B.mean(axis=1, keepdims=True)
    An = np.linalg.norm(A, axis=1, keepdims=True)
    Bn = np.linalg.norm(B, axis=1, keepdims=True)
    live = (An[:, 0] > 0) & (Bn[:, 0] > 0)
    C = (A / np.where(An > 0, An, 1)) @ (B / np.where(Bn > 0, Bn, 1)).T
    r, c = linear_sum_assignment(-C)
    return C[r, c], live


def test_mu_trajectory_matches_sklearn_per_iteration(golden):
    X = golden["ingest_group"]
    xdev, st = _state(X, 10, "mu", golden["regularization"])
    st.set_factors(golden["init_nndsvd_W"], golden["init_nndsvd_H"])
    assert abs(st.error() - golden["mu_err0"]) <= 1e-5 * golden["mu_err0"]
    for it in range(1, 41):
        st.iterate(1)
        if f"mu_W_{it}" in golden:
            W, H = st.get_factors()
            assert rel_fro(W, golden[f"mu_W_{it}"]) < MU_TOL, it
            assert rel_fro(H, golden[f"mu_H_{it}"]) < MU_TOL, it
            assert abs(st.error() - golden[f"mu_err_{it}"]) <= 1e-5 * golden[f"mu_err_{it}"], it


def test_mu_with_l1_and_l2(golden):
    X = golden["ingest_group"]
    xdev, st = _state(X, 10, "mu", golden["mu_l2_regs"])
    st.set_factors(golden["init_random_W"], golden["init_random_H"])
    st.iterate(5)
    W, H = st.get_factors()
    assert rel_fro(W, golden["mu_l2_W_5"]) < MU_TOL
    assert rel_fro(H, golden["mu_l2_H_5"]) < MU_TOL


def test_mu_full_fit_through_nmf_decomp(golden):
    import nfact_b200
    X = golden["ingest_group"]
    p = dict(nfact_b200.get_parameters(None, "nmf", 10), solver="mu")
    res = nfact_b200.nmf_decomp(p, X, W=golden["init_nndsvd_W"].copy(), H=golden["init_nndsvd_H"].copy())
    assert res["grey_components"].dtype == np.float32 and res["grey_components"].shape == (99, 10)
    assert res["white_components"].shape == (10, 199)
    assert rel_fro(res["grey_components"], golden["mu_fit_W"]) < 1e-3
    assert rel_fro(res["white_components"], golden["mu_fit_H"]) < 1e-3


def test_mu_stopping_rule(golden):
    import nfact_b200
    X = golden["ingest_group"]
    est = nfact_b200.NMF(**dict(nfact_b200.get_parameters(None, "nmf", 10), solver="mu", init="custom"))
    est.fit_transform(X, W=golden["init_nndsvd_W"].copy(), H=golden["init_nndsvd_H"].copy())
    assert est.n_iter_ == int(golden["mu_fit_n_iter"])


def test_cd_sweeps_match_sklearn(golden):
    X = golden["ingest_group"]
    xdev, st = _state(X, 10, "cd", golden["regularization"])
    st.set_factors(golden["init_nndsvd_W"], golden["init_nndsvd_H"])
    for it in range(1, 4):
        viol = st.iterate(1, want_violation=True)
        W, H = st.get_factors()
        assert rel_fro(W, golden[f"cd_W_{it}"]) < 1e-4, it
        assert rel_fro(H, golden[f"cd_H_{it}"]) < 1e-4, it
        assert abs(viol - golden["cd_violation"][it - 1]) <= 1e-3 * golden["cd_violation"][it - 1], it


def test_cd_full_fit_nfact_defaults(golden):
    import nfact_b200
    from oracle import nfact_oracle as orc
    X = golden["ingest_group"]
    p = nfact_b200.get_parameters(None, "nmf", 10)
    est = nfact_b200.NMF(**dict(p, init="custom"))
    W = est.fit_transform(X, W=golden["init_nndsvd_W"].copy(), H=golden["init_nndsvd_H"].copy())
    H = est.components_
    regs = golden["regularization"]
    obj = orc.nmf_objective(X, W, H, *regs)
    ref = orc.nmf_objective(X, golden["cd_fit_W"], golden["cd_fit_H"], *regs)
    assert abs(obj - ref) <= CD_OBJ_TOL * ref
    corr, live = matched_correlation(H, golden["cd_fit_H"])
    assert corr[live].min() >= CD_CORR
    corr_w, live_w = matched_correlation(W.T, golden["cd_fit_W"].T)
    assert corr_w[live_w].min() >= CD_CORR
    assert abs(est.n_iter_ - int(golden["cd_fit_n_iter"])) <= 5
    assert abs(est.reconstruction_err_ - golden["cd_fit_err"]) <= 1e-3 * golden["cd_fit_err"]


def test_cd_second_case_l2(golden):
    import nfact_b200
    from oracle import nfact_oracle as orc
    X = golden["ingest_group"]
    p = dict(nfact_b200.get_parameters(None, "nmf", 10), alpha_W=0.001, l1_ratio=0.5, max_iter=60)
    res = nfact_b200.nmf_decomp(p, X, W=golden["init_random_W"].copy(), H=golden["init_random_H"].copy())
    regs = orc.compute_regularization(99, 199, 0.001, "same", 0.5)
    obj = orc.nmf_objective(X, res["grey_components"], res["white_components"], *regs)
    ref = orc.nmf_objective(X, golden["cd2_fit_W"], golden["cd2_fit_H"], *regs)
    assert abs(obj - ref) <= CD_OBJ_TOL * ref
    corr, live = matched_correlation(res["white_components"], golden["cd2_fit_H"])
    assert corr[live].min() >= CD_CORR


@pytest.mark.parametrize("solver", ["mu", "cd"])
def test_medium_random_against_oracle(solver):
    """A larger ragged case (m, n not multiples of the tile sizes, k not a multiple of 16)."""
    from oracle import nfact_oracle as orc
    rng = np.random.default_rng(3)
    m, n, k = 517, 1203, 23
    Wt, Ht = rng.gamma(0.3, 1.0, (m, k)), rng.gamma(0.3, 1.0, (k, n))
    X = (np.floor(300 * (Wt @ Ht)) * (rng.random((m, n)) < 0.4) * (1e8 / 1.5e7)).astype(np.float32)
    W0, H0 = orc.init_random(X, k, 1)
    regs = orc.compute_regularization(m, n, 0.01, "same", 1.0)
    xdev, st = _state(X, k, solver, regs)
    st.set_factors(W0, H0)
    rec = []
    (orc.fit_mu if solver == "mu" else orc.fit_cd)(X, W0, H0, *regs, tol=0, max_iter=6, record=rec)
    for it in range(1, 7):
        st.iterate(1)
        W, H = st.get_factors()
        assert rel_fro(W, rec[it - 1][1]) < MU_TOL, (solver, it)
        assert rel_fro(H, rec[it - 1][2]) < MU_TOL, (solver, it)


def test_inits_and_error_paths(golden):
    import nfact_b200
    X = golden["ingest_group"]
    p = nfact_b200.get_parameters(None, "nmf", 10)
    res = nfact_b200.nmf_decomp(dict(p, init="random", random_state=7, max_iter=3), X)
    assert res["grey_components"].shape == (99, 10) and (res["grey_components"] >= 0).all()
    res = nfact_b200.nmf_decomp(dict(p, random_state=0, max_iter=5), X)          # nndsvd on the device
    assert np.isfinite(res["white_components"]).all()
    with pytest.raises(SystemExit):                                              # nmfsso.py:57-58
        nfact_b200.nmf_decomp(dict(p, beta_loss="kullback-leibler"), X)
    with pytest.raises(SystemExit):
        nfact_b200.nmf_decomp(p, X, W=np.ones((5, 10), np.float32), H=np.ones((10, 199), np.float32))


def test_nndsvd_device_matches_sklearn(golden):
    import nfact_b200
    from nfact_b200.device import DeviceMatrix
    from nfact_b200.nndsvd import nndsvd_device
    X = golden["ingest_group"]
    xdev = DeviceMatrix.from_host(X)
    W, H = nndsvd_device(xdev, 10, np.random.RandomState(0))
    assert rel_fro(W, golden["init_nndsvd_W"]) < 1e-3
    assert rel_fro(H, golden["init_nndsvd_H"]) < 1e-3


def test_nmf_sso_runs_on_resident_matrix():
    """NMF-sso orchestration: N random solves share one resident X, clustering, final custom solve."""
    import nfact_b200
    rng = np.random.default_rng(12)
    k = 4
    Wp = rng.gamma(0.4, 1.0, (120, k)) * (rng.random((120, k)) < 0.4)
    Hp = rng.gamma(0.4, 1.0, (k, 210)) * (rng.random((k, 210)) < 0.4)
    X = (Wp @ Hp * 100 + 0.01 * rng.random((120, 210))).astype(np.float32)
    p = dict(nfact_b200.get_parameters(None, "nmf", k), alpha_W=1e-4, max_iter=150)
    res = nfact_b200.nmf_sso(X, p, {"no_sso": False, "iterations": 6, "n_cores": 1, "dim": k, "outdir": "."})
    kf = res["white_components"].shape[0]
    assert 2 <= kf <= k and res["grey_components"].shape == (120, kf)
    assert np.isfinite(res["white_components"]).all() and (res["grey_components"] >= 0).all()
    rel = np.linalg.norm(X - res["grey_components"] @ res["white_components"]) / np.linalg.norm(X)
    assert rel < 0.35
    # nmf_sso mutates its parameter dict like the reference does (init -> custom): start from fresh ones
    fresh = dict(nfact_b200.get_parameters(None, "nmf", k), alpha_W=1e-4, max_iter=50, random_state=0)
    single = nfact_b200.nmf_sso(X, fresh, {"no_sso": True})
    assert single["white_components"].shape == (k, 210)


def test_nmf_sso_resident_flow_matches_host_stack_flow(monkeypatch):
    """nmf_sso with the runs' components resident (device random starts, stacks, similarity from the
    stack, centrotypes gathered into the final solve) against the host-stack flow (NMFsso: every run's W, H copied
    out, stacked and re-uploaded) on a planted study whose runs agree: same number of components, matched
    components correlate, same quality of fit.  The starts differ (device generator vs numpy), the fixed point of
    a well-separated study does not."""
    import nfact_b200
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(5)
    k, m, n = 4, 150, 260
    Wp = np.zeros((m, k)); Hp = np.zeros((k, n))
    for t in range(k):                                   # disjoint supports: the runs agree up to permutation
        r0, r1, c0, c1 = t * m // k, (t + 1) * m // k, t * n // k, (t + 1) * n // k
        Wp[r0:r1, t] = rng.gamma(2.0, 1.0, r1 - r0)
        Hp[t, c0:c1] = rng.gamma(2.0, 1.0, c1 - c0)
    X = (Wp @ Hp * 50 + 0.01 * rng.random((m, n))).astype(np.float32)
    args = {"no_sso": False, "iterations": 5, "n_cores": 1, "dim": k}
    monkeypatch.setenv("NFB_SSO_SEED", "11")
    res_dev = nfact_b200.nmf_sso(X, dict(nfact_b200.get_parameters(None, "nmf", k), alpha_W=1e-4, max_iter=200), dict(args))
    monkeypatch.setenv("NFB_SSO_RESIDENT", "0")
    res_host = nfact_b200.nmf_sso(X, dict(nfact_b200.get_parameters(None, "nmf", k), alpha_W=1e-4, max_iter=200), dict(args))
    assert res_dev["white_components"].shape == res_host["white_components"].shape == (k, n)
    assert res_dev["grey_components"].shape == (m, k)
    corr = np.corrcoef(res_dev["white_components"], res_host["white_components"])[:k, k:]
    rows, cols = linear_sum_assignment(-corr)
    assert corr[rows, cols].min() > 0.99
    for res in (res_dev, res_host):
        rel = np.linalg.norm(X - res["grey_components"] @ res["white_components"]) / np.linalg.norm(X)
        assert rel < 0.05


def test_device_random_start_has_sklearns_distribution():
    """nfb_nmf_init_random: |avg N(0, 1)| (sklearn _nmf.py:296-307) -- half-normal moments of both factors, different
    seeds give different draws, the same seed the same."""
    from nfact_b200.device import DeviceMatrix, NmfState, make_params
    rng = np.random.default_rng(0)
    m, n, k = 700, 1500, 12
    X = rng.random((m, n)).astype(np.float32)
    xdev = DeviceMatrix.from_host(X)
    st = NmfState(xdev, k, make_params("cd", 5, 0.0, (0.0, 0.0, 0.0, 0.0)))
    try:
        avg = 0.37
        st.init_random(avg, 123)
        W1, H1 = st.get_factors()
        st.init_random(avg, 123)
        W2, H2 = st.get_factors()
        st.init_random(avg, 124)
        W3, H3 = st.get_factors()
    finally:
        st.free(); xdev.free()
    assert np.array_equal(W1, W2) and np.array_equal(H1, H2)
    assert not np.array_equal(W1, W3) and not np.array_equal(H1, H3)
    for F in (W1, H1, H3):
        assert (F >= 0).all()
        z = F.astype(np.float64) / avg
        assert abs(z.mean() - np.sqrt(2 / np.pi)) < 0.02          # E|N| = sqrt(2/pi)
        assert abs((z ** 2).mean() - 1.0) < 0.03                 # E N^2 = 1
        assert abs((z > 1.0).mean() - 0.3173) < 0.02             # P(|N| > 1)
    assert abs(np.corrcoef(W1.ravel()[:8000], H1.ravel()[:8000])[0, 1]) < 0.05


@pytest.mark.parametrize("solver", ["mu", "cd"])
def test_more_than_256_components(solver):
    """k > 256 (configs[4] uses k = 500): the component dimension is tiled over several GEMM launches."""
    from oracle import nfact_oracle as orc
    rng = np.random.default_rng(8)
    m, n, k = 640, 901, 300
    X = (rng.gamma(0.3, 1.0, (m, 40)) @ rng.gamma(0.3, 1.0, (40, n)) * (rng.random((m, n)) < 0.5) * 50).astype(np.float32)
    W0, H0 = orc.init_random(X, k, 2)
    regs = orc.compute_regularization(m, n, 0.001, "same", 1.0)
    xdev, st = _state(X, k, solver, regs)
    st.set_factors(W0, H0)
    rec = []
    (orc.fit_mu if solver == "mu" else orc.fit_cd)(X, W0, H0, *regs, tol=0, max_iter=3, record=rec)
    st.iterate(3)
    W, H = st.get_factors()
    assert rel_fro(W, rec[-1][1]) < MU_TOL
    assert rel_fro(H, rec[-1][2]) < MU_TOL


def test_sso_similarity_matrix_matches_corrcoef():
    """compute_similairty_matrix (sso_functions.py:49-68) on the device vs np.corrcoef(rownorm(.)) in float64:
    ragged sizes (rows not a multiple of the 512-row factor tile is covered by r = 700), a dead component."""
    from nfact_b200.nmfsso import compute_similairty_matrix, rownorm
    rng = np.random.default_rng(12)
    for r, n in ((37, 501), (700, 3001)):
        comps = (rng.gamma(0.4, 1.0, (r, n)) * (rng.random((r, n)) < 0.3)).astype(np.float32)
        comps[5] = 0.0                                        # a component a run drove to all-zero
        sim = compute_similairty_matrix(comps)
        with np.errstate(invalid="ignore", divide="ignore"):
            ref = np.corrcoef(rownorm(comps.astype(np.float64)))
        ref[~np.isfinite(ref)] = 0.0
        np.fill_diagonal(ref, 1.0)
        np.clip(ref, 0, 1, out=ref)
        assert sim.shape == (r, r) and sim.dtype == np.float64
        assert np.abs(sim - ref).max() < 5e-6
        assert np.array_equal(sim, sim.T) and (np.diag(sim) == 1.0).all()


def test_cd_run_ahead_loop_equals_stepwise_loop(golden):
    """nfb_nmf_run keeps one CD iteration in flight while it looks at the previous violation and rolls the
    speculative iteration back when sklearn's rule says stop: same n_iter and bit-identical factors as
    stepping one iteration at a time with a synchronous check."""
    from nfact_b200.device import DeviceMatrix, NmfState, make_params
    from oracle import nfact_oracle as orc
    X = golden["ingest_group"]
    k = 10
    W0, H0 = golden["init_nndsvd_W"], golden["init_nndsvd_H"]
    regs = orc.compute_regularization(X.shape[0], X.shape[1], 0.1, "same", 1.0)
    xdev = DeviceMatrix.from_host(X)
    for tol, max_iter in ((1e-4, 200), (5e-2, 200), (1e-4, 7), (1e-4, 1), (0.9, 200)):
        fast = NmfState(xdev, k, make_params("cd", max_iter, tol, regs))
        fast.set_factors(W0.copy(), H0.copy())
        n_fast, err_fast = fast.run()
        Wf, Hf = fast.get_factors()
        fast.free()
        slow = NmfState(xdev, k, make_params("cd", max_iter, tol, regs))
        slow.set_factors(W0.copy(), H0.copy())
        n_slow, v_init = 0, None
        for it in range(1, max_iter + 1):
            v = slow.iterate(1, want_violation=True)
            n_slow = it
            v_init = v if it == 1 else v_init
            if v_init == 0 or v / v_init <= tol:
                break
        Ws, Hs = slow.get_factors()
        err_slow = slow.error()
        slow.free()
        assert n_fast == n_slow, (tol, max_iter, n_fast, n_slow)
        assert np.array_equal(Wf, Ws) and np.array_equal(Hf, Hs)
        assert abs(err_fast - err_slow) <= 1e-9 * err_slow
    xdev.free()


def test_sso_similarity_matches_reference_golden():
    """Device similarity matrix vs the reference's compute_similairty_matrix output (golden), and the clustering
    that follows from it: same partition and centrotypes as the reference run."""
    import os
    from conftest import GOLDEN_DIR
    from nfact_b200 import nmfsso
    g = np.load(os.path.join(GOLDEN_DIR, "golden_sso.npz"))
    sim = nmfsso.compute_similairty_matrix(g["components"])
    # float32 standardised rows + fp32 accumulation over n terms: a few 1e-6 absolute on correlations near 1
    assert np.abs(sim - g["sim"]).max() < 2e-5
    part = nmfsso.clustering_components(nmfsso.sim2dis(sim), int(g["dim"]))
    assert np.array_equal(part, g["partition"])
    assert np.array_equal(nmfsso.idx2centrotype(sim, part), g["centrotypes"])
