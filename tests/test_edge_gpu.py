"""GPU: edge cases of the hot path -- tiny / ragged shapes, k = 1, zero rows and columns, all-zero X,
single iterations, shapes straddling the 128/256-row and 64-column tile boundaries."""
import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu


def _run(X, k, solver, iters, regs=None, seed=0):
    from nfact_b200.device import DeviceMatrix, NmfState, make_params
    from oracle import nfact_oracle as orc
    m, n = X.shape
    W0, H0 = orc.init_random(X, k, seed) if X.any() else (np.ones((m, k), np.float32), np.ones((k, n), np.float32))
    regs = regs or orc.compute_regularization(m, n, 0.01, "same", 1.0)
    xdev = DeviceMatrix.from_host(X)
    st = NmfState(xdev, k, make_params(solver, 200, 1e-4, regs))
    st.set_factors(W0, H0)
    st.iterate(iters)
    W, H = st.get_factors()
    rec = []
    (orc.fit_mu if solver == "mu" else orc.fit_cd)(X, W0, H0, *regs, tol=0, max_iter=iters, record=rec)
    return W, H, rec[-1][1], rec[-1][2]


@pytest.mark.parametrize("solver", ["mu", "cd"])
@pytest.mark.parametrize("m,n,k", [(3, 5, 2), (1, 9, 1), (7, 1, 1), (129, 65, 3), (257, 130, 17), (64, 256, 16),
                                   (255, 63, 31), (130, 700, 33)])
def test_small_and_boundary_shapes(m, n, k, solver):
    rng = np.random.default_rng(m * 31 + n)
    X = (rng.gamma(0.5, 1.0, (m, n)) * (rng.random((m, n)) < 0.6) * 40).astype(np.float32)
    if not X.any():
        X[0, 0] = 1.0
    W, H, Wr, Hr = _run(X, k, solver, 3)
    assert W.shape == (m, k) and H.shape == (k, n)
    assert rel_fro(W, Wr) < 1e-4 and rel_fro(H, Hr) < 1e-4


@pytest.mark.parametrize("solver", ["mu", "cd"])
def test_zero_rows_and_columns(solver):
    rng = np.random.default_rng(2)
    X = (rng.gamma(0.5, 1.0, (90, 140)) * 30).astype(np.float32)
    X[10:25, :] = 0
    X[:, 100:] = 0            # like targets that are never hit (the reference fixtures have 100 such columns)
    W, H, Wr, Hr = _run(X, 6, solver, 4)
    assert np.isfinite(W).all() and np.isfinite(H).all()
    assert rel_fro(W, Wr) < 1e-4 and rel_fro(H, Hr) < 1e-4
    assert not H[:, 100:].any() or solver == "mu"      # CD drives the dead columns to exactly zero


def test_all_zero_matrix_stops_like_sklearn():
    """sklearn: violation_init == 0 -> break after the first iteration (_nmf.py:507-508)."""
    import nfact_b200
    X = np.zeros((20, 30), np.float32)
    est = nfact_b200.NMF(n_components=3, init="custom", solver="cd", alpha_W=0.0)
    W = est.fit_transform(X, W=np.ones((20, 3), np.float32), H=np.ones((3, 30), np.float32))
    assert est.n_iter_ <= 2 and np.isfinite(W).all() and est.reconstruction_err_ < 1e-3


def test_max_iter_one_and_float64_input(golden):
    import nfact_b200
    X64 = golden["ingest_group"].astype(np.float64)
    p = dict(nfact_b200.get_parameters(None, "nmf", 10), max_iter=1)
    res = nfact_b200.nmf_decomp(p, X64, W=golden["init_nndsvd_W"].copy(), H=golden["init_nndsvd_H"].copy())
    assert rel_fro(res["grey_components"], golden["cd_W_1"]) < 1e-4
    assert rel_fro(res["white_components"], golden["cd_H_1"]) < 1e-4


def test_negative_input_rejected(golden):
    import nfact_b200
    X = golden["ingest_group"].copy()
    X[3, 150] = -1.0
    with pytest.raises(SystemExit):
        nfact_b200.nmf_decomp(nfact_b200.get_parameters(None, "nmf", 4), X)


def test_dr_tiny_and_zero_component(golden):
    import nfact_b200
    from oracle import nfact_oracle as orc
    rng = np.random.default_rng(6)
    X = (rng.gamma(0.5, 1.0, (40, 33)) * 10).astype(np.float32)
    G = rng.gamma(0.5, 1.0, (40, 4))
    G[:, 2] = 0.0                                            # a dead group component
    dr = nfact_b200.nmf_dual_regression({"grey_components": G}, X)
    ref = orc.nmf_dual_regression({"grey_components": G}, X)
    assert rel_fro(dr["white_components"], ref["white_components"]) < 1e-4
    assert rel_fro(dr["grey_components"], ref["grey_components"]) < 1e-4
    assert not dr["white_components"][2].any()


@pytest.mark.parametrize("tile", [32, 64, 96, 128, 160, 192, 224, 256])
@pytest.mark.parametrize("k", [16, 37, 50])
def test_cd_sweep_every_tile_size(tile, k, monkeypatch):
    """Every instantiation of the blocked sweep (sample tile x thread count) against the oracle, with k a
    multiple of the 16-wide component block, one component past it, and a ragged last block; the column
    count leaves a partial tile and is not a multiple of 4 (scalar seed path)."""
    monkeypatch.setenv("NFB_CD_TILE", str(tile))
    rng = np.random.default_rng(tile + k)
    m, n = 611, 1030
    X = (rng.gamma(0.5, 1.0, (m, n)) * (rng.random((m, n)) < 0.5) * 40).astype(np.float32)
    W, H, Wr, Hr = _run(X, k, "cd", 3)
    assert rel_fro(W, Wr) < 1e-4 and rel_fro(H, Hr) < 1e-4


def test_results_are_page_locked_numpy_arrays(golden):
    """W / H come back as ordinary writable float32 arrays (page-locked underneath, recycled by size)."""
    import gc
    import nfact_b200
    p = dict(nfact_b200.get_parameters(None, "nmf", 10), max_iter=2)
    res = nfact_b200.nmf_decomp(p, golden["ingest_group"], W=golden["init_nndsvd_W"].copy(),
                                H=golden["init_nndsvd_H"].copy())
    W, H = res["grey_components"], res["white_components"]
    assert isinstance(W, np.ndarray) and W.dtype == np.float32 and W.flags.writeable and W.flags.c_contiguous
    addrs = [W.ctypes.data]
    nfact_b200.thresholding(H, 1)          # downstream code mutates the arrays in place
    keep = W.copy()
    del res, W, H
    gc.collect()
    # the pool hands out the most recently freed block of a size first
    for _ in range(8):
        res2 = nfact_b200.nmf_decomp(p, golden["ingest_group"], W=golden["init_nndsvd_W"].copy(),
                                     H=golden["init_nndsvd_H"].copy())
        assert np.array_equal(res2["grey_components"], keep)        # deterministic, and not clobbered
        addrs.append(res2["grey_components"].ctypes.data)
        del res2
        gc.collect()
        if len(set(addrs)) < len(addrs):
            break
    assert len(set(addrs)) < len(addrs)                             # a block was recycled
