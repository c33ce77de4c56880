"""GPU: size-independent properties at BASELINE.json's full single-GPU shape (C3: 59 412 x 110 000, k = 200),
where the oracle cannot run.  Inputs are generated on the device exactly like bench.py does.

  * adjoint identity  <W, X H'> == <H, W'X>: the K-major and the MN-major read of the same fp16 hi/lo planes
    (the two X-streaming GEMMs of an iteration) compute the same trace(W' X H') -- to fp32 accuracy;
  * both against a float64 evaluation of the same trace on a row sample (anchors the value, not only the symmetry);
  * coordinate descent is monotone: the penalised objective 0.5 ||X - WH||^2 + l1_W |W|_1 + l1_H |H|_1 never
    increases from one iteration to the next (sklearn's solver has the same property);
  * factors stay non-negative and finite; the multiplicative update is monotone in ||X - WH||_F;
  * the dual regression of the full-size subject satisfies the NNLS optimality (KKT) conditions of both stages,
    evaluated in float64 on sampled target columns / seed rows, for group components whose norms spread over four
    orders of magnitude (the solver works in diagonally scaled variables).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M, N, K = 59412, 110000, 200


@pytest.fixture(scope="module")
def resident():
    import torch
    from bench import random_init, synth_rows_device
    from nfact_b200.device import DeviceMatrix, get_handle
    dev = torch.device("cuda", 0)
    x = synth_rows_device(torch, dev, 0, M, N, K, 0.05)
    torch.cuda.synchronize()
    handle = get_handle()
    sample_rows = torch.arange(0, M, 997, device=dev)
    x_sample = x[sample_rows].double().cpu().numpy()
    sample_cols = torch.arange(0, N, 1999, device=dev)
    x_cols = x[:, sample_cols].double().cpu().numpy()
    x_mean = float(x.sum(dtype=torch.float64)) / (M * N)
    xmat = DeviceMatrix.from_device_ptr(x.data_ptr(), M, N, N, handle=handle)
    del x
    torch.cuda.empty_cache()
    w0, h0 = random_init(torch, dev, x_mean, M, N, K, 0)
    yield {"xmat": xmat, "w0": w0, "h0": h0, "rows": sample_rows.cpu().numpy(), "x_sample": x_sample, "torch": torch,
           "cols": sample_cols.cpu().numpy(), "x_cols": x_cols}
    xmat.free()


def test_adjoint_identity_of_the_two_gemms(resident):
    torch = resident["torch"]
    xmat, w0, h0 = resident["xmat"], resident["w0"], resident["h0"]
    dev = w0.device
    wt = w0.t().contiguous()                                  # [k, m] component-major
    xht = torch.empty((K, M), dtype=torch.float32, device=dev)
    wtx = torch.empty((K, N), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    xmat.matmul_dev(False, h0.data_ptr(), K, xht.data_ptr(), M)      # (X H')' : K-major read of X
    xmat.matmul_dev(True, wt.data_ptr(), K, wtx.data_ptr(), N)       # W'X     : MN-major read of the same planes
    xmat.handle.synchronize()
    lhs = float((xht.double() * wt.double()).sum())
    rhs = float((wtx.double() * h0.double()).sum())
    assert abs(lhs - rhs) <= 2e-6 * abs(lhs), (lhs, rhs)
    # anchor: float64 evaluation on every 997th seed row
    rows = resident["rows"]
    ref = resident["x_sample"] @ h0.double().cpu().numpy().T          # [rows, k]
    got = xht[:, torch.as_tensor(rows, device=dev)].t().double().cpu().numpy()
    assert np.linalg.norm(got - ref) <= 2e-6 * np.linalg.norm(ref)


@pytest.mark.parametrize("solver", ["cd", "mu"])
def test_objective_is_monotone_at_full_size(resident, solver):
    from nfact_b200.device import NmfState, make_params
    xmat, w0, h0 = resident["xmat"], resident["w0"], resident["h0"]
    l1_w, l1_h = 0.1 * N, 0.1 * M                 # NFACT defaults: alpha_W = 0.1, l1_ratio = 1
    regs = (l1_w, l1_h, 0.0, 0.0) if solver == "cd" else (0.0, 0.0, 0.0, 0.0)
    st = NmfState(xmat, K, make_params(solver, 10 ** 6, 0.0, regs))
    st.set_factors_dev(w0.data_ptr(), h0.data_ptr())

    def objective():
        err = st.error()
        w, h = st.get_factors()
        assert np.isfinite(w).all() and np.isfinite(h).all() and w.min() >= 0 and h.min() >= 0
        return 0.5 * err * err + regs[0] * float(w.sum(dtype=np.float64)) + regs[1] * float(h.sum(dtype=np.float64))

    values = [objective()]
    for _ in range(3):
        st.iterate(1)
        values.append(objective())
    st.free()
    for before, after in zip(values, values[1:]):
        assert after <= before * (1 + 1e-6), values
    assert values[-1] < values[0]


def test_dual_regression_kkt_at_full_size(resident):
    """nmf_dual_regression at C3: h >= 0, gradient >= 0 off the support and == 0 on it (scipy's nnls returns the
    point with exactly these properties), both stages, on every 1999th target column and every 997th seed row."""
    import nfact_b200
    xmat = resident["xmat"]
    rng = np.random.default_rng(4)
    G = rng.gamma(0.3, 1.0, (M, K)) * (rng.random((M, K)) < 0.3)
    G *= 10.0 ** rng.uniform(-2, 2, K)                       # component norms over four orders of magnitude
    G[:, 7] = 0.0                                            # a dead component
    dr = nfact_b200.nmf_dual_regression({"grey_components": G}, xmat)
    Hs, Gs = dr["white_components"], dr["grey_components"]
    assert Hs.shape == (K, N) and Gs.shape == (M, K) and Hs.dtype == np.float64 and Gs.dtype == np.float64
    assert np.isfinite(Hs).all() and np.isfinite(Gs).all() and Hs.min() >= 0 and Gs.min() >= 0
    assert not Hs[7].any()
    # stage 1: min_h ||G h - x_j||, gradient G'(G h - x_j)
    cols, xc = resident["cols"], resident["x_cols"]          # [M, c]
    hc = Hs[:, cols]
    grad1 = G.T @ (G @ hc - xc)
    scale1 = np.abs(G.T @ xc).max()
    assert grad1.min() > -1e-5 * scale1, grad1.min() / scale1
    assert np.abs(grad1[hc > 0]).max() < 1e-5 * scale1, np.abs(grad1[hc > 0]).max() / scale1
    # stage 2: min_g ||Hs' g - x_i||, gradient (g Hs - x_i) Hs'
    rows, xr = resident["rows"], resident["x_sample"]        # [r, N]
    gr = Gs[rows]
    grad2 = (gr @ Hs - xr) @ Hs.T
    scale2 = np.abs(xr @ Hs.T).max()
    assert grad2.min() > -1e-5 * scale2, grad2.min() / scale2
    assert np.abs(grad2[gr > 0]).max() < 1e-5 * scale2, np.abs(grad2[gr > 0]).max() / scale2
