"""GPU: the split-precision tcgen05 GEMM against a float64 reference (through the C ABI)."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run_gemm(a, b, transposed, cta_group, splits=0):
    """a: numpy float32 [rows, cols] as stored; returns out [n_b, M] float32."""
    import torch
    import nfact_b200
    from nfact_b200._lib import check
    h = nfact_b200.get_handle()
    dev = torch.device("cuda", 0)
    ta = torch.from_numpy(a).to(dev)
    tb = torch.from_numpy(b).to(dev)
    m_total = a.shape[1] if transposed else a.shape[0]
    ld_out = (m_total + 7) // 8 * 8
    out = torch.zeros((b.shape[0], ld_out), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    check(h.lib.nfb_test_gemm(h.ptr, ctypes.c_void_p(ta.data_ptr()), a.shape[0], a.shape[1], int(transposed),
                              ctypes.c_void_p(tb.data_ptr()), b.shape[0], ctypes.c_void_p(out.data_ptr()), ld_out,
                              cta_group, splits))
    torch.cuda.synchronize()
    return out[:, :m_total].cpu().numpy()


def _reference(a, b, transposed):
    a64 = a.astype(np.float64)
    op = a64.T if transposed else a64
    return b.astype(np.float64) @ op.T        # [n_b, M]


CASES = [
    # (rows, cols, n_b, transposed)
    (128, 64, 16, False),
    (200, 333, 50, False),
    (300, 517, 10, False),
    (1000, 2100, 200, False),
    (64, 128, 16, True),
    (333, 200, 50, True),
    (517, 300, 10, True),
    (2100, 1000, 200, True),
    (4100, 3000, 100, False),
    (3000, 4100, 100, True),
]


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("rows,cols,n_b,transposed", CASES)
def test_gemm_matches_float64(rows, cols, n_b, transposed, cta_group):
    rng = np.random.default_rng(rows * 7 + cols + n_b)
    a = (rng.gamma(0.3, 1.0, size=(rows, cols)) * 300 * (rng.random((rows, cols)) < 0.3)).astype(np.float32)
    k_total = rows if transposed else cols
    b = rng.gamma(0.5, 2.0, size=(n_b, k_total)).astype(np.float32)
    b[n_b // 2] *= 1e-6          # a nearly dead component keeps its relative accuracy (per-row scales)
    out = _run_gemm(a, b, transposed, cta_group)
    ref = _reference(a, b, transposed)
    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-300
    err = np.abs(out - ref) / scale
    assert err.max() < 1e-5, f"max rel err {err.max():.3e}"
    rel_fro = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert rel_fro < 5e-6, f"rel fro {rel_fro:.3e}"


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_signed_values_and_splits(cta_group):
    rng = np.random.default_rng(5)
    a = rng.standard_normal((700, 1500)).astype(np.float32)
    b = rng.standard_normal((48, 1500)).astype(np.float32)
    ref = _reference(a, b, False)
    for splits in (1, 3, 8):
        out = _run_gemm(a, b, False, cta_group, splits)
        assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-5, splits


STREAMK_CASES = [
    # (rows, cols, n_b, transposed): M blocks against 74 CTA pairs / 148 CTAs -- fewer tiles than workers (every tile
    # shared by several workers), one tile more than a wave, a ragged last tile, a K that is not a multiple of 64
    (1300, 9000, 50, False),      # 6 / 11 tiles, 141 K blocks
    (5000, 10000, 50, False),     # C1's X H': 20 / 40 tiles on 74 / 148 workers
    (9000, 2100, 200, True),      # M = 2100 (9 / 17 tiles), K = 9000
    (2050, 19300, 64, True),      # M = 19300: 76 / 151 tiles = one wave + a few tiles, K = 2050 (33 K blocks)
    (20000, 2050, 208, False),    # M = 20000 (79 / 157 tiles), K = 2050
]


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("rows,cols,n_b,transposed", STREAMK_CASES)
def test_gemm_streamk_tail_matches_float64(rows, cols, n_b, transposed, cta_group):
    """Unsplit launches with the stream-K tail (splits = -1; launched twice on one workspace): the tiles of the last,
    partial wave are shared between the workers and fixed up inside the kernel -- same accuracy as whole tiles, and
    identical results from run to run (fixed summation order)."""
    rng = np.random.default_rng(rows + 3 * cols + n_b)
    a = (rng.gamma(0.3, 1.0, size=(rows, cols)) * 300 * (rng.random((rows, cols)) < 0.3)).astype(np.float32)
    k_total = rows if transposed else cols
    b = rng.gamma(0.5, 2.0, size=(n_b, k_total)).astype(np.float32)
    out = _run_gemm(a, b, transposed, cta_group, splits=-1)
    ref = _reference(a, b, transposed)
    scale = np.abs(ref).max(axis=1, keepdims=True) + 1e-300
    assert (np.abs(out - ref) / scale).max() < 1e-5
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 5e-6
    again = _run_gemm(a, b, transposed, cta_group, splits=-1)
    assert np.array_equal(out, again)
