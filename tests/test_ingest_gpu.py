"""GPU: ingest / group average is bit-identical to the reference (golden) and to the oracle."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def test_single_subject_bit_exact(golden, golden_subjects):
    import nfact_b200
    for s, path in enumerate(golden_subjects):
        X = nfact_b200.load_fdt_matrix(path)
        assert X.dtype == np.float32 and X.flags.c_contiguous
        assert np.array_equal(X, golden[f"ingest_single_{s + 1}"])


def test_non_integer_values_bit_exact(golden):
    import nfact_b200
    X = nfact_b200.load_fdt_matrix(os.path.join(GOLDEN_DIR, "data", "sub-pd", "fdt_matrix2.dot"))
    assert np.array_equal(X, golden["ingest_single_pd"])


def test_group_average_bit_exact(golden, golden_subjects):
    import nfact_b200
    assert np.array_equal(nfact_b200.avg_fdt(golden_subjects), golden["ingest_group"])
    assert np.array_equal(nfact_b200.avg_fdt(golden_subjects[:2]), golden["ingest_group2"])
    folders = [os.path.dirname(p) for p in golden_subjects]
    assert np.array_equal(nfact_b200.process_fdt_matrix2(folders, True), golden["ingest_group"])
    assert np.array_equal(nfact_b200.process_fdt_matrix2(folders[:1], False), golden["ingest_single_1"])


def test_random_against_oracle(tmp_path):
    """Ragged shapes, duplicates, empty rows/columns, several subjects; compared with the oracle."""
    import nfact_b200
    from oracle import nfact_oracle as orc
    rng = np.random.default_rng(11)
    files = []
    for s in range(4):
        folder = tmp_path / f"sub-{s}"
        folder.mkdir()
        nnz = 5000 + 1000 * s
        rows = rng.integers(1, 258, size=nnz)
        cols = rng.integers(5, 391, size=nnz)
        vals = rng.integers(1, 2000, size=nnz)
        with open(folder / "fdt_matrix2.dot", "w") as fh:
            for r, c, v in zip(rows, cols, vals):
                fh.write(f"{r} {c} {v}\n")
            fh.write("257 401 0\n")
        (folder / "waytotal").write_text(str(10_000_000 + 777_777 * s))
        files.append(str(folder / "fdt_matrix2.dot"))
    assert np.array_equal(nfact_b200.load_fdt_matrix(files[2]), orc.load_fdt_matrix(files[2]))
    assert np.array_equal(nfact_b200.avg_fdt(files), orc.avg_fdt(files))
    assert np.array_equal(nfact_b200.avg_fdt(files[:3]), orc.avg_fdt(files[:3]))


def test_empty_matrix_and_errors(tmp_path):
    import nfact_b200
    folder = tmp_path / "sub-empty"
    folder.mkdir()
    (folder / "fdt_matrix2.dot").write_text("7 9 0\n")
    (folder / "waytotal").write_text("1000")
    X = nfact_b200.load_fdt_matrix(str(folder / "fdt_matrix2.dot"))
    assert X.shape == (7, 9) and not X.any()
    (folder / "fdt_matrix2.dot").write_text("8 1 5\n7 9 0\n")       # row index beyond the shape line
    with pytest.raises(ValueError):
        nfact_b200.load_fdt_matrix(str(folder / "fdt_matrix2.dot"))
    with pytest.raises(SystemExit):                                   # reference: error_and_exit -> exit(1)
        nfact_b200.process_fdt_matrix2([str(folder)], False)


def test_resident_ingest_feeds_the_solvers(golden, golden_subjects):
    """load_fdt_matrix(resident=True): the matrix never leaves HBM; the dual regression on it equals the
    one on the host copy bit for bit (same planes, same kernels)."""
    import nfact_b200
    from nfact_b200.device import DeviceMatrix
    xres = nfact_b200.load_fdt_matrix(golden_subjects[0], resident=True)
    assert isinstance(xres, DeviceMatrix) and xres.shape == golden["ingest_single_1"].shape
    assert abs(xres.sumsq - float((golden["ingest_single_1"].astype(np.float64) ** 2).sum())) <= 1e-9 * xres.sumsq
    comps = {"grey_components": golden["cd_fit_W"].astype(np.float64)}
    a = nfact_b200.nmf_dual_regression(comps, xres)
    b = nfact_b200.nmf_dual_regression(comps, golden["ingest_single_1"])
    xres.free()
    assert np.array_equal(a["white_components"], b["white_components"])
    assert np.array_equal(a["grey_components"], b["grey_components"])


def test_lz4_subjects_bit_exact(tmp_path, golden, golden_subjects):
    """fdt_matrix2.dot.lz4 (NFACT/base/matrix_handling.py:108-110) through the native frame decoder: the same
    float32 matrices as the text files, single subject and group average, linked and independent blocks."""
    import shutil
    import nfact_b200
    from oracle import lz4_frame as z
    folders = []
    for s, path in enumerate(golden_subjects):
        folder = tmp_path / f"sub-{s + 1}"
        folder.mkdir()
        shutil.copy(os.path.join(os.path.dirname(path), "waytotal"), folder / "waytotal")
        with open(folder / "fdt_matrix2.dot.lz4", "wb") as fh:
            fh.write(z.compress(open(path, "rb").read(), independent=bool(s % 2), content_checksum=bool(s % 2)))
        folders.append(str(folder))
    for s, folder in enumerate(folders):
        X = nfact_b200.load_fdt_matrix(os.path.join(folder, "fdt_matrix2.dot.lz4"))
        assert np.array_equal(X, golden[f"ingest_single_{s + 1}"])
    assert np.array_equal(nfact_b200.process_fdt_matrix2(folders, True), golden["ingest_group"])
    assert np.array_equal(nfact_b200.process_fdt_matrix2(folders[:1], False), golden["ingest_single_1"])


def test_direct_plane_ingest_equals_the_exact_path(golden):
    """A subject without duplicate coordinates goes straight into the operand planes (no float64 scratch, no
    dense float32 matrix); the planes must be the ones the exact path builds: products with them and the dual
    regression on them are bit-identical.  Subjects with duplicates fall back to the exact path."""
    import torch
    import nfact_b200
    from nfact_b200.device import DeviceMatrix, ingest_coo_resident
    rng = np.random.default_rng(77)
    m, n, k = 203, 333, 5                              # n is not a multiple of 8: padded plane rows
    cells = rng.choice(m * n, size=9000, replace=False)
    rows, cols = (cells // n).astype(np.int32), (cells % n).astype(np.int32)
    vals = rng.uniform(0.5, 900.0, size=cells.size)
    scale = 1e8 / 9876543.0
    dense = np.zeros((m, n), np.float64)
    dense[rows, cols] = vals * scale
    X = dense.astype(np.float32)
    direct = ingest_coo_resident([(rows, cols, vals, scale)], m, n, False)
    exact = DeviceMatrix.from_host(X)
    assert direct.shape == (m, n)
    assert abs(direct.sumsq - exact.sumsq) <= 1e-12 * exact.sumsq and abs(direct.mean - exact.mean) <= 1e-12 * exact.mean
    assert not direct.has_negative
    B = torch.from_numpy(rng.random((k, n)).astype(np.float32)).cuda()
    outs = []
    for mat in (direct, exact):
        out = torch.zeros((k, m), dtype=torch.float32, device="cuda")
        mat.matmul_dev(False, B.data_ptr(), k, out.data_ptr(), m)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    comps = {"grey_components": rng.gamma(0.5, 1.0, (m, k))}
    a = nfact_b200.nmf_dual_regression(comps, direct)
    b = nfact_b200.nmf_dual_regression(comps, exact)
    assert np.array_equal(a["white_components"], b["white_components"])
    assert np.array_equal(a["grey_components"], b["grey_components"])
    direct.free()
    exact.free()
    # the duplicate-free golden subject (probtrackx --pd style) through the file loader
    pd_path = os.path.join(GOLDEN_DIR, "data", "sub-pd", "fdt_matrix2.dot")
    xres = nfact_b200.load_fdt_matrix(pd_path, resident=True)
    want = golden["ingest_single_pd"]
    assert abs(xres.sumsq - float((want.astype(np.float64) ** 2).sum())) <= 1e-12 * xres.sumsq
    g = {"grey_components": rng.gamma(0.5, 1.0, (want.shape[0], 4))}
    assert np.array_equal(nfact_b200.nmf_dual_regression(g, xres)["grey_components"],
                          nfact_b200.nmf_dual_regression(g, want)["grey_components"])
    xres.free()
    # duplicates (summed in float64 by the reference) and a negative value: exact path, flags kept
    rows2 = np.concatenate([rows, rows[:50]])
    cols2 = np.concatenate([cols, cols[:50]])
    vals2 = np.concatenate([vals, vals[:50] * 0.25])
    dense2 = np.zeros((m, n), np.float64)
    np.add.at(dense2, (rows2, cols2), vals2)
    dup = ingest_coo_resident([(rows2, cols2, vals2, scale)], m, n, False)
    want2 = (dense2 * scale).astype(np.float32).astype(np.float64)
    assert abs(dup.sumsq - float((want2 ** 2).sum())) <= 1e-12 * dup.sumsq
    dup.free()
    vals3 = vals.copy()
    vals3[7] = -3.0
    neg = ingest_coo_resident([(rows, cols, vals3, scale)], m, n, False)
    assert neg.has_negative
    neg.free()
    bad = rows.copy()
    bad[3] = m
    with pytest.raises(nfact_b200.NfbError, match="exceeds matrix dimensions"):
        ingest_coo_resident([(bad, cols, vals, scale)], m, n, False)
    empty = ingest_coo_resident([(rows[:0], cols[:0], vals[:0], scale)], m, n, False)
    assert empty.sumsq == 0.0 and empty.shape == (m, n)
    empty.free()
