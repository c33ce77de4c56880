"""CPU: the C-ABI library loads and exports every symbol include/nfact_b200.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nfact_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nfb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import nfact_b200
    from nfact_b200 import _lib
    lib = nfact_b200.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nfact_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"
    assert b"sm_100a" in lib.nfb_version()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path fails loudly instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import nfact_b200
    from nfact_b200._lib import NfbError
    with pytest.raises((NfbError, SystemExit)):
        nfact_b200.nmf_decomp(nfact_b200.get_parameters(None, "nmf", 3), np.ones((8, 9), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "nfact_b200")
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, name)).read()
                assert "oracle" not in text.lower() or name == "nothing", f"{name} mentions the oracle"


def test_get_parameters_matches_reference(golden):
    import nfact_b200
    p = nfact_b200.get_parameters(None, "nmf", 10)
    assert sorted(p.keys()) == list(golden["params_keys"])
    assert repr(sorted(p.items(), key=lambda kv: kv[0])) == str(golden["params_repr"])
    cfg = {"init": "random", "solver": "mu"}
    assert nfact_b200.get_parameters(cfg, "nmf", 7) == {"init": "random", "solver": "mu", "n_components": 7}
    regs = nfact_b200.compute_regularization(99, 199, p["alpha_W"], p["alpha_H"], p["l1_ratio"])
    assert np.array_equal(np.array(regs, dtype=np.float64), golden["regularization"])


def test_host_parsing_matches_oracle(golden_subjects):
    import nfact_b200
    from oracle import nfact_oracle as orc
    mat = nfact_b200.read_in_matrix(golden_subjects[0])
    assert np.array_equal(mat, orc.read_dot(golden_subjects[0]))
    coo = nfact_b200.load_single_matrix(golden_subjects[0])
    assert coo.shape == (99, 199) and coo.rows.dtype == np.int32 and coo.rows.min() >= 0
    assert coo.scale == 1e8 / 15000000.0


def test_parameter_validation():
    import nfact_b200
    with pytest.raises(TypeError):
        nfact_b200.NMF(n_components=3, not_a_param=1)
    with pytest.raises(ValueError):
        nfact_b200.NMF(n_components=3, beta_loss="kullback-leibler")
    with pytest.raises(ValueError):
        nfact_b200.NMF(n_components=3, solver="xx")


def test_native_dot_parser_matches_loadtxt(tmp_path):
    """nfb_dot_parse == np.loadtxt on awkward but legal text; malformed input raises like loadtxt."""
    import nfact_b200
    text = ("1 101 117\n  2\t106   234.5  \n3 107 9.6e1\r\n\n4 5 +7\n5 6 1e-3\n6 7 0.1\n"
            "7 8 12345678901234567890\n99 199 438")      # no trailing newline, a blank line, CRLF, tabs
    path = tmp_path / "fdt_matrix2.dot"
    path.write_text(text)
    got = nfact_b200.read_in_matrix(str(path))
    ref = np.loadtxt(str(path))
    assert got.shape == ref.shape and np.array_equal(got, ref)
    rng = np.random.default_rng(0)
    big = tmp_path / "big.dot"
    rows = rng.integers(1, 5000, 300000)
    vals = rng.uniform(0, 1e6, 300000)
    with open(big, "w") as fh:
        for r, v in zip(rows, vals):
            fh.write(f"{r} {r + 1} {float(v)!r}\n")
        fh.write("5000 5001 0\n")
    got = nfact_b200.read_in_matrix(str(big))
    assert got.shape == (300001, 3)
    assert np.array_equal(got[:-1, 2], vals) and np.array_equal(got[:-1, 0], rows.astype(np.float64))
    bad = tmp_path / "bad.dot"
    bad.write_text("1 2 3\n4 5\n6 7 8\n")
    with pytest.raises(ValueError):
        nfact_b200.read_in_matrix(str(bad))
    bad.write_text("1 2 x\n")
    with pytest.raises(ValueError):
        nfact_b200.read_in_matrix(str(bad))
    bad.write_text("")
    with pytest.raises(ValueError):
        nfact_b200.read_in_matrix(str(bad))


def test_native_dot_parser_comments_and_signs(tmp_path):
    """np.loadtxt (comments='#') drops everything from a '#' to the end of the line and skips what is empty then;
    a leading '+' is a sign, '+-' is not a number."""
    import nfact_b200
    from nfact_b200.matrix_handling import parse_dot_coo
    text = ("# probtrackx matrix2\n1 101 117\n   # indented comment\n2 106 234.5 # trailing note\n3 107 96#x\n"
            "\n#\n+4 +5 +7.5\n5 6 +.5\n99 199 438 # shape")
    path = tmp_path / "fdt_matrix2.dot"
    path.write_text(text)
    got = nfact_b200.read_in_matrix(str(path))
    ref = np.loadtxt(str(path))
    assert got.shape == ref.shape == (6, 3) and np.array_equal(got, ref)
    rows, cols, vals, shape = parse_dot_coo(text.encode())
    assert shape == (99, 199) and np.array_equal(rows, ref[:-1, 0].astype(np.int32) - 1)
    assert np.array_equal(cols, ref[:-1, 1].astype(np.int32) - 1) and np.array_equal(vals, ref[:-1, 2])
    for bad in ("1 2 #3\n4 5 6\n", "+-1 2 3\n", "1 2 +\n", "1 + 2 3\n"):
        path.write_text(bad)
        with pytest.raises(ValueError):
            np.loadtxt(str(path))
        with pytest.raises(ValueError):
            nfact_b200.read_in_matrix(str(path))


def test_native_dot_parser_fuzz_against_loadtxt(tmp_path):
    """Random legal files in the number formats probtrackx and printf produce: the native parser returns exactly the
    float64 array np.loadtxt returns (correct rounding of every decimal string)."""
    import nfact_b200
    rng = np.random.default_rng(20261020)
    formats = ["{:d}", "{:.6g}", "{:.9e}", "{:.3f}", "{:.17g}", "{!r}", "{:.1f}", "{:E}"]
    seps = [" ", "  ", "\t", " \t "]
    path = tmp_path / "fuzz.dot"
    for trial in range(25):
        n = int(rng.integers(1, 400))
        lines = []
        for _ in range(n):
            fmt = formats[int(rng.integers(len(formats)))]
            mag = 10.0 ** rng.uniform(-8, 12)
            v = float(rng.uniform(0, 1) * mag)
            val = fmt.format(int(v) if fmt == "{:d}" else v)
            sep = seps[int(rng.integers(len(seps)))]
            lead = " " * int(rng.integers(0, 3))
            tail = ["", " ", "\r", " # c"][int(rng.integers(4))]
            lines.append(f"{lead}{int(rng.integers(1, 70000))}{sep}{int(rng.integers(1, 120000))}{sep}{val}{tail}")
            if rng.random() < 0.05:
                lines.append(["", "   ", "# comment", "\t"][int(rng.integers(4))])
        lines.append("70000 120000 0")
        path.write_text("\n".join(lines) + ("\n" if trial % 2 else ""))
        got = nfact_b200.read_in_matrix(str(path))
        ref = np.loadtxt(str(path), ndmin=2)
        assert got.shape == ref.shape and np.array_equal(got, ref), trial


def test_native_dot_parser_fast_path_boundaries(tmp_path):
    """The one-operation decimal fast path of the host parser and its hand-over to std::from_chars: digit strings
    around 2^53 and 19 significant digits, up to 25 fraction digits, signs, bare '.', exponents -- every value equal to
    Python's float() of the same token (np.loadtxt's conversion), bit for bit."""
    import nfact_b200
    rng = np.random.default_rng(7)
    tokens = ["9007199254740992", "9007199254740993", "9007199254740991", "90071992547409925", "1234567890123456789",
              "12345678901234567890", "0.1", "0.30000000000000004", "5.", ".5", "-0", "-0.0", "+7.25", "007.5000",
              "0.0000000000000000000001", "0.00000000000000000000001", "123456789.123456789", "1.7976931348623157e308",
              "4.9e-324", "1e22", "1e23", "8.5e-5", "2.5E+3", "123456789012345.6789", "0.1234567890123456789012",
              "99999999999999999999", "0.000000000000000000000000000001", "1.5e", "3.25e+"]
    for _ in range(400):
        digits = int(rng.integers(1, 22))
        frac = int(rng.integers(0, 26))
        body = "".join(str(int(d)) for d in rng.integers(0, 10, digits))
        tail = "".join(str(int(d)) for d in rng.integers(0, 10, frac))
        tokens.append(("-" if rng.random() < 0.2 else "") + body + ("." + tail if frac else ""))
    good = [t for t in tokens if not t.endswith(("e", "e+"))]
    path = tmp_path / "fdt_matrix2.dot"
    path.write_text("".join(f"{i + 1} {i + 2} {t}\n" for i, t in enumerate(good)) + "5000 5000 0\n")
    got = nfact_b200.read_in_matrix(str(path))
    want = np.array([float(t) for t in good])
    assert got.shape == (len(good) + 1, 3)
    assert np.array_equal(got[:-1, 2].view(np.uint64), want.view(np.uint64))        # bit for bit, also -0.0
    assert np.array_equal(got, np.loadtxt(str(path)))
    for bad in ("1.5e", "3.25e+", ".", "-", "+", "1..2", "--1"):
        path.write_text(f"1 2 {bad}\n3 3 0\n")
        with pytest.raises(ValueError):
            nfact_b200.read_in_matrix(str(path))


def test_fused_coo_parser_matches_the_reference_construction(tmp_path, golden_subjects):
    """nfb_dot_parse_coo == loadtxt + `[:-1] - 1` + integer cast + last-line shape
    (NFACT/base/matrix_handling.py:131-137), with scipy's index errors."""
    import nfact_b200
    from nfact_b200.matrix_handling import parse_dot_coo
    for path in golden_subjects:
        ref = np.loadtxt(path)
        rows, cols, vals, shape = parse_dot_coo(open(path, "rb").read(), path)
        assert rows.dtype == np.int32 and cols.dtype == np.int32 and vals.dtype == np.float64
        assert np.array_equal(rows, (ref[:-1, 0] - 1).astype(np.int64))
        assert np.array_equal(cols, (ref[:-1, 1] - 1).astype(np.int64))
        assert np.array_equal(vals, ref[:-1, 2])
        assert shape == (int(ref[-1, 0]), int(ref[-1, 1]))
    rng = np.random.default_rng(1)
    big = tmp_path / "sub" / "fdt_matrix2.dot"
    big.parent.mkdir()
    r = rng.integers(1, 4001, 250000)
    c = rng.integers(1, 9001, 250000)
    v = rng.uniform(0, 1e5, 250000)
    with open(big, "w") as fh:
        for a, b, x in zip(r, c, v):
            fh.write(f"{a}  {b}  {float(x)!r}\n")
        fh.write("4000  9000  0\n")
    (big.parent / "waytotal").write_text("12500000\n")
    coo = nfact_b200.load_single_matrix(str(big))
    assert coo.shape == (4000, 9000) and coo.scale == 1e8 / 12500000.0
    assert np.array_equal(coo.rows, r - 1) and np.array_equal(coo.cols, c - 1) and np.array_equal(coo.data, v)
    only_shape = tmp_path / "empty.dot"
    only_shape.write_text("7 9 0\n")
    rows, cols, vals, shape = parse_dot_coo(only_shape.read_bytes())
    assert rows.size == 0 and shape == (7, 9)
    for text, msg in (("0 1 5\n3 3 0\n", "negative row"), ("1 0 5\n3 3 0\n", "negative column"),
                      ("4 1 5\n3 3 0\n", "row index exceeds"), ("1 4 5\n3 3 0\n", "column index exceeds"),
                      ("1 2\n3 3 0\n", "could not convert"), ("1 2 3 4\n3 3 0\n", "could not convert")):
        with pytest.raises(ValueError, match=msg):
            parse_dot_coo(text.encode())


def test_sso_clustering_glue_matches_reference():
    """sim2dis / clustering_components / idx2centrotype (host code of nfact_b200.nmfsso) against the outputs of
    the reference's own sso_functions.py on the committed golden stack (tests/golden/make_golden_sso.py)."""
    from conftest import GOLDEN_DIR
    from nfact_b200 import nmfsso
    g = np.load(os.path.join(GOLDEN_DIR, "golden_sso.npz"))
    dis = nmfsso.sim2dis(g["sim"])
    assert np.array_equal(dis, g["dis"])
    part = nmfsso.clustering_components(dis, int(g["dim"]))
    assert np.array_equal(part, g["partition"])
    assert np.array_equal(nmfsso.idx2centrotype(g["sim"], part), g["centrotypes"])
