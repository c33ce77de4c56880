"""GPU: the device .dot parser (nfb_dot_parse_coo_gpu / nfb_ingest_dot_matrix; replaces np.loadtxt + the index
arithmetic of NFACT/base/matrix_handling.py:93-137) against the host parser (std::from_chars, itself checked against
np.loadtxt on the CPU): identical triplets bit for bit, identical resident matrices, the same errors."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

pytestmark = pytest.mark.gpu


def _both(text: bytes):
    import nfact_b200
    from nfact_b200.device import parse_dot_coo_device
    from nfact_b200.matrix_handling import parse_dot_coo
    host = parse_dot_coo(text)
    dev = parse_dot_coo_device(text)
    return host, dev


def _assert_same(host, dev):
    assert dev is not None, "device parser asked for the host fallback"
    for a, b in zip(host[:3], dev[:3]):
        assert a.dtype == b.dtype and a.shape == b.shape
        assert np.array_equal(a.view(np.uint8), np.asarray(b).view(np.uint8))      # bit for bit (also -0.0)
    assert tuple(host[3]) == tuple(dev[3])


def test_golden_subjects_parse_identically(golden_subjects):
    for path in golden_subjects + [os.path.join(GOLDEN_DIR, "data", "sub-pd", "fdt_matrix2.dot")]:
        text = open(path, "rb").read()
        _assert_same(*_both(text))


def test_number_formats_and_layouts():
    lines = [
        "1 1 5", "2 3 123.456", "  3\t4   1.5e3  ", "4 5 2E-2", "5 6 +7", "6 7 -0", "7 8 007", "8 9 .5", "9 10 5.",
        "10 11 0.000123", "11 12 9007199254740992", "12 13 1e22", "13 14 1e-22", "14 15 123456789012345.6",
        "15 16 0.1", "16 17 299792458", "17 18 6.02214076e23".replace("e23", "e20"), "18 19 4.9e-20", "19 20 1.0E+05",
        "", "   ", "20 21 3.14159\r", "21.0 22.0 1", "22 23 0.30000000000000004"[:-2],
    ]
    text = ("\n".join(lines) + "\n\n40 40 0").encode()          # blank lines, CRLF, no trailing newline
    host, dev = _both(text)
    _assert_same(host, dev)
    assert host[3] == (40, 40) and host[0].size == 22
    with_trailing = text + b"\n  \n"
    _assert_same(*_both(with_trailing))


def test_random_files_match_host_parser():
    rng = np.random.default_rng(31)
    n_lines = 300_000
    r = rng.integers(1, 5000, n_lines)
    c = rng.integers(1, 9000, n_lines)
    ints = rng.integers(0, 100000, n_lines)
    text_int = ("".join(f"{a} {b} {v}\n" for a, b, v in zip(r, c, ints)) + "5000 9000 17\n").encode()
    _assert_same(*_both(text_int))
    vals = rng.uniform(1e-6, 1e6, n_lines)
    text_g = ("".join(f"  {a}  {b}  {v:.6g}\n" for a, b, v in zip(r, c, vals)) + "5000 9000 0\n").encode()
    _assert_same(*_both(text_g))
    text_e = ("".join(f"{a}\t{b}\t{v:.9e}\n" for a, b, v in zip(r[:50000], c[:50000], vals[:50000])) + "5000 9000 0").encode()
    _assert_same(*_both(text_e))
    text_f = ("".join(f"{a} {b} {v:.6f}\n" for a, b, v in zip(r[:50000], c[:50000], vals[:50000])) + "5000 9000 0").encode()
    _assert_same(*_both(text_f))
    # 18 significant digits do not fit the exact fast path (mantissa > 2^53): the device parser says so
    text_long = ("".join(f"{a} {b} {v:.12f}\n" for a, b, v in zip(r[:1000], c[:1000], vals[:1000])) + "5000 9000 0").encode()
    host, dev = _both(text_long)
    assert dev is None and host[0].size == 1000


def test_hard_numbers_fall_back_to_the_host_parser(tmp_path):
    import nfact_b200
    from nfact_b200.device import ingest_dot_resident, parse_dot_coo_device
    for hard in ("9007199254740993", "0.1234567890123456789012", "1e23", "1e-23", "inf", "nan", "-inf"):
        text = f"1 1 5\n2 2 {hard}\n3 3 0\n".encode()
        assert parse_dot_coo_device(text) is None, hard
        assert ingest_dot_resident(text, 1.0) is None
    # load_fdt_matrix takes the host parser then and still produces the matrix
    sub = tmp_path / "sub"
    sub.mkdir()
    (sub / "fdt_matrix2.dot").write_text("1 1 0.1234567890123456789012\n2 3 7\n2 3 1\n")
    (sub / "waytotal").write_text("100000000")
    x = nfact_b200.load_fdt_matrix(str(sub / "fdt_matrix2.dot"), resident=True)
    assert x.shape == (2, 3) and abs(x.sumsq - (np.float32(0.1234567890123456789012) ** 2 + 49.0)) < 1e-6
    x.free()


def test_comments_are_dropped_like_loadtxt():
    """np.loadtxt(comments='#'): everything from a '#' to the end of its line is ignored, on both parsers."""
    text = (b"# probtrackx matrix2\n1 101 117\n   # indented comment\n2 106 234.5 # trailing note\n3 107 96#x\n"
            b"\n#\n+4 +5 +7.5\n5 6 +.5\n99 199 438 # shape")
    host, dev = _both(text)
    _assert_same(host, dev)
    assert host[3] == (99, 199) and host[0].tolist() == [0, 1, 2, 3, 4] and host[2].tolist() == [117, 234.5, 96, 7.5, 0.5]
    # a comment longer than a parser segment, and comments at segment boundaries
    rng = np.random.default_rng(5)
    lines = []
    for i in range(20000):
        lines.append(f"{int(rng.integers(1, 900))} {int(rng.integers(1, 700))} {int(rng.integers(0, 5000))}")
        if i % 37 == 0:
            lines.append("#" + "c" * int(rng.integers(0, 400)))
        if i % 53 == 0:
            lines[-1] += "   # " + "n" * int(rng.integers(0, 200))
    _assert_same(*_both(("\n".join(lines) + "\n900 700 0\n").encode()))


def test_errors_match_the_host_parser():
    import nfact_b200
    from nfact_b200.device import parse_dot_coo_device
    from nfact_b200.matrix_handling import parse_dot_coo
    bad = {
        b"1 1 5\n2 2\n3 3 0\n": "malformed",
        b"1 1 5 9\n3 3 0\n": "malformed",
        b"1 x 5\n3 3 0\n": "malformed",
        b"1 1 5e\n3 3 0\n": "malformed",
        b"4 1 5\n3 3 0\n": "row index exceeds",
        b"1 4 5\n3 3 0\n": "column index exceeds",
        b"0 1 5\n3 3 0\n": "negative row",
        b"1 -1 5\n3 3 0\n": "negative column",
        b"1e12 1 5\n3 3 0\n": "32-bit",
        b"+-1 1 5\n3 3 0\n": "malformed",
        b"1 1 +\n3 3 0\n": "malformed",
        b"1 1 #5\n3 3 0\n": "malformed",
    }
    for text, what in bad.items():
        with pytest.raises(Exception, match=what):
            parse_dot_coo_device(text)
        with pytest.raises(Exception, match=what):
            parse_dot_coo(text)
    with pytest.raises(Exception, match="empty"):
        parse_dot_coo_device(b"   \n\n")
    only_shape = parse_dot_coo_device(b"7 9 0\n")
    assert only_shape[3] == (7, 9) and only_shape[0].size == 0


def test_device_parsed_subject_equals_host_parsed_subject(golden):
    """The whole load on the device (text -> planes) against host parse + exact ingest: bit-identical regressions."""
    import nfact_b200
    from nfact_b200.device import ingest_dot_resident
    path = os.path.join(GOLDEN_DIR, "data", "sub-pd", "fdt_matrix2.dot")
    text = open(path, "rb").read()
    scale = 1e8 / float(open(os.path.join(os.path.dirname(path), "waytotal")).read())
    xdev = ingest_dot_resident(text, scale)
    assert xdev is not None and xdev.shape == golden["ingest_single_pd"].shape
    want = golden["ingest_single_pd"]
    assert abs(xdev.sumsq - float((want.astype(np.float64) ** 2).sum())) <= 1e-12 * xdev.sumsq
    rng = np.random.default_rng(2)
    g = {"grey_components": rng.gamma(0.5, 1.0, (want.shape[0], 4))}
    a = nfact_b200.nmf_dual_regression(g, xdev)
    b = nfact_b200.nmf_dual_regression(g, want)
    xdev.free()
    assert np.array_equal(a["grey_components"], b["grey_components"])
    assert np.array_equal(a["white_components"], b["white_components"])
    # duplicate coordinates (the integer golden subjects): reported for the exact path, which load_fdt_matrix takes
    dup_text = open(os.path.join(GOLDEN_DIR, "data", "sub-1", "fdt_matrix2.dot"), "rb").read()
    assert ingest_dot_resident(dup_text, 1.0) is None
    xres = nfact_b200.load_fdt_matrix(os.path.join(GOLDEN_DIR, "data", "sub-1", "fdt_matrix2.dot"), resident=True)
    assert abs(xres.sumsq - float((golden["ingest_single_1"].astype(np.float64) ** 2).sum())) <= 1e-9 * xres.sumsq
    xres.free()
