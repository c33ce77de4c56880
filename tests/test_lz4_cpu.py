"""CPU: the native LZ4 frame decoder (nfb_lz4_frame_*; replaces lz4.frame.open at
NFACT/base/matrix_handling.py:108-110) against the published format's known answers and the oracle codec."""
import os
import shutil
import struct

import numpy as np
import pytest

from conftest import GOLDEN_DIR


def _native(frame: bytes) -> bytes:
    import nfact_b200
    return nfact_b200.decompress_lz4_frame(frame).tobytes()


def test_xxh32_and_header_known_answers():
    from oracle import lz4_frame as z
    # xxHash's published XXH32 values (seed 0)
    assert z.xxh32(b"") == 0x02CC5D05
    assert z.xxh32(b"a") == 0x550D7456
    assert z.xxh32(b"abc") == 0x32D153FF
    assert z.xxh32(b"Nobody inspects the spammish repetition") == 0xE2293B2F
    # the frame headers every `lz4` command-line file starts with, and the empty frame
    assert z.compress(b"", 4) == bytes.fromhex("04224d186440a700000000055dcc02")
    assert z.compress(b"", 7)[:7] == bytes.fromhex("04224d186470b9")
    assert _native(bytes.fromhex("04224d186440a700000000055dcc02")) == b""
    assert _native(bytes.fromhex("04224d186470b900000000055dcc02")) == b""
    assert _native(b"") == b""


def test_hand_assembled_frame():
    """A frame written out by hand from the specification (no encoder involved): literals 'abcd', a match of
    length 12 at offset 4 (overlapping, run-length style), then the closing 5 literals."""
    from oracle import lz4_frame as z
    block = bytes([0x48]) + b"abcd" + struct.pack("<H", 4) + bytes([0x50]) + b"vwxyz"
    content = b"abcd" + b"abcd" * 3 + b"vwxyz"
    desc = bytes([0x60, 0x40])                      # version 01, independent blocks, no checksums; 64 KB
    frame = struct.pack("<I", 0x184D2204) + desc + bytes([(z.xxh32(desc) >> 8) & 0xFF]) + \
        struct.pack("<I", len(block)) + block + struct.pack("<I", 0)
    assert z.decompress(frame) == content
    assert _native(frame) == content


def _payloads():
    rng = np.random.default_rng(5)
    text = "".join(f"{r} {c} {v}\n" for r, c, v in zip(rng.integers(1, 3000, 40000), rng.integers(1, 9000, 40000),
                                                       rng.integers(1, 700, 40000))).encode() + b"3000 9000 0\n"
    return {
        "dot_text": text,
        "random_bytes": rng.integers(0, 256, 200_001, dtype=np.uint8).tobytes(),   # stored blocks
        "zeros": bytes(300_000),                                                   # offset-1 runs, long lengths
        "period3": b"xyz" * 50_000,
        "tiny": b"1 1 1\n",
        "exactly_one_block": text[:65536],
        "one_block_plus_one": text[:65537],
    }


@pytest.mark.parametrize("options", [
    dict(),                                                         # lz4 CLI style: independent + content checksum
    dict(independent=False, content_checksum=False),                # python-lz4 defaults (what nfact_config writes)
    dict(independent=False, content_checksum=True, block_checksum=True, content_size=True),
    dict(block_size_id=5, block_checksum=True),
    dict(block_size_id=4, store_incompressible=False),
])
def test_round_trips(options):
    from oracle import lz4_frame as z
    for name, payload in _payloads().items():
        frame = z.compress(payload, **options)
        assert z.decompress(frame) == payload, name
        assert _native(frame) == payload, (name, options)


def test_concatenated_and_skippable_frames():
    from oracle import lz4_frame as z
    a, b = b"first frame " * 3000, b"second " * 20000
    skippable = struct.pack("<II", 0x184D2A53, 7) + b"ignored"
    blob = z.compress(a, independent=False) + skippable + z.compress(b, block_checksum=True) + skippable
    assert z.decompress(blob) == a + b
    assert _native(blob) == a + b


def test_corruption_is_detected():
    from oracle import lz4_frame as z
    payload = _payloads()["dot_text"]
    good = z.compress(payload, block_checksum=True, content_checksum=True, content_size=True)
    cases = {
        "magic": b"\x00" + good[1:],
        "header checksum": good[:6] + bytes([good[6] ^ 0x40]) + good[7:],
        "truncated": good[:len(good) // 2],
        "no end mark": good[:-8],
        "block payload": good[:5000] + bytes([good[5000] ^ 0xFF]) + good[5001:],
        "content checksum": good[:-1] + bytes([good[-1] ^ 1]),
    }
    for name, bad in cases.items():
        with pytest.raises(RuntimeError, match="LZ4F"):
            _native(bad)
    # no checksums at all: a damaged sequence stream must still fail cleanly (bounds checks), never crash
    plain = bytearray(z.compress(payload, content_checksum=False))
    rng = np.random.default_rng(0)
    for _ in range(200):
        bad = bytearray(plain)
        for pos in rng.integers(7, len(bad), 3):
            bad[pos] = int(rng.integers(0, 256))
        try:
            out = _native(bytes(bad))
            assert len(out) <= 2 * len(payload) + (1 << 22)
        except RuntimeError:
            pass
    # wrong content size in the header
    wrong = bytearray(z.compress(payload, content_size=True, content_checksum=False))
    wrong[6:14] = struct.pack("<Q", len(payload) + 1)
    desc = bytes(wrong[4:14])
    wrong[14] = (z.xxh32(desc) >> 8) & 0xFF
    with pytest.raises(RuntimeError, match="frameSize"):
        _native(bytes(wrong))


def test_compressed_subject_parses_like_the_text(tmp_path, golden_subjects):
    """read_in_matrix / load_single_matrix / process_fdt_matrix2's file pick on fdt_matrix2.dot.lz4."""
    import nfact_b200
    from oracle import lz4_frame as z
    from oracle import nfact_oracle as orc
    src = os.path.dirname(golden_subjects[0])
    folder = tmp_path / "sub-1"
    folder.mkdir()
    shutil.copy(os.path.join(src, "waytotal"), folder / "waytotal")
    text = open(golden_subjects[0], "rb").read()
    lz4_path = str(folder / "fdt_matrix2.dot.lz4")
    with open(lz4_path, "wb") as fh:
        fh.write(z.compress(text, independent=False, content_checksum=False))   # python-lz4's defaults
    assert np.array_equal(nfact_b200.read_in_matrix(lz4_path), orc.read_dot(golden_subjects[0]))
    coo, ref = nfact_b200.load_single_matrix(lz4_path), nfact_b200.load_single_matrix(golden_subjects[0])
    assert coo.shape == ref.shape and coo.scale == ref.scale
    for got, want in zip(coo[:3], ref[:3]):
        assert np.array_equal(got, want)
    with open(lz4_path, "wb") as fh:
        fh.write(b"definitely not lz4")
    with pytest.raises(RuntimeError, match="frameType"):
        nfact_b200.read_in_matrix(lz4_path)
    assert os.path.exists(os.path.join(GOLDEN_DIR, "data", "sub-1", "fdt_matrix2.dot"))


def _decode_with(frame: bytes, nthreads: int):
    import ctypes
    from nfact_b200._lib import load
    lib = load()
    src = np.frombuffer(frame, dtype=np.uint8)
    size, written = ctypes.c_int64(), ctypes.c_int64()
    assert lib.nfb_lz4_frame_size(ctypes.c_void_p(src.ctypes.data), src.size, nthreads, ctypes.byref(size)) == 0
    out = np.empty(max(size.value, 1), dtype=np.uint8)
    rc = lib.nfb_lz4_frame_decompress(ctypes.c_void_p(src.ctypes.data), src.size, ctypes.c_void_p(out.ctypes.data),
                                      size.value, nthreads, ctypes.byref(written))
    assert rc == 0 and written.value == size.value
    chunks, deferred = ctypes.c_int64(), ctypes.c_int64()
    lib.nfb_lz4_last_chunks(ctypes.byref(chunks), ctypes.byref(deferred))
    return out[:size.value].tobytes(), chunks.value, deferred.value


def test_linked_frames_are_decoded_in_parallel_chunks(monkeypatch):
    """python-lz4's default frames (linked 64 KB blocks) are split into chunks decoded side by side: a chunk
    reconstructs the bytes in front of it in a scratch stream with exact taint tracking and starts on its own once its
    first 64 KB are untainted; chunks that stay tainted (copy chains back to the start of the frame) wait for their
    predecessor.  Either way the output equals the sequential decode and the format oracle."""
    from oracle import lz4_frame as z
    rng = np.random.default_rng(11)
    n = 120_000
    text = ("\n".join(f"{a} {b} {v}" for a, b, v in zip(np.sort(rng.integers(1, 60000, n)).tolist(),
                                                          rng.integers(1, 110000, n).tolist(),
                                                          rng.integers(1, 5000, n).tolist())) + "\n").encode()
    periodic = b"abcdefgh" * 150_000 + text[:300_000]          # every match chains back: the taint never clears
    mixed = text[:500_000] + bytes(rng.integers(0, 256, 200_000, dtype=np.uint8)) + text[500_000:]   # stored blocks inside
    monkeypatch.setenv("NFB_LZ4_CHUNK_BYTES", str(1 << 17))
    saw_clean = saw_deferred = False
    for warm in (1 << 16, 1 << 18, 1 << 20):
        monkeypatch.setenv("NFB_LZ4_WARM_BYTES", str(warm))
        for name, payload in (("text", text), ("periodic", periodic), ("mixed", mixed)):
            for opts in (dict(independent=False, content_checksum=False),
                         dict(independent=False, content_checksum=True, block_checksum=True, block_size_id=5)):
                frame = z.compress(payload, **opts)
                sequential, chunks1, _ = _decode_with(frame, 1)
                assert sequential == payload and chunks1 == 0
                for nthreads in (2, 5, 8):
                    got, chunks, deferred = _decode_with(frame, nthreads)
                    assert got == payload, (name, warm, nthreads)
                    assert chunks >= 1 and 0 <= deferred <= chunks
                    saw_clean |= deferred < chunks
                    saw_deferred |= deferred > 0
                # two frames back to back, the second one chunked as well
                got, chunks, _ = _decode_with(frame + frame, 4)
                assert got == payload + payload and chunks >= 2
    assert saw_clean and saw_deferred


def test_damaged_linked_frames_fail_cleanly_in_chunked_mode(monkeypatch):
    from oracle import lz4_frame as z
    monkeypatch.setenv("NFB_LZ4_CHUNK_BYTES", str(1 << 17))
    monkeypatch.setenv("NFB_LZ4_WARM_BYTES", str(1 << 17))
    payload = _payloads()["dot_text"] * 6
    frame = bytearray(z.compress(payload, independent=False, content_checksum=False))
    rng = np.random.default_rng(2)
    for _ in range(150):
        bad = bytearray(frame)
        for pos in rng.integers(7, len(bad), 3):
            bad[pos] = int(rng.integers(0, 256))
        try:
            out = _native(bytes(bad))
            assert len(out) <= 2 * len(payload) + (1 << 22)
        except RuntimeError:
            pass
