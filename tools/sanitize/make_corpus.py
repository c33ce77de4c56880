"""Corpora for tools/sanitize/run.sh: valid LZ4 frames (all options of the test encoder) and .dot files plus randomly
damaged and truncated copies of them."""
import os
import sys

import numpy as np

from oracle import lz4_frame as z

out = sys.argv[1]
os.makedirs(os.path.join(out, "lz4")), os.makedirs(os.path.join(out, "dot"))
rng = np.random.default_rng(77)
n = 20000
text = ("\n".join(f"{a} {b} {x}" for a, b, x in zip(rng.integers(1, 60000, n).tolist(), rng.integers(1, 110000, n).tolist(),
                                                      rng.integers(1, 5000, n).tolist())) + "\n").encode()
payloads = [text, bytes(rng.integers(0, 4, 150000, dtype=np.uint8)), b"abcdefgh" * 20000 + b"x" * 70000, b"", b"a",
            bytes(rng.integers(0, 256, 70000, dtype=np.uint8))]
i = 0
for payload in payloads:
    for opts in (dict(independent=False, content_checksum=False),
                 dict(independent=True, content_checksum=True, block_checksum=True),
                 dict(independent=False, content_checksum=False, block_size_id=5, content_size=True)):
        frame = z.compress(payload, **opts)
        open(os.path.join(out, "lz4", f"good_{i}.lz4"), "wb").write(frame)
        i += 1
        if len(frame) < 16:
            continue
        for trial in range(120):
            bad = bytearray(frame)
            for pos in rng.integers(7, len(bad), int(rng.integers(1, 5))):
                bad[pos] = int(rng.integers(0, 256))
            if trial % 8 == 0:
                bad = bad[:int(rng.integers(8, len(bad)))]
            open(os.path.join(out, "lz4", f"bad_{i}.lz4"), "wb").write(bytes(bad))
            i += 1

fmts = ["{:d}", "{:.6g}", "{:.9e}", "{:.3f}", "{!r}"]


def make_dot():
    lines = []
    for _ in range(int(rng.integers(1, 3000))):
        f = fmts[int(rng.integers(len(fmts)))]
        v = float(rng.uniform(0, 1) * 10.0 ** rng.uniform(-6, 10))
        lines.append(f"{int(rng.integers(1, 60000))} {int(rng.integers(1, 110000))} "
                     + f.format(int(v) if f == "{:d}" else v) + ["", " # c", "\r", "  "][int(rng.integers(4))])
        if rng.random() < 0.03:
            lines.append(["", "#x", "   "][int(rng.integers(3))])
    return ("\n".join(lines) + "\n60000 110000 0" + ("\n" if rng.random() < 0.5 else "")).encode()


alphabet = list(b"0123456789 .-+eE#\n\t\rxinfa\x00\xff")
i = 0
for k in range(12):
    t = make_dot()
    open(os.path.join(out, "dot", f"good_{i}.dot"), "wb").write(t)
    i += 1
    for trial in range(60):
        bad = bytearray(t)
        for pos in rng.integers(0, len(bad), int(rng.integers(1, 6))):
            bad[pos] = int(rng.choice(alphabet))
        if trial % 6 == 0:
            bad = bad[:int(rng.integers(0, len(bad)))]
        open(os.path.join(out, "dot", f"bad_{i}.dot"), "wb").write(bytes(bad))
        i += 1
for t in (b"", b"\n", b"#", b"1", b"1 2", b"1 2 3", b"+", b"-", b".", b"1 2 3e", b"1 2 " + b"9" * 400,
          b"1 2 0." + b"0" * 400 + b"1", b" \t\r\n" * 50):
    open(os.path.join(out, "dot", f"edge_{i}.dot"), "wb").write(t)
    i += 1
