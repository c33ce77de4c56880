"""Probe: .dot text -> triplets, host parser (all cores) against the device parser, and text -> resident matrix."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import nfact_b200  # noqa: E402
from nfact_b200.device import get_handle, ingest_dot_resident, parse_dot_coo_device  # noqa: E402
from nfact_b200.matrix_handling import parse_dot_coo  # noqa: E402

n_lines = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
m, n = 59412, 110000
rng = np.random.default_rng(0)
cells = np.unique(rng.integers(0, m * n, int(n_lines * 1.01)))[:n_lines]
rows, cols = cells // n + 1, cells % n + 1
vals = rng.integers(1, 5000, cells.size)
t0 = time.perf_counter()
block = np.empty((cells.size, 3), dtype=np.int64)
block[:, 0], block[:, 1], block[:, 2] = rows, cols, vals
import io
buf = io.BytesIO()
np.savetxt(buf, block, fmt="%d")
text = buf.getvalue() + f"{m} {n} 0\n".encode()
print(f"{cells.size / 1e6:.1f} M lines, {len(text) / 1e6:.0f} MB of text (generated in {time.perf_counter() - t0:.1f} s), "
      f"host cores: {os.cpu_count()}", flush=True)
get_handle()
for rep in range(3):
    t0 = time.perf_counter(); host = parse_dot_coo(text, pinned=True); t_host = time.perf_counter() - t0
    t0 = time.perf_counter(); dev = parse_dot_coo_device(text); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); x = ingest_dot_resident(text, 6.67); t_res = time.perf_counter() - t0
    assert dev is not None and x is not None and np.array_equal(host[2], dev[2]) and np.array_equal(host[0], dev[0])
    x.free()
    del host, dev
    print(f"rep{rep}: host parse {t_host * 1e3:.0f} ms ({cells.size / t_host / 1e6:.0f} M lines/s)   device parse -> host "
          f"triplets {t_dev * 1e3:.0f} ms ({cells.size / t_dev / 1e6:.0f} M lines/s)   text -> resident matrix "
          f"{t_res * 1e3:.0f} ms", flush=True)
