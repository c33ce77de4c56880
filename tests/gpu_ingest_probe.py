"""Probe: resident ingest of a C3-shaped subject from page-locked COO triplets (host-clock phases)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import nfact_b200  # noqa: E402
from bench import WORKLOADS, synth_rows_device  # noqa: E402
from nfact_b200.device import get_handle, ingest_coo_resident  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
m, n, k, density = WORKLOADS[name]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
idx = x.nonzero()
rows = idx[:, 0].to(torch.int32).cpu().pin_memory().numpy()
cols = idx[:, 1].to(torch.int32).cpu().pin_memory().numpy()
vals = x[idx[:, 0], idx[:, 1]].to(torch.float64).cpu().pin_memory().numpy()
del x, idx
torch.cuda.empty_cache()
handle = get_handle()
print(f"{name}: {rows.size / 1e6:.1f} M triplets", flush=True)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    t0 = time.perf_counter()
    xres = ingest_coo_resident([(rows, cols, vals, 1.0)], m, n, False, handle)
    dt = time.perf_counter() - t0
    print(f"rep{rep}: ingest {dt * 1e3:.1f} ms", flush=True)
    xres.free()
