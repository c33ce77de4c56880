"""Probe: dual-regression timing on synthetic C2/C3-shaped subjects (resident X)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import nfact_b200  # noqa: E402
from bench import WORKLOADS, synth_rows_device  # noqa: E402
from nfact_b200.device import DeviceMatrix  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
m, n, k, density = WORKLOADS[name]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
torch.cuda.synchronize()
xmat = DeviceMatrix.from_device_ptr(x.data_ptr(), m, n, n)
del x
rng = np.random.default_rng(0)
G = rng.gamma(0.3, 1.0, (m, k)) * (rng.random((m, k)) < 0.3)        # sparse-ish group components
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    t0 = time.perf_counter()
    dr = nfact_b200.nmf_dual_regression({"grey_components": G}, xmat)
    dt = time.perf_counter() - t0
    Hs, Gs = dr["white_components"], dr["grey_components"]
    print(f"{name} rep{rep}: {dt:.3f}s  Hs nnz={np.mean(Hs > 0):.3f} Gs nnz={np.mean(Gs > 0):.3f}", flush=True)
