"""Probe: host-clock time of NmfState creation / factor upload / short solves at C3 (first-use costs of the streams)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import nfact_b200  # noqa: E402
from bench import WORKLOADS, random_init, regularization, synth_rows_device  # noqa: E402
from nfact_b200.device import DeviceMatrix, NmfState, make_params  # noqa: E402

m, n, k, density = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
x_mean = float(x.sum(dtype=torch.float64)) / (m * n)
xmat = DeviceMatrix.from_device_ptr(x.data_ptr(), m, n, n)
del x
torch.cuda.empty_cache()
w0, h0 = random_init(torch, dev, x_mean, m, n, k, 0)
w0h, h0h = w0.cpu().numpy(), h0.cpu().numpy()
regs = regularization(m, n, k)
for rep in range(4):
    t0 = time.perf_counter()
    st = NmfState(xmat, k, make_params("cd", 3, 0.0, regs))
    t1 = time.perf_counter()
    st.set_factors(w0h, h0h)
    t2 = time.perf_counter()
    st.run()
    t3 = time.perf_counter()
    W, H = st.get_factors()
    t4 = time.perf_counter()
    st.free()
    t5 = time.perf_counter()
    print(f"rep {rep}: create {t1 - t0:.3f}s  set_factors {t2 - t1:.3f}s  run(3 it) {t3 - t2:.3f}s  get {t4 - t3:.3f}s  free {t5 - t4:.3f}s", flush=True)
