"""Probe: dual-regression / NNLS timing at C3 on the components of a real NMF solve, one process, several kernel
settings (the NFB_NNLS_* knobs are read on every call):

    NFB_PHASE_TIMING=1 NFB_DR_STATS=1 python tests/gpu_nnls_probe.py [C2|C3] [nmf iterations] ["K=V,K=V;K=V;..."]

Each ';'-separated group is a set of environment assignments applied before one timed regression (default: the
shipped settings, then sweeps + CG only, then one greedy step per key pass)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch  # noqa: E402
import nfact_b200  # noqa: E402
from bench import WORKLOADS, random_init, synth_rows_device  # noqa: E402
from nfact_b200.device import DeviceMatrix  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
groups = sys.argv[3].split(";") if len(sys.argv) > 3 else ["", "NFB_NNLS_GREEDY=0", "NFB_NNLS_BATCH=1", "NFB_NNLS_BATCH=2"]
m, n, k, density = WORKLOADS[name]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
x_mean = float(x.sum(dtype=torch.float64)) / (m * n)
torch.cuda.synchronize()
xmat = DeviceMatrix.from_device_ptr(x.data_ptr(), m, n, n)
dump = os.environ.get("NFB_PROBE_DUMP")        # path: Gram matrices + sampled right-hand sides for CPU simulations
if not dump:
    del x
    torch.cuda.empty_cache()
w0, h0 = random_init(torch, dev, x_mean, m, n, k, 0)
params = dict(nfact_b200.get_parameters(None, "nmf", k), tol=0.0, max_iter=iters)
t0 = time.perf_counter()
res = nfact_b200.nmf_decomp(params, xmat, W=w0.cpu().numpy(), H=h0.cpu().numpy())
print(f"{name}: {iters} CD iterations in {time.perf_counter() - t0:.2f}s; grey nnz={np.mean(res['grey_components'] > 0):.3f}",
      flush=True)
comps = {"grey_components": res["grey_components"]}
if dump:
    rng = np.random.default_rng(0)
    cols = np.sort(rng.choice(n, 256, replace=False))
    rows = np.sort(rng.choice(m, 256, replace=False))
    wd = torch.from_numpy(res["grey_components"]).to(dev, torch.float64)
    g1 = (wd.T @ wd).cpu().numpy()
    r1 = (wd.T @ x[:, torch.from_numpy(cols).to(dev)].to(torch.float64)).cpu().numpy()
    dr0 = nfact_b200.nmf_dual_regression(comps, xmat)
    hs = torch.from_numpy(dr0["white_components"]).to(dev, torch.float64)
    g2 = (hs @ hs.T).cpu().numpy()
    r2 = (hs @ x[torch.from_numpy(rows).to(dev), :].to(torch.float64).T).cpu().numpy()
    np.savez(dump, G1=g1, R1=r1, G2=g2, R2=r2, H1=dr0["white_components"][:, cols],
             H2=dr0["grey_components"][rows, :].T)
    del x, wd, hs, dr0
    torch.cuda.empty_cache()
reference = None
touched = set()
for group in groups:
    for key in touched:
        os.environ.pop(key, None)
    touched = set()
    for assignment in filter(None, group.split(",")):
        key, value = assignment.split("=")
        os.environ[key] = value
        touched.add(key)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        dr = nfact_b200.nmf_dual_regression(comps, xmat)
        times.append(time.perf_counter() - t0)
    if reference is None:
        reference = dr
    diff = max(np.linalg.norm(dr[key] - reference[key]) / np.linalg.norm(reference[key])
               for key in ("grey_components", "white_components"))
    print(f"[{group or 'default'}] best of 3: {min(times) * 1e3:.1f} ms   rel diff to the first setting {diff:.1e}",
          flush=True)
