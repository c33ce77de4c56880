"""Probe: the NFACT default call nmf_decomp(get_parameters(None, "nmf", k), X) -- init='nndsvd', solver='cd',
tol=1e-4, max_iter=200 -- on a synthetic C3 (or C2) matrix held in host memory.  NFB_HOST_TIMING=1 prints phases."""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402
import nfact_b200  # noqa: E402
from bench import WORKLOADS, synth_rows_device  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
m, n, k, density = WORKLOADS[name]
dev = torch.device("cuda", 0)
x = synth_rows_device(torch, dev, 0, m, n, k, density)
x_host = torch.empty((m, n), dtype=torch.float32, pin_memory=True)
x_host.copy_(x)
del x
torch.cuda.empty_cache()
params = nfact_b200.get_parameters(None, "nmf", k)
for rep in range(2):
    t0 = time.perf_counter()
    res = nfact_b200.nmf_decomp(dict(params, random_state=0), x_host.numpy())
    dt = time.perf_counter() - t0
    w, h = res["grey_components"], res["white_components"]
    print(f"{name} rep{rep}: {dt:.2f} s  W nnz={float((w > 0).mean()):.3f} H nnz={float((h > 0).mean()):.3f}", flush=True)
