"""First-light script for a fresh B200 box: each stage prints before it runs so a hang is attributable.
Run under `timeout`."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import torch  # noqa: E402

print("torch", torch.__version__, "cuda", torch.cuda.is_available(), torch.cuda.get_device_name(0), flush=True)
import nfact_b200  # noqa: E402
from test_gemm_gpu import _reference, _run_gemm  # noqa: E402

print(nfact_b200.load_library().nfb_version(), flush=True)
for cta_group in (1, 2):
    for (rows, cols, n_b, tr) in [(128, 64, 16, False), (64, 128, 16, True), (300, 517, 48, False),
                                  (517, 300, 48, True), (4100, 3000, 208, False), (3000, 4100, 208, True)]:
        rng = np.random.default_rng(1)
        a = (rng.gamma(0.3, 1.0, size=(rows, cols)) * 300).astype(np.float32)
        b = rng.gamma(0.5, 2.0, size=(n_b, rows if tr else cols)).astype(np.float32)
        print(f"gemm cta_group={cta_group} a={rows}x{cols} n_b={n_b} transposed={tr} ...", end="", flush=True)
        t0 = time.time()
        out = _run_gemm(a, b, tr, cta_group)
        ref = _reference(a, b, tr)
        rel = (out - ref) / (np.abs(ref) + 1e-30)
        print(f" ok {time.time() - t0:.2f}s max|rel|={np.abs(rel).max():.3e} mean(rel)={rel.mean():+.3e}", flush=True)
print("FIRST LIGHT DONE", flush=True)
