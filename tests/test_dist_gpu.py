"""GPU, >= 2 devices: the row-sharded solver under torchrun (tests/dist_parity_gpu.py) as a collected test.

Spawns one rank per GPU (2 ranks; all visible GPUs with NFB_DIST_TEST_RANKS=all), checks the exit status and the
"DIST PARITY OK" verdict, and leaves the ranks' output in gpurun_out/dist_parity_n<N>.log (copied to profiles/
as the tracked evidence of the multi-GPU parity run).  Skipped on a one-GPU box.
"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_row_sharded_solver_parity_under_torchrun():
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least 2 GPUs (run through `gpurun --gpus 2`)")
    ranks = n_gpus if os.environ.get("NFB_DIST_TEST_RANKS") == "all" else 2
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_parity_gpu.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"dist_parity_n{ranks}.log"), "w") as fh:
        fh.write("$ " + " ".join(cmd[1:]) + "\n" + proc.stdout + "\n---- stderr (tail) ----\n" + proc.stderr[-4000:])
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "DIST PARITY OK" in proc.stdout
