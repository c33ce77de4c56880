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


def _torchrun(ranks, script_args, log_name, env=None):
    port = 29500 + (os.getpid() + len(log_name)) % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}",
           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, **(env or {})))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, log_name), "a") as fh:
        fh.write("$ " + " ".join(cmd[1:]) + "\n" + proc.stdout[-6000:] + "\n---- stderr (tail) ----\n" + proc.stderr[-3000:] + "\n")
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    return proc


def test_reference_tools_on_two_gpus(tmp_path):
    """`torchrun ... nfact_b200.dropin nfact_decomp / nfact_dr`: rank 0 runs the reference's main, rank 1 serves.
    nfact_decomp without sso = ONE row-sharded solve with NFACT's default NNDSVD init computed across the ranks;
    with sso = the random-init runs spread over the GPUs (replicas) + final solve; nfact_dr = subjects round-robin.
    Outputs against a single-process run of the same tool (CD gate) / against direct calls."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs (run through `gpurun --gpus 2`)")
    from test_reference_mains_gpu import M, N, _cd_gate, _reference_root, _study
    ref = _reference_root()
    if ref is None:
        pytest.skip("no reference checkout (NFB_REFERENCE_PATH, /root/reference, baseline/_ref)")
    script = os.path.join(ROOT, "tests", "dist_dropin_gpu.py")
    subs, sub_list, seed_list, out = _study(tmp_path)
    out1 = str(tmp_path / "out1")
    os.makedirs(out1)
    from test_reference_mains_gpu import _config
    config, _ = _config(tmp_path, random_state=0)      # repeatable NNDSVD draw: N = 2 vs N = 1 differ by rounding only
    common = ["-l", sub_list, "-s", seed_list, "-d", "5", "--disk", "-f", config, "-t", "0"]
    _torchrun(2, [script, ref, "nfact_decomp", "-o", out] + common + ["-X", "1"], "dist_dropin_n2.log")
    _torchrun(1, [script, ref, "nfact_decomp", "-o", out1] + common + ["-X", "1"], "dist_dropin_n2.log")
    comp = os.path.join(out, "nfact_decomp", "components", "NMF", "decomp")
    comp1 = os.path.join(out1, "nfact_decomp", "components", "NMF", "decomp")
    avg = np.load(os.path.join(out, "nfact_decomp", "group_averages", "average_matrix2.npy"))
    assert np.array_equal(avg, np.load(os.path.join(out1, "nfact_decomp", "group_averages", "average_matrix2.npy")))
    W, G = np.load(os.path.join(comp, "W_NMF_dim5.npy")), np.load(os.path.join(comp, "G_NMF_dim5.npy"))
    W1, G1 = np.load(os.path.join(comp1, "W_NMF_dim5.npy")), np.load(os.path.join(comp1, "G_NMF_dim5.npy"))
    assert W.shape == (5, N) and G.shape == (M, 5)
    _cd_gate(G, W, G1, W1, avg, 5)
    # NMF-sso over two GPUs (the random-init runs are the replicas axis), planted study + config file
    from test_reference_mains_gpu import _study_planted
    from test_nmf_gpu import matched_correlation
    planted = tmp_path / "planted"
    planted.mkdir()
    Xp, p_list, p_seeds, p_config, out_sso = _study_planted(planted)
    out_sso1 = str(planted / "out1")
    os.makedirs(out_sso1)
    sso_args = ["-l", p_list, "-s", p_seeds, "-d", "4", "--disk", "-i", "6", "-f", p_config, "-t", "0"]
    _torchrun(2, [script, ref, "nfact_decomp", "-o", out_sso] + sso_args, "dist_dropin_n2.log")
    _torchrun(1, [script, ref, "nfact_decomp", "-o", out_sso1] + sso_args, "dist_dropin_n2.log")
    comp_sso = os.path.join(out_sso, "nfact_decomp", "components", "NMF", "decomp")
    comp_sso1 = os.path.join(out_sso1, "nfact_decomp", "components", "NMF", "decomp")
    Ws, Gs = np.load(os.path.join(comp_sso, "W_NMF_dim4.npy")), np.load(os.path.join(comp_sso, "G_NMF_dim4.npy"))
    Ws1 = np.load(os.path.join(comp_sso1, "W_NMF_dim4.npy"))
    assert Ws.shape == Ws1.shape == (4, Xp.shape[1]) and Gs.shape == (Xp.shape[0], 4)
    assert np.isfinite(Ws).all() and np.isfinite(Gs).all() and Ws.min() >= 0 and Gs.min() >= 0
    corr, live = matched_correlation(Ws, Ws1)
    assert live.all() and corr.min() >= 0.99, corr
    # nfact_dr: subjects round-robin over the two ranks
    import nfact_b200
    _torchrun(2, [script, ref, "nfact_dr", "-l", sub_list, "-s", seed_list, "-d", os.path.join(out, "nfact_decomp"),
                  "-o", out, "-a", "NMF"], "dist_dropin_n2.log", env={"NFB_TEST_DR_NPY": comp})
    group = {"white_components": W, "grey_components": G.astype(np.float64)}
    ranks_seen = set()
    for idx, s in enumerate(subs):
        sub_id = os.path.basename(s)
        ranks_seen.add(open(os.path.join(out, "nfact_dr", "NMF", f"rank_of_{sub_id}.txt")).read())
        Gd = np.load(os.path.join(out, "nfact_dr", "NMF", f"G_{sub_id}_dim5.npy"))
        Wd = np.load(os.path.join(out, "nfact_dr", "NMF", f"W_{sub_id}_dim5.npy"))
        x = nfact_b200.load_fdt_matrix(os.path.join(s, "fdt_matrix2.dot"))
        direct = nfact_b200.nmf_dual_regression(group, x)
        direct = nfact_b200.thresholding_components(3, os.path.join(s, "coords_for_fdt_matrix2"), [seed_list], direct)
        assert np.allclose(Gd, direct["grey_components"], rtol=1e-9, atol=1e-12)
        assert np.allclose(Wd, direct["white_components"], rtol=1e-9, atol=1e-12)
    assert ranks_seen == {"0", "1"}
