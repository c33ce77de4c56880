"""The reference's command-line tools on several GPUs (run under torchrun, one rank per GPU; spawned by
tests/test_dist_gpu.py):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_dropin_gpu.py \
        <reference root> nfact_decomp -l subs.txt -s seeds.txt -o out -d 5 --disk ...
Same as `-m nfact_b200.dropin <tool> ...` (rank 0 runs the reference's main, the other ranks serve its solves:
nfact_b200/multigpu.py), plus what a test box needs: the reference checkout on sys.path and stubs for the reference's
optional dependencies that are not installed here (nibabel, lz4, matplotlib, ...)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_dropin_cpu import _Stub  # noqa: E402


def main():
    ref_root, argv = sys.argv[1], sys.argv[2:]
    for name in ("lz4", "lz4.frame", "nibabel", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "seaborn",
                 "file_tree", "networkx", "fsl", "fsl.wrappers", "fsl_sub"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    sys.path.insert(0, ref_root)
    from nfact_b200 import dropin
    if os.environ.get("NFB_TEST_DR_NPY"):
        # nfact_dr: .npy stand-ins for the two NIfTI / GIFTI functions (nibabel is not installed), on every rank
        import numpy as np
        drf = importlib.import_module("NFACT.dual_reg.nfact_dr_functions")
        comp_dir = os.environ["NFB_TEST_DR_NPY"]

        def load(comp, avg, seeds, roi):
            k = [f for f in os.listdir(comp_dir) if f.startswith("W_NMF_dim")][0][len("W_NMF_dim"):-4]
            return {"white_components": np.load(os.path.join(comp_dir, f"W_NMF_dim{k}.npy")),
                    "grey_components": np.load(os.path.join(comp_dir, f"G_NMF_dim{k}.npy")).astype(np.float64)}

        def save(dr_results, out_dir, seeds, algo, dim, sub_id, subject_dir, roi, cifti):
            np.save(os.path.join(out_dir, algo, f"G_{sub_id}_dim{dim}.npy"), dr_results["grey_components"])
            np.save(os.path.join(out_dir, algo, f"W_{sub_id}_dim{dim}.npy"), dr_results["white_components"])
            with open(os.path.join(out_dir, algo, f"rank_of_{sub_id}.txt"), "w") as fh:
                fh.write(os.environ.get("RANK", "0"))
        drf.get_group_level_components = load
        drf.save_dual_regression_images = save
    try:
        dropin.main(argv)
    except SystemExit as e:
        if e.code not in (0, None):
            raise


if __name__ == "__main__":
    main()
