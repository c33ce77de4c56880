"""GPU: the reference's OWN command-line mains, unmodified, on the B200 path through ``nfact_b200.dropin``.

``nfact_decomp_main`` (NFACT/decomp/__main__.py:42-211) and ``nfact_dr_main`` (NFACT/dual_reg/__main__.py:27-96)
are imported from a reference checkout -- ``$NFB_REFERENCE_PATH``, ``/root/reference`` (build container) or
``baseline/_ref`` (the pip-installed copy that travels to the GPU box; git-ignored, never part of the repo's
history) -- and run on the golden subjects after ``dropin.install()``.  The reference's optional dependencies that
are not installed (nibabel, lz4, matplotlib, ...) are stubbed as in tests/test_dropin_cpu.py; image I/O therefore
goes through the reference's own ``.npy`` writer (``--disk`` -> ``save_components_to_disk``,
decomp/pipes/image_handling.py:78-110) for ``nfact_decomp``, and through ``.npy`` stand-ins for the two NIfTI/GIFTI
functions of ``nfact_dr`` (component loader, image writer).  Every output file is checked against the same steps
called directly on ``nfact_b200``.  Skipped when no reference checkout is present.
"""
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

from conftest import GOLDEN_DIR, ROOT
from test_dropin_cpu import _Stub
from test_nmf_gpu import matched_correlation

pytestmark = pytest.mark.gpu

M, N = 99, 199            # the golden subjects' matrix


def _reference_root():
    for cand in (os.environ.get("NFB_REFERENCE_PATH"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "NFACT")):
            return cand
    return None


@pytest.fixture
def reference_tools():
    root = _reference_root()
    if root is None:
        pytest.skip("no reference checkout (NFB_REFERENCE_PATH, /root/reference, baseline/_ref)")
    from nfact_b200 import dropin
    added = []
    for name in ("lz4", "lz4.frame", "nibabel", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "seaborn",
                 "file_tree", "networkx", "fsl", "fsl.wrappers", "fsl_sub"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)
            added.append(name)
    sys.path.insert(0, root)
    argv = list(sys.argv)
    yield dropin
    sys.argv = argv
    dropin.uninstall()
    sys.path.remove(root)
    for name in added:
        sys.modules.pop(name, None)
    for name in [n for n in sys.modules if n == "NFACT" or n.startswith("NFACT.")]:
        del sys.modules[name]


def _study(tmp_path):
    """Three probtrackx-like subject folders from the golden fixtures + the list files the tools take."""
    subs = []
    for s in (1, 2, 3):
        d = tmp_path / f"sub_{s}"          # get_subject_id (nfact_dr_functions.py:484-508) looks for sub_<id>
        d.mkdir()
        for f in ("fdt_matrix2.dot", "waytotal"):
            shutil.copy(os.path.join(GOLDEN_DIR, "data", f"sub-{s}", f), d)
        coords = np.zeros((M, 5), dtype=int)
        coords[:, 3] = 0                                   # one seed: every row belongs to seed 0
        coords[:, 4] = np.arange(M)
        np.savetxt(d / "coords_for_fdt_matrix2", coords, fmt="%d")
        np.savetxt(d / "tract_space_coords_for_fdt_matrix2", np.zeros((N, 3), dtype=int), fmt="%d")
        (d / "lookup_tractspace_fdt_matrix2.nii.gz").write_bytes(b"x")
        subs.append(str(d))
    (tmp_path / "subs.txt").write_text("\n".join(subs) + "\n")
    seed = tmp_path / "seed.nii.gz"
    seed.write_bytes(b"x")
    (tmp_path / "seeds.txt").write_text(str(seed) + "\n")
    out = tmp_path / "out"
    out.mkdir()
    return subs, str(tmp_path / "subs.txt"), str(tmp_path / "seeds.txt"), str(out)


def _run_tool(dropin, argv):
    try:
        dropin.main(argv)
    except SystemExit as e:            # the reference ends a command-line run with exit(0)
        assert e.code in (0, None), f"tool exited with {e.code}"


def _cd_gate(W, H, Wd, Hd, X, k):
    from oracle import nfact_oracle as orc
    regs = orc.compute_regularization(X.shape[0], X.shape[1], 0.1, "same", 1.0)
    obj, obj_d = orc.nmf_objective(X, W, H, *regs), orc.nmf_objective(X, Wd, Hd, *regs)
    assert abs(obj - obj_d) <= 5e-3 * obj_d
    corr, live = matched_correlation(H, Hd)
    assert corr[live].min() >= 0.99


def _config(tmp_path, **overrides):
    """nfact_config-style hyper-parameter file (-f): sklearn NMF's signature with NFACT's defaults
    (matrix_decomposition.py:100-107) and the given overrides."""
    import json
    params = {"init": "nndsvd", "solver": "cd", "beta_loss": "frobenius", "tol": 1e-4, "max_iter": 200,
              "random_state": None, "alpha_W": 0.1, "alpha_H": "same", "l1_ratio": 1, "verbose": 0, "shuffle": False}
    params.update(overrides)
    path = tmp_path / "config.json"
    path.write_text(json.dumps({"nmf": params, "ica": {}}))
    return str(path), params


def test_nfact_decomp_main_single_solve(reference_tools, tmp_path):
    """nfact_decomp -l subs -s seeds -o out -d 5 --disk -X 1 -f config (no sso): ingest -> average_matrix2.npy ->
    NNDSVD + CD solve with NFACT's default penalties -> thresholding -> W / G .npy files.  random_state = 0 in the
    config file makes the randomised SVD of the initialisation repeatable, so the files must equal the same steps
    called directly."""
    import nfact_b200
    from oracle import nfact_oracle as orc
    dropin = reference_tools
    subs, sub_list, seed_list, out = _study(tmp_path)
    config, params = _config(tmp_path, random_state=0)
    _run_tool(dropin, ["nfact_decomp", "-l", sub_list, "-s", seed_list, "-o", out, "-d", "5", "--disk", "-X", "1",
                       "-f", config])
    mod = sys.modules["NFACT.decomp.__main__"]
    assert mod.process_fdt_matrix2 is nfact_b200.process_fdt_matrix2          # the run went through the rebound seams
    avg = np.load(os.path.join(out, "nfact_decomp", "group_averages", "average_matrix2.npy"))
    files = [os.path.join(s, "fdt_matrix2.dot") for s in subs]
    assert avg.dtype == np.float32 and np.array_equal(avg, orc.avg_fdt(files))           # bit-exact ingest
    comp_dir = os.path.join(out, "nfact_decomp", "components", "NMF", "decomp")
    W = np.load(os.path.join(comp_dir, "W_NMF_dim5.npy"))       # white components [k, n]
    G = np.load(os.path.join(comp_dir, "G_NMF_dim5.npy"))       # grey components [m, k]
    assert W.shape == (5, N) and G.shape == (M, 5)
    direct = nfact_b200.nmf_decomp(nfact_b200.get_parameters(dict(params), "nmf", 5), avg)
    # un-thresholded solve against sklearn itself from the same (sklearn-made) NNDSVD start: the CD gate
    from sklearn.decomposition import NMF
    from sklearn.decomposition._nmf import _initialize_nmf
    W0, H0 = _initialize_nmf(avg, 5, init="nndsvd", random_state=0)
    sk = NMF(**dict(params, n_components=5, init="custom"))
    Wsk = sk.fit_transform(avg, W=W0.copy(), H=H0.copy())
    _cd_gate(direct["grey_components"], direct["white_components"], Wsk, sk.components_, avg, 5)
    coords = os.path.join(out, "nfact_decomp", "group_averages", "coords_for_fdt_matrix2")
    direct = nfact_b200.thresholding_components(3, coords, [seed_list], direct)
    assert np.allclose(G, direct["grey_components"], rtol=1e-6, atol=0)
    assert np.allclose(W, direct["white_components"], rtol=1e-6, atol=0)
    assert (W == 0).mean() > 0.3                                # the z = 3 threshold did its work


def _study_planted(tmp_path, k=4, seed=12):
    """One subject whose matrix has k planted sparse components (integer streamline counts, waytotal 1e8 -> scale
    1): the NMF-sso flow needs runs that agree, which the 99 x 199 fixtures do not give (with them the reference's
    own clustering fails on all-zero thresholded components).  Returns (X, subject list, seed list, config, out)."""
    rng = np.random.default_rng(seed)
    m, n = 120, 210
    Wp = rng.gamma(0.4, 1.0, (m, k)) * (rng.random((m, k)) < 0.4)
    Hp = rng.gamma(0.4, 1.0, (k, n)) * (rng.random((k, n)) < 0.4)
    X = np.floor(Wp @ Hp * 100 + 3 * rng.random((m, n)))
    d = tmp_path / "sub_1"
    d.mkdir()
    rows, cols = np.nonzero(X)
    with open(d / "fdt_matrix2.dot", "wb") as fh:
        np.savetxt(fh, np.column_stack([rows + 1, cols + 1, X[rows, cols].astype(np.int64)]), fmt="%d %d %d")
        fh.write(f"{m} {n} 0\n".encode())
    (d / "waytotal").write_text("100000000\n")
    coords = np.zeros((m, 5), dtype=int)
    coords[:, 4] = np.arange(m)
    np.savetxt(d / "coords_for_fdt_matrix2", coords, fmt="%d")
    np.savetxt(d / "tract_space_coords_for_fdt_matrix2", np.zeros((n, 3), dtype=int), fmt="%d")
    (d / "lookup_tractspace_fdt_matrix2.nii.gz").write_bytes(b"x")
    (tmp_path / "subs.txt").write_text(str(d) + "\n")
    seed_img = tmp_path / "seed.nii.gz"
    seed_img.write_bytes(b"x")
    (tmp_path / "seeds.txt").write_text(str(seed_img) + "\n")
    config, _ = _config(tmp_path, alpha_W=1e-4, max_iter=150)      # a small penalty: the planted factors survive
    out = tmp_path / "out"
    out.mkdir(exist_ok=True)
    return X.astype(np.float32), str(tmp_path / "subs.txt"), str(tmp_path / "seeds.txt"), config, str(out)


def reference_sso_flow(X, params, iterations, dim):
    """The reference's NMF-sso flow with its own functions on the CPU (sklearn), run serially: NMFsso's parallel
    branch (nmfsso.py:120-138: random init, thresholding(H, 3) per run) -- its single-core branch cannot run
    (`range(colours)`, nmfsso.py:207) and joblib workers cannot see the test's stub modules --, then nmf_sso's own
    lines (nmfsso.py:346-358)."""
    sso = importlib.import_module("NFACT.decomp.decomposition.sso.nmfsso")
    fn = importlib.import_module("NFACT.decomp.decomposition.sso.sso_functions")
    bmh = importlib.import_module("NFACT.base.matrix_handling")
    run_params = dict(params, init="random", random_state=None)
    white, grey = [], []
    for _ in range(iterations):
        res = sso.nmf_decomp(run_params, X)
        grey.append(res["grey_components"])
        white.append(bmh.thresholding(res["white_components"], 3))
    w_components, g_components = np.vstack(white), np.hstack(grey)
    sim = fn.compute_similairty_matrix(w_components)
    partitions = fn.clustering_components(fn.sim2dis(sim), dim)
    centroids = fn.idx2centrotype(sim, partitions)
    final = dict(params, random_state=None, init="custom", n_components=centroids.shape[0])
    return sso.nmf_decomp(final, X, W=np.ascontiguousarray(g_components[:, centroids]),
                          H=np.ascontiguousarray(w_components[centroids, :]))


def test_nfact_decomp_main_default_sso_flow(reference_tools, tmp_path):
    """NFACT's default flow through the reference's main: NMF-sso (6 random-init runs) -> similarity -> clustering ->
    centrotypes -> final custom-init solve, hyper-parameters from a -f config file.  Compared with the reference's
    own functions run on the CPU BEFORE the drop-in is installed: same number of components, Hungarian-matched
    component correlation.  (The reference's CLI cannot run this configuration itself: `range(colours)` TypeError at
    nmfsso.py:207 with n_cores = 1; the drop-in keeps the intended behaviour.)"""
    import json
    dropin = reference_tools
    X, sub_list, seed_list, config, out = _study_planted(tmp_path)
    params = dict(json.load(open(config))["nmf"], n_components=4)
    ref = reference_sso_flow(X, params, 6, 4)                              # the reference's arithmetic, untouched
    _run_tool(dropin, ["nfact_decomp", "-l", sub_list, "-s", seed_list, "-o", out, "-d", "4", "--disk", "-i", "6",
                       "-f", config, "-t", "0"])
    comp_dir = os.path.join(out, "nfact_decomp", "components", "NMF", "decomp")
    W = np.load(os.path.join(comp_dir, "W_NMF_dim4.npy"))
    G = np.load(os.path.join(comp_dir, "G_NMF_dim4.npy"))
    kf = W.shape[0]
    assert kf == ref["white_components"].shape[0] == 4 and W.shape == (kf, X.shape[1]) and G.shape == (X.shape[0], kf)
    assert np.isfinite(W).all() and np.isfinite(G).all() and W.min() >= 0 and G.min() >= 0
    corr, live = matched_correlation(W, ref["white_components"])
    assert live.all() and corr.min() >= 0.99, corr
    corr_g, live_g = matched_correlation(G.T, ref["grey_components"].T)
    assert live_g.all() and corr_g.min() >= 0.99, corr_g
    rel = np.linalg.norm(X - G @ W) / np.linalg.norm(X)
    rel_ref = np.linalg.norm(X - ref["grey_components"] @ ref["white_components"]) / np.linalg.norm(X)
    assert abs(rel - rel_ref) < 0.01, (rel, rel_ref)


def test_nfact_dr_main(reference_tools, tmp_path, monkeypatch):
    """nfact_dr -l subs -s seeds -d out/nfact_decomp -o out -a NMF: the rebound subject loop -> per-subject files,
    equal to nmf_dual_regression + thresholding called directly."""
    import nfact_b200
    dropin = reference_tools
    subs, sub_list, seed_list, out = _study(tmp_path)
    _run_tool(dropin, ["nfact_decomp", "-l", sub_list, "-s", seed_list, "-o", out, "-d", "5", "--disk", "-X", "1"])
    comp_dir = os.path.join(out, "nfact_decomp", "components", "NMF", "decomp")
    group = {"white_components": np.load(os.path.join(comp_dir, "W_NMF_dim5.npy")),
             "grey_components": np.load(os.path.join(comp_dir, "G_NMF_dim5.npy")).astype(np.float64)}
    drf = importlib.import_module("NFACT.dual_reg.nfact_dr_functions")
    # the two NIfTI / GIFTI functions of nfact_dr, as .npy stand-ins (nibabel is not installed)
    monkeypatch.setattr(drf, "get_group_level_components", lambda comp, avg, seeds, roi: dict(group))
    saved = {}

    def save_npy(dr_results, out_dir, seeds, algo, dim, sub_id, subject_dir, roi, cifti):
        np.save(os.path.join(out_dir, algo, f"G_{sub_id}_dim{dim}.npy"), dr_results["grey_components"])
        np.save(os.path.join(out_dir, algo, f"W_{sub_id}_dim{dim}.npy"), dr_results["white_components"])
        saved[sub_id] = subject_dir
    monkeypatch.setattr(drf, "save_dual_regression_images", save_npy)
    _run_tool(dropin, ["nfact_dr", "-l", sub_list, "-s", seed_list, "-d", os.path.join(out, "nfact_decomp"),
                       "-o", out, "-a", "NMF"])
    assert len(saved) == 3
    for sub_id, subject_dir in saved.items():
        Gs = np.load(os.path.join(out, "nfact_dr", "NMF", f"G_{sub_id}_dim5.npy"))
        Ws = np.load(os.path.join(out, "nfact_dr", "NMF", f"W_{sub_id}_dim5.npy"))
        assert Gs.shape == (M, 5) and Ws.shape == (5, N) and Gs.dtype == np.float64
        x = nfact_b200.load_fdt_matrix(os.path.join(subject_dir, "fdt_matrix2.dot"))
        direct = nfact_b200.nmf_dual_regression(group, x)
        direct = nfact_b200.thresholding_components(3, os.path.join(subject_dir, "coords_for_fdt_matrix2"), [seed_list],
                                                    direct)
        assert np.allclose(Gs, direct["grey_components"], rtol=1e-9, atol=1e-12)
        assert np.allclose(Ws, direct["white_components"], rtol=1e-9, atol=1e-12)
