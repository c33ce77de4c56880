"""Generate the golden vectors under tests/golden/ by running the REFERENCE itself.

Run in the build container only (needs /root/reference plus the installed sklearn/scipy):

    python tests/golden/make_golden.py

The reference's own functions are imported unmodified from /root/reference (optional
dependencies that are not installed here -- lz4, nibabel, matplotlib, seaborn, file_tree --
are replaced by empty modules; none of them is touched on the NMF path).  Inputs are our own
seeded synthetic ``fdt_matrix2.dot`` files that keep the properties of the reference's
test data (NFACT/testing/generate_connectivity_matrix.py: integer counts 50-499, ~half of the
coordinates duplicated, targets numbered so that the first columns are all zero, last line =
matrix shape).  Nothing under tests/, smoke() or bench.py reads /root/reference at run time:
they read the files this script commits.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def import_reference():
    for name in ("lz4", "lz4.frame", "nibabel", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.colors", "seaborn", "file_tree", "networkx"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    sys.path.insert(0, REF)


def write_subject(folder, rng, n_seeds, n_targets, first_target, n_conn, waytotal, integer=True):
    """Our own seeded .dot writer (1-based ``row col value`` lines, last line = shape)."""
    os.makedirs(folder, exist_ok=True)
    seeds = rng.integers(1, n_seeds + 1, size=n_conn)
    targets = rng.integers(first_target, n_targets + 1, size=n_conn)
    if integer:
        vals = rng.integers(50, 500, size=n_conn).astype(np.float64)
    else:
        # probtrackx --pd style: non-integer, and made duplicate-free below
        vals = rng.uniform(0.5, 900.0, size=n_conn)
        _, keep = np.unique(seeds * (n_targets + 1) + targets, return_index=True)
        seeds, targets, vals = seeds[keep], targets[keep], vals[keep]
    order = np.lexsort((targets, seeds))
    seeds, targets, vals = seeds[order], targets[order], vals[order]
    with open(os.path.join(folder, "fdt_matrix2.dot"), "w") as fh:
        for r, c, v in zip(seeds, targets, vals):
            fh.write(f"{r} {c} {int(v)}\n" if integer else f"  {r}  {c}  {v:.6g}\n")
        fh.write(f"{n_seeds} {n_targets} {int(rng.integers(50, 500))}\n")
    with open(os.path.join(folder, "waytotal"), "w") as fh:
        fh.write(str(waytotal))


def main():
    import_reference()
    from NFACT.base.matrix_handling import load_fdt_matrix, normalise_components, thresholding
    from NFACT.decomp.decomposition.matrix_handling import avg_fdt
    from NFACT.decomp.decomposition.matrix_decomposition import get_parameters
    from NFACT.decomp.decomposition.sso.nmfsso import nmf_decomp
    from NFACT.dual_reg.dual_regression.dual_regression_methods import nmf_dual_regression
    import sklearn.decomposition._nmf as sknmf
    from scipy.optimize import nnls as scipy_nnls

    rng = np.random.default_rng(20261019)
    data_dir = os.path.join(HERE, "data")
    ways = [15000000, 12345678, 21987654]
    for s in range(3):
        write_subject(os.path.join(data_dir, f"sub-{s + 1}"), rng, 99, 199, 101, 20000, ways[s])
    write_subject(os.path.join(data_dir, "sub-pd"), rng, 61, 83, 1, 2500, 9876543, integer=False)
    subs = [os.path.join(data_dir, f"sub-{s + 1}", "fdt_matrix2.dot") for s in range(3)]

    out = {}
    # ---- ingest (a1-a5) -------------------------------------------------------------
    for s, path in enumerate(subs):
        out[f"ingest_single_{s + 1}"] = load_fdt_matrix(path)
    out["ingest_single_pd"] = load_fdt_matrix(os.path.join(data_dir, "sub-pd", "fdt_matrix2.dot"))
    out["ingest_group"] = avg_fdt(subs)
    out["ingest_group2"] = avg_fdt(subs[:2])
    X = out["ingest_group"]

    # ---- hyper-parameters (a6, a9) --------------------------------------------------
    params = get_parameters(None, "nmf", 10)
    out["params_keys"] = np.array(sorted(params.keys()))
    out["params_repr"] = np.array(repr(sorted(params.items(), key=lambda kv: kv[0])))
    est = sknmf.NMF(**params)
    est._check_params(X)
    out["regularization"] = np.array(est._compute_regularization(X), dtype=np.float64)

    # ---- init (a8) ------------------------------------------------------------------
    W0r, H0r = sknmf._initialize_nmf(X, 10, init="random", random_state=7)
    out["init_random_W"], out["init_random_H"] = W0r.copy(), H0r.copy()
    W0, H0 = sknmf._initialize_nmf(X, 10, init="nndsvd", random_state=0)
    # NB: in the transposed branch of _randomized_svd H0 comes back Fortran-ordered, so H0.T is
    # C-contiguous and np.ascontiguousarray(H0.T) would alias it -- normalise to C order copies.
    W0, H0 = np.array(W0, order="C", copy=True), np.array(H0, order="C", copy=True)
    out["init_nndsvd_W"], out["init_nndsvd_H"] = W0.copy(), H0.copy()

    # ---- MU trajectory (a11, a12) from the fixed nndsvd start -------------------------
    regs = est._compute_regularization(X)
    l1W, l1H, l2W, l2H = regs
    W, H = W0.copy(), H0.copy()
    out["mu_err0"] = np.array(sknmf._beta_divergence(X, W, H, 2, square_root=True))
    keep = (1, 2, 3, 5, 10, 20, 40)
    for it in range(1, 41):
        W, *_ = sknmf._multiplicative_update_w(X, W, H, 2, l1W, l2W, 1.0)
        H = sknmf._multiplicative_update_h(X, W, H, 2, l1H, l2H, 1.0)
        if it in keep:
            out[f"mu_W_{it}"], out[f"mu_H_{it}"] = W.copy(), H.copy()
            out[f"mu_err_{it}"] = np.array(sknmf._beta_divergence(X, W, H, 2, square_root=True))
    # MU with L2 + L1 on a random (dense, strictly positive) start
    regs2 = (0.7, 0.4, 1.3, 0.9)
    W, H = W0r.copy(), H0r.copy()
    for it in range(1, 6):
        W, *_ = sknmf._multiplicative_update_w(X, W, H, 2, regs2[0], regs2[2], 1.0)
        H = sknmf._multiplicative_update_h(X, W, H, 2, regs2[1], regs2[3], 1.0)
    out["mu_l2_W_5"], out["mu_l2_H_5"] = W, H
    out["mu_l2_regs"] = np.array(regs2)
    # full MU solve through the reference's own nmf_decomp
    pm = dict(params, solver="mu", max_iter=200)
    res = nmf_decomp(pm, X, W=W0.copy(), H=H0.copy())
    out["mu_fit_W"], out["mu_fit_H"] = res["grey_components"], res["white_components"]
    est_mu = sknmf.NMF(**dict(pm, init="custom"))
    est_mu.fit_transform(X, W=W0.copy(), H=H0.copy())
    out["mu_fit_n_iter"] = np.array(est_mu.n_iter_)

    # ---- CD (a10) ----------------------------------------------------------------------
    W, Ht = W0.copy(), np.array(H0.T, order="C", copy=True)
    viol = []
    rs = np.random.RandomState(0)
    for it in range(1, 4):
        v = sknmf._update_coordinate_descent(X, W, Ht, l1W, l2W, False, rs)
        v += sknmf._update_coordinate_descent(X.T, Ht, W, l1H, l2H, False, rs)
        viol.append(v)
        out[f"cd_W_{it}"], out[f"cd_H_{it}"] = W.copy(), Ht.T.copy()
    out["cd_violation"] = np.array(viol, dtype=np.float64)
    res = nmf_decomp(dict(params), X, W=W0.copy(), H=H0.copy())   # NFACT defaults: cd, tol 1e-4, 200 iters
    out["cd_fit_W"], out["cd_fit_H"] = res["grey_components"], res["white_components"]
    est_cd = sknmf.NMF(**dict(params, init="custom"))
    est_cd.fit_transform(X, W=W0.copy(), H=H0.copy())
    out["cd_fit_n_iter"] = np.array(est_cd.n_iter_)
    out["cd_fit_err"] = np.array(est_cd.reconstruction_err_)
    # a second CD case with weaker regularisation and l2
    p2 = dict(params, alpha_W=0.001, l1_ratio=0.5, max_iter=60)
    res2 = nmf_decomp(p2, X, W=W0r.copy(), H=H0r.copy())
    out["cd2_fit_W"], out["cd2_fit_H"] = res2["grey_components"], res2["white_components"]

    # ---- post-ops (a14) ---------------------------------------------------------------------
    norm = normalise_components(out["cd_fit_W"].copy(), out["cd_fit_H"].copy())
    out["norm_grey"], out["norm_white"] = norm["grey_matter"], norm["white_matter"]
    out["thresh_white_3"] = thresholding(out["cd_fit_H"].copy(), 3)
    out["thresh_white_1"] = thresholding(out["mu_fit_H"].copy(), 1)

    # ---- NNLS dual regression (a15) -------------------------------------------------------
    Xs = out["ingest_single_1"]
    comps = {"grey_components": out["cd_fit_W"].astype(np.float64),
             "white_components": out["cd_fit_H"].astype(np.float64)}
    dr = nmf_dual_regression(comps, Xs, 1)
    out["dr_grey"], out["dr_white"] = dr["grey_components"], dr["white_components"]
    comps_mu = {"grey_components": out["mu_fit_W"], "white_components": out["mu_fit_H"]}
    dr2 = nmf_dual_regression(comps_mu, out["ingest_single_2"], 1)
    out["dr2_grey"], out["dr2_white"] = dr2["grey_components"], dr2["white_components"]
    # raw scipy nnls known answers (scipy/optimize/_nnls.py docstring + random cases)
    rs = np.random.RandomState(3)
    A = np.abs(rs.standard_normal((40, 12)))
    A[:, 5] = A[:, 2] * 0.5 + A[:, 7] * 0.5 + 1e-3 * rs.standard_normal(40)   # near-collinear
    B = rs.standard_normal((40, 9)) + 0.3
    out["nnls_A"], out["nnls_B"] = A, B
    out["nnls_X"] = np.array([scipy_nnls(A, B[:, j])[0] for j in range(B.shape[1])])

    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_small.npz"), "with", len(out), "arrays")
    import sklearn, scipy
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as fh:
        fh.write(f"numpy {np.__version__}\nscipy {scipy.__version__}\nscikit-learn {sklearn.__version__}\n")


if __name__ == "__main__":
    main()
