"""Golden vectors for the component post-processing, produced by the REFERENCE's own functions.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_postops.py

Imports NFACT/base/matrix_handling.py unmodified (``thresholding`` :218, ``threshold_grey_components`` :184,
``thresholding_components`` :241, ``normalise_components`` :45) and applies them to seeded synthetic component
maps: float32 (what nfact_decomp produces) and float64 (what nfact_dr produces), two seeds with interleaved rows
plus rows that belong to no listed seed, a constant component, an all-zero component.  The coordinate file the
grey thresholding reads is committed beside the vectors (tests/golden/data/postops/coords_for_fdt_matrix2).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    import_reference()
    from NFACT.base.matrix_handling import (normalise_components, threshold_grey_components, thresholding,
                                            thresholding_components)
    rng = np.random.default_rng(414)
    m, n, k = 301, 517, 9
    folder = os.path.join(HERE, "data", "postops")
    os.makedirs(folder, exist_ok=True)
    seeds_id = rng.integers(0, 3, size=m)            # ids 0, 1 are the two listed seeds; 2 belongs to neither
    coords = np.stack([rng.integers(0, 90, m), rng.integers(0, 90, m), rng.integers(0, 90, m), seeds_id,
                       np.arange(m)], axis=1)
    coord_path = os.path.join(folder, "coords_for_fdt_matrix2")
    np.savetxt(coord_path, coords, fmt="%d")
    seeds = ["left.surf.gii", "right.surf.gii"]

    def maps(dtype):
        grey = (rng.gamma(0.4, 1.0, (m, k)) * (rng.random((m, k)) < 0.5)).astype(dtype)
        white = (rng.gamma(0.4, 1.0, (k, n)) * (rng.random((k, n)) < 0.6)).astype(dtype)
        grey[:, 3] = 0.0                               # dead component
        white[3, :] = 0.0
        grey[:, 5] = 0.75                              # constant component (StandardScaler: scale -> 1)
        white[5, :] = 0.75
        return grey, white

    out = {"seeds_id": seeds_id.astype(np.int32), "n_seeds": np.int64(len(seeds))}
    for tag, dtype in (("f32", np.float32), ("f64", np.float64)):
        grey, white = maps(dtype)
        out[f"grey_{tag}"], out[f"white_{tag}"] = grey.copy(), white.copy()
        for z in (1, 3):
            out[f"white_thr{z}_{tag}"] = thresholding(white.copy(), z)
            out[f"grey_thr{z}_{tag}"] = threshold_grey_components(grey.copy(), coord_path, seeds, z)
        comps = thresholding_components(2, coord_path, seeds, {"grey_components": grey.copy(),
                                                               "white_components": white.copy()})
        out[f"grey_thr2_{tag}"], out[f"white_thr2_{tag}"] = comps["grey_components"], comps["white_components"]
        norm = normalise_components(comps["grey_components"].copy(), comps["white_components"].copy())
        out[f"grey_thr2_norm_{tag}"], out[f"white_thr2_norm_{tag}"] = norm["grey_matter"], norm["white_matter"]
        norm = normalise_components(grey.copy(), white.copy())
        out[f"grey_norm_{tag}"], out[f"white_norm_{tag}"] = norm["grey_matter"], norm["white_matter"]
    np.savez_compressed(os.path.join(HERE, "golden_postops.npz"), **out)
    print("wrote golden_postops.npz:", {k_: v.shape for k_, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
