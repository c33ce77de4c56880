"""Golden vectors for the winner-takes-all maps, produced by the REFERENCE's own function.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_wta.py

Imports ``create_wta_map`` (NFACT/decomp/pipes/image_handling.py:216-245) unmodified -- nibabel, which the module
pulls in for the image writers, is replaced by an empty stand-in; the function itself is StandardScaler + argmax --
and applies it as ``winner_takes_all`` does (:186-187): white components [k, n] with axis 0, grey components [m, k]
with axis 1, float32 (nfact_decomp's dtype) and float64, with a dead component, a constant component, targets no
component reaches (all-zero column: scale 1, z = 0), duplicated maxima (first index wins) and several thresholds.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(f"{self.__name__}.{name}")
        setattr(self, name, child)
        return child


def main():
    sys.modules["nibabel"] = _Stub("nibabel")
    import_reference()
    from NFACT.decomp.pipes.image_handling import create_wta_map
    rng = np.random.default_rng(2718)
    m, n, k = 263, 411, 11
    out = {}
    for tag, dtype in (("f32", np.float32), ("f64", np.float64)):
        grey = (rng.gamma(0.4, 1.0, (m, k)) * (rng.random((m, k)) < 0.5)).astype(dtype)
        white = (rng.gamma(0.4, 1.0, (k, n)) * (rng.random((k, n)) < 0.6)).astype(dtype)
        grey[:, 3] = 0.0                    # dead component
        white[3, :] = 0.0
        grey[:, 5] = 0.75                   # constant component
        white[5, :] = 0.75
        white[:, 17] = 0.0                  # a target no component reaches
        white[:, 18] = 2.5                  # ... and one every component reaches equally
        grey[40, :] = 0.0
        white[7, 30:60] = white[2, 30:60]   # duplicated maxima: the first index wins
        grey[100:130, 8] = grey[100:130, 1]
        out[f"grey_{tag}"], out[f"white_{tag}"] = grey.copy(), white.copy()
        for z in (0.0, 1.0, 2.5):
            key = str(z).replace(".", "p")
            out[f"white_wta_{key}_{tag}"] = create_wta_map(white.copy(), 0, z)
            out[f"grey_wta_{key}_{tag}"] = create_wta_map(grey.copy(), 1, z)
    np.savez_compressed(os.path.join(HERE, "golden_wta.npz"), **out)
    print("wrote golden_wta.npz:", {k_: (v.shape, v.dtype) for k_, v in out.items()})
    for key in ("white_wta_1p0_f32", "grey_wta_1p0_f32"):
        print(key, np.bincount(out[key].ravel(), minlength=k + 1))


if __name__ == "__main__":
    main()
