"""Golden vectors for the NMF-sso clustering glue, produced by the REFERENCE's own functions
(NFACT/decomp/decomposition/sso/sso_functions.py: compute_similairty_matrix, sim2dis, clustering_components,
idx2centrotype) on a seeded synthetic stack of components.  Run in the build container only:

    python tests/golden/make_golden_sso.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402


def main():
    import_reference()
    from NFACT.decomp.decomposition.sso import sso_functions as ref
    rng = np.random.default_rng(2024)
    runs, k, n = 6, 7, 900
    # every run finds (noisy, permuted, rescaled) versions of the same k sparse components
    base = rng.gamma(0.4, 1.0, (k, n)) * (rng.random((k, n)) < 0.25)
    stack = []
    for _ in range(runs):
        noisy = base * rng.uniform(0.8, 1.2, (k, 1)) + 0.05 * rng.gamma(0.4, 1.0, (k, n)) * (rng.random((k, n)) < 0.1)
        stack.append(noisy[rng.permutation(k)])
    comps = np.vstack(stack).astype(np.float32)
    sim = ref.compute_similairty_matrix(comps)
    dis = ref.sim2dis(sim)
    part = ref.clustering_components(dis, k)
    cent = ref.idx2centrotype(sim, part)
    np.savez_compressed(os.path.join(HERE, "golden_sso.npz"), components=comps, sim=sim, dis=dis,
                        partition=np.asarray(part), centrotypes=np.asarray(cent), dim=np.int64(k))
    print("sim", sim.shape, "clusters", int(np.max(part)), "centrotypes", cent)


if __name__ == "__main__":
    main()
