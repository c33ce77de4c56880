"""CPU: host logic of the seed-row sharded NNDSVD / random initialisation (nfact_b200/nndsvd.py, nmf.py).

The products with X are stood in by a float32 torch matmul on the CPU (on the GPU they are libnfact_b200's
split-precision GEMM); everything else -- the panel bookkeeping, TSQR, the all-reduces, the global sign flip,
the norms of the sharded side -- is the product's own code, run here at world size 1 and, over gloo, 2 and 3.
Reference: sklearn.decomposition._nmf._initialize_nmf(X, k, 'nndsvd', random_state=seed) as NFACT's default
``get_parameters`` asks for it (NFACT/decomp/decomposition/matrix_decomposition.py:100-107).
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_fro


def _problem(m, n, k, seed=5):
    rng = np.random.default_rng(seed)
    W = rng.gamma(0.5, 1.0, (m, k + 3))
    H = rng.gamma(0.5, 1.0, (k + 3, n))
    return (np.floor(40.0 * (W @ H)) * (rng.random((m, n)) < 0.5)).astype(np.float32)


def _cpu_operator(x_local):
    import torch
    xt = torch.from_numpy(x_local)

    def op(trans, q):
        return (xt.t() @ q) if trans else (xt @ q)
    return op


def _worker(rank, world, port, shape, k, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from nfact_b200.nndsvd import Comm, nndsvd_from_operator
    from nfact_b200.sharding import row_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X = _problem(*shape, k)
    r0, r1 = row_range(X.shape[0], rank, world)
    W, H = nndsvd_from_operator(_cpu_operator(np.ascontiguousarray(X[r0:r1])), r1 - r0, X.shape[1], k,
                                np.random.RandomState(0), Comm(rank, world))
    np.save(os.path.join(out_dir, f"W_{rank}.npy"), W)
    np.save(os.path.join(out_dir, f"H_{rank}.npy"), H)
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(90, 140), (150, 70)])      # m < n: sklearn works on X'; m > n: on X
def test_single_rank_matches_sklearn(shape):
    from sklearn.decomposition._nmf import _initialize_nmf
    from nfact_b200.nndsvd import nndsvd_from_operator
    k = 6
    X = _problem(*shape, k)
    Wr, Hr = _initialize_nmf(X, k, init="nndsvd", random_state=0)
    W, H = nndsvd_from_operator(_cpu_operator(X), X.shape[0], X.shape[1], k, np.random.RandomState(0))
    assert rel_fro(W, Wr) < 1e-3 and rel_fro(H, Hr) < 1e-3


@pytest.mark.parametrize("shape,world", [((90, 140), 2), ((150, 70), 2), ((61, 200), 3)])
def test_sharded_equals_unsharded(tmp_path, shape, world):
    import torch.multiprocessing as mp
    from nfact_b200.nndsvd import nndsvd_from_operator
    k = 6
    port = 29800 + (os.getpid() % 150) + world + shape[0] % 7
    mp.start_processes(_worker, args=(world, port, shape, k, str(tmp_path)), nprocs=world, join=True,
                       start_method="spawn")
    X = _problem(*shape, k)
    W1, H1 = nndsvd_from_operator(_cpu_operator(X), X.shape[0], X.shape[1], k, np.random.RandomState(0))
    W = np.concatenate([np.load(tmp_path / f"W_{r}.npy") for r in range(world)])
    assert W.shape == W1.shape
    assert rel_fro(W, W1) < 1e-3
    for r in range(world):                                   # H is replicated: the same on every rank
        H = np.load(tmp_path / f"H_{r}.npy")
        assert rel_fro(H, H1) < 1e-3
        assert np.array_equal(H, np.load(tmp_path / "H_0.npy"))


def test_random_init_is_sliced_by_global_row():
    """initialize_nmf(init='random') on a shard = the rows of the unsharded draw (same RandomState stream)."""
    import nfact_b200.nmf as nmf
    from nfact_b200.nndsvd import Comm

    class FakeComm(Comm):
        def __init__(self, rank, world, layout):
            super().__init__(rank, 1)
            self.world, self._layout = world, layout

        def row_layout(self, m_local, like):
            return self._layout

    class FakeX:
        def __init__(self, shape):
            self.shape = shape
            self.handle = type("H", (), {"device": 0})()

    m, n, k = 11, 7, 3
    W_full, H_full = nmf.initialize_nmf(FakeX((m, n)), 2.5, k, "random", 4)
    # world > 1 with an int seed: no broadcast needed, every rank draws the same stream
    import torch  # noqa: F401  (row_layout is overridden, torch is only imported by initialize_nmf)
    W_a, H_a = nmf.initialize_nmf(FakeX((5, n)), 2.5, k, "random", 4, FakeComm(0, 2, (0, m)))
    W_b, H_b = nmf.initialize_nmf(FakeX((6, n)), 2.5, k, "random", 4, FakeComm(1, 2, (5, m)))
    assert np.array_equal(np.concatenate([W_a, W_b]), W_full)
    assert np.array_equal(H_a, H_full) and np.array_equal(H_b, H_full)
