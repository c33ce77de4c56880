"""CPU, world_size 2 over gloo: the row-sharded NMF iteration (what each GPU rank computes plus the two
all-reduces) reproduces the unsharded solve.  The per-rank arithmetic here is the oracle's; the same
decomposition is what libnfact_b200 runs with NCCL (csrc/solver.cu gemm_wtx / compute_gram)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def _worker(rank, world, port, solver, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from nfact_b200.sharding import row_range
    from oracle import nfact_oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(4)
    m, n, k = 61, 47, 5
    X = (rng.gamma(0.5, 1.0, (m, n)) * (rng.random((m, n)) < 0.5) * 100).astype(np.float32)
    W0, H0 = orc.init_random(X, k, 3)
    regs = orc.compute_regularization(m, n, 0.01, "same", 0.5)
    l1W, l1H, l2W, l2H = regs
    r0, r1 = row_range(m, rank, world)
    Xs, W, H = X[r0:r1], W0[r0:r1].copy(), H0.copy()

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    for _ in range(4):
        if solver == "mu":
            W = orc.mu_update_w(Xs, W, H, l1W, l2W)                      # local: needs only HH'
            num = allreduce(W.T @ Xs)                                    # k x n numerator
            gram = allreduce(W.T @ W)                                    # k x k Gram
            den = gram @ H
            if l1H > 0:
                den = den + l1H
            if l2H > 0:
                den = den + l2H * H
            den[den == 0] = orc.EPSILON
            H = H * (num / den)
        else:
            Ht = np.ascontiguousarray(H.T)
            orc.cd_half_update(Xs, W, Ht, l1W, l2W)                      # local sweep over the W shard
            num = allreduce(W.T @ Xs).T                                  # [n, k]
            gram = allreduce(W.T @ W)
            lib = orc._clib()
            gram = np.ascontiguousarray(gram, dtype=np.float32)
            if l2H != 0.0:
                gram.flat[:: k + 1] += l2H
            num = np.ascontiguousarray(num - l1H, dtype=np.float32)
            lib.oracle_cd_sweep_f32(Ht.ctypes.data, gram.ctypes.data, num.ctypes.data, n, k)
            H = np.ascontiguousarray(Ht.T)
    np.save(os.path.join(out_dir, f"W_{rank}.npy"), W)
    if rank == 0:
        np.save(os.path.join(out_dir, "H.npy"), H)
        np.save(os.path.join(out_dir, "X.npy"), X)
    dist.destroy_process_group()


@pytest.mark.parametrize("solver", ["mu", "cd"])
def test_row_sharded_iteration_matches_unsharded(tmp_path, solver):
    import torch.multiprocessing as mp
    from oracle import nfact_oracle as orc
    port = 29600 + (os.getpid() % 200) + (0 if solver == "mu" else 1)
    mp.start_processes(_worker, args=(2, port, solver, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    X = np.load(tmp_path / "X.npy")
    W = np.concatenate([np.load(tmp_path / "W_0.npy"), np.load(tmp_path / "W_1.npy")])
    H = np.load(tmp_path / "H.npy")
    m, n, k = X.shape[0], X.shape[1], H.shape[0]
    W0, H0 = orc.init_random(X, k, 3)
    regs = orc.compute_regularization(m, n, 0.01, "same", 0.5)
    fit = orc.fit_mu if solver == "mu" else orc.fit_cd
    Wr, Hr, _ = fit(X, W0, H0, *regs, tol=0, max_iter=4)
    assert np.linalg.norm(W - Wr) / np.linalg.norm(Wr) < 1e-5
    assert np.linalg.norm(H - Hr) / np.linalg.norm(Hr) < 1e-5


def test_row_range_is_a_partition():
    from nfact_b200.sharding import row_range
    for m in (1, 7, 59412):
        for world in (1, 2, 3, 8):
            edges = [row_range(m, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == m
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        row_range(10, 2, 2)


def test_dual_regression_subjects_are_partitioned_over_ranks(monkeypatch, tmp_path):
    """run_locally(rank, world): rank r takes subjects r, r + world, ... (no collective); every subject is done by
    exactly one rank, in list order, and `save` sees each of them once.  The device work is replaced by a recorder:
    this is the host-side distribution logic only."""
    import nfact_b200.dual_regression as drmod
    subjects = [str(tmp_path / f"sub-{i:02d}") for i in range(11)]
    seen = {}

    def fake_stream(load_matrix, n_subjects, components, device=0, prefetch=True, options=None):
        for idx in range(n_subjects):
            yield idx, {"grey_components": np.zeros((2, 1)), "white_components": np.zeros((1, 3))}

    monkeypatch.setattr(drmod, "stream_subjects", fake_stream)
    monkeypatch.setattr(drmod, "_postprocess_options", lambda *a, **k: {})
    for world in (1, 2, 3, 8):
        done = []
        for rank in range(world):
            saved = []
            out = drmod.run_locally(subjects, {"grey_components": np.zeros((2, 1))}, ["l.gii"], rank=rank, world=world,
                                    save=lambda d, res: saved.append(d))
            assert list(out) == saved == subjects[rank::world]
            assert all(v is None for v in out.values())          # results are handed to `save`, not kept
            done += saved
        assert sorted(done) == subjects
    kept = drmod.run_locally(subjects, {"grey_components": np.zeros((2, 1))}, ["l.gii"], rank=1, world=4)
    assert list(kept) == subjects[1::4] and all(v is not None for v in kept.values())


def test_read_ahead_reads_every_subject_once_and_hands_the_bytes_on(monkeypatch, tmp_path):
    """run_locally(read_ahead=N): N host threads read / decompress the next subjects' files while the current one is
    loaded; every file is read exactly once, in time, and its bytes reach load_fdt_matrix(raw=...).  Host logic only:
    the reader, the loader and the subject stream are recorders."""
    import threading
    import time
    import nfact_b200.dual_regression as drmod
    import nfact_b200.matrix_handling as mh
    subjects = []
    for i in range(7):
        d = tmp_path / f"sub-{i:02d}"
        d.mkdir()
        (d / ("fdt_matrix2.dot.lz4" if i % 2 else "fdt_matrix2.dot")).write_bytes(b"x")
        subjects.append(str(d))
    reads, loads, lock = [], [], threading.Lock()
    in_flight = {"now": 0, "max": 0}

    def fake_read(path):
        with lock:
            in_flight["now"] += 1
            in_flight["max"] = max(in_flight["max"], in_flight["now"])
        time.sleep(0.02)
        with lock:
            in_flight["now"] -= 1
            reads.append(path)
        return ("bytes of " + path).encode()

    def fake_load(matfile, resident=False, handle=None, raw=None):
        loads.append((matfile, raw))
        return object()

    def fake_stream(load_matrix, n_subjects, components, device=0, prefetch=True, options=None):
        for idx in range(n_subjects):
            load_matrix(idx, None)
            yield idx, {"grey_components": np.zeros((2, 1)), "white_components": np.zeros((1, 3))}

    monkeypatch.setattr(mh, "_read_bytes", fake_read)
    monkeypatch.setattr(mh, "load_fdt_matrix", fake_load)
    monkeypatch.setattr(drmod, "stream_subjects", fake_stream)
    monkeypatch.setattr(drmod, "_postprocess_options", lambda *a, **k: {})
    comps = {"grey_components": np.zeros((2, 1))}
    for ahead in (1, 3):
        reads.clear(), loads.clear()
        in_flight.update(now=0, max=0)
        out = drmod.run_locally(subjects, comps, ["l.gii"], read_ahead=ahead)
        files = [drmod._subject_matrix_file(d) for d in subjects]
        assert list(out) == subjects and sorted(reads) == sorted(files)           # each file once
        assert [m for m, _ in loads] == files                                     # loaded in order
        assert all(raw == ("bytes of " + m).encode() for m, raw in loads)         # the reader's bytes
        assert 1 <= in_flight["max"] <= ahead
    reads.clear(), loads.clear()
    drmod.run_locally(subjects[:3], comps, ["l.gii"])                            # default: no reader threads
    assert reads == [] and all(raw is None for _, raw in loads)
    reads.clear(), loads.clear()
    drmod.run_locally(subjects, comps, ["l.gii"], rank=1, world=3, read_ahead=2)  # with subject sharding
    assert [m for m, _ in loads] == [drmod._subject_matrix_file(d) for d in subjects[1::3]]
