"""Multi-GPU parity (run under torchrun, one rank per GPU; collected by tests/test_dist_gpu.py when >= 2 GPUs
are visible):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_parity_gpu.py
Row-sharded MU and CD iterations (W shards local, H columns sliced over ranks) against the unsharded CPU oracle,
once through the default NVLink peer-memory exchange (csrc/p2p.cu) and once with NFB_P2P=0 (NCCL reduce-scatter /
all-gather inside libnfact_b200); sklearn's stopping rule through nfb_nmf_run; and NFACT's default call
(init='nndsvd', computed across the ranks) against the same call on one GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nfact_b200.device import DeviceMatrix, Handle, NmfState, make_params  # noqa: E402
from nfact_b200.sharding import init_nccl_from_torch, row_range  # noqa: E402
from oracle import nfact_oracle as orc  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    handle = Handle(local)
    init_nccl_from_torch(handle)
    ok = True
    for path in ("p2p", "nccl"):
        if path == "nccl":
            os.environ["NFB_P2P"] = "0"
        else:
            os.environ.pop("NFB_P2P", None)
        ok &= solver_parity(rank, world, handle, path)
    os.environ.pop("NFB_P2P", None)
    ok &= default_flow_parity(rank, world, handle)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST PARITY", "OK" if int(flag) == 1 else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


def default_flow_parity(rank, world, handle):
    """nmf_decomp(get_parameters(None, 'nmf', k) | {random_state: 0}, X_rows): NFACT's default (NNDSVD init through
    the sharded randomized SVD, CD with the default penalties and stopping rule) on `world` ranks vs one GPU."""
    import nfact_b200
    from scipy.optimize import linear_sum_assignment
    ok = True
    for (m, n, k, init) in [(1201, 1733, 12, "nndsvd"), (1733, 1201, 12, "nndsvda"), (900, 1400, 9, "random"),
                            (700, 1100, 8, "nndsvdar")]:
        rng = np.random.default_rng(m)
        Wt, Ht = rng.gamma(0.3, 1.0, (m, k)), rng.gamma(0.3, 1.0, (k, n))
        X = (np.floor(300 * (Wt @ Ht)) * (rng.random((m, n)) < 0.4) * (1e8 / 1.5e7)).astype(np.float32)
        r0, r1 = row_range(m, rank, world)
        params = dict(nfact_b200.get_parameters(None, "nmf", k), random_state=0, init=init)
        xdev = DeviceMatrix.from_host(X[r0:r1], handle)
        res = nfact_b200.nmf_decomp(dict(params), xdev)
        xdev.free()
        parts = [None] * world
        dist.all_gather_object(parts, res["grey_components"])
        if rank == 0:
            single = Handle(handle.device)
            x1 = DeviceMatrix.from_host(X, single)
            one = nfact_b200.nmf_decomp(dict(params), x1)
            x1.free()
            W = np.concatenate(parts)
            regs = orc.compute_regularization(m, n, 0.1, "same", 1.0)
            obj = orc.nmf_objective(X, W, res["white_components"], *regs)
            obj1 = orc.nmf_objective(X, one["grey_components"], one["white_components"], *regs)

            def corr(a, b):
                a = a - a.mean(axis=1, keepdims=True)
                b = b - b.mean(axis=1, keepdims=True)
                an, bn = np.linalg.norm(a, axis=1, keepdims=True), np.linalg.norm(b, axis=1, keepdims=True)
                live = (an[:, 0] > 0) & (bn[:, 0] > 0)
                c = (a / np.where(an > 0, an, 1)) @ (b / np.where(bn > 0, bn, 1)).T
                rr, cc = linear_sum_assignment(-c)
                return float(c[rr, cc][live].min())
            ch, cw = corr(res["white_components"], one["white_components"]), corr(W.T, one["grey_components"].T)
            print(f"default flow {m}x{n} k={k} init={init}: objective rel diff {abs(obj - obj1) / obj1:.2e}, "
                  f"min matched corr H {ch:.5f} W {cw:.5f}", flush=True)
            ok &= abs(obj - obj1) <= 5e-3 * obj1 and ch >= 0.99 and cw >= 0.99
    return ok


def solver_parity(rank, world, handle, path):
    rng = np.random.default_rng(3)
    ok = True
    for (m, n, k) in [(517, 1203, 23), (300, 77, 5)]:
        Wt, Ht = rng.gamma(0.3, 1.0, (m, k)), rng.gamma(0.3, 1.0, (k, n))
        X = (np.floor(300 * (Wt @ Ht)) * (rng.random((m, n)) < 0.4) * (1e8 / 1.5e7)).astype(np.float32)
        W0, H0 = orc.init_random(X, k, 1)
        regs = orc.compute_regularization(m, n, 0.01, "same", 0.7)
        r0, r1 = row_range(m, rank, world)
        xdev = DeviceMatrix.from_host(X[r0:r1], handle)
        for solver in ("mu", "cd"):
            st = NmfState(xdev, k, make_params(solver, 200, 0.0, regs))
            st.set_factors(W0[r0:r1], H0)
            st_mode = st.exchange_mode
            ok &= st_mode == path or (path == "p2p" and os.environ.get("NFB_ALLOW_NO_P2P") == "1")
            viol = st.iterate(4, want_violation=(solver == "cd"))
            err = st.error()
            W, H = st.get_factors()
            st.free()
            rec = []
            (orc.fit_mu if solver == "mu" else orc.fit_cd)(X, W0, H0, *regs, tol=0, max_iter=4, record=rec)
            Wr, Hr = rec[-1][1], rec[-1][2]
            ew = np.linalg.norm(W - Wr[r0:r1]) / np.linalg.norm(Wr[r0:r1])
            eh = np.linalg.norm(H - Hr) / np.linalg.norm(Hr)
            eref = orc.frob_error(X, Wr, Hr)
            mode = st_mode
            line = f"rank {rank} [{path}->{mode}] {m}x{n} k={k} {solver}: relW={ew:.2e} relH={eh:.2e} err={err:.6g} (oracle {eref:.6g})"
            if solver == "cd":
                line += f" violation={viol:.6g} (oracle {rec[-1][3]:.6g})"
                ok &= abs(viol - rec[-1][3]) <= 1e-3 * rec[-1][3]
            print(line, flush=True)
            ok &= ew < 1e-4 and eh < 1e-4 and abs(err - eref) <= 1e-4 * eref
        # sklearn's stopping rule through nfb_nmf_run: the loop keeps one iteration in flight and rolls the
        # speculative one back when it stops early -- same n_iter and factors as the oracle, on every rank
        for tol in (3e-1, 1e-2):
            st = NmfState(xdev, k, make_params("cd", 60, tol, regs))
            st.set_factors(W0[r0:r1], H0)
            n_iter, err = st.run()
            W, H = st.get_factors()
            st.free()
            Wr, Hr, n_ref = orc.fit_cd(X, W0, H0, *regs, tol=tol, max_iter=60)
            ew = np.linalg.norm(W - Wr[r0:r1]) / np.linalg.norm(Wr[r0:r1])
            eh = np.linalg.norm(H - Hr) / np.linalg.norm(Hr)
            print(f"rank {rank} {m}x{n} k={k} cd run tol={tol}: n_iter={n_iter} (oracle {n_ref}) relW={ew:.2e} "
                  f"relH={eh:.2e}", flush=True)
            ok &= n_iter == n_ref and ew < 1e-3 and eh < 1e-3
        xdev.free()
    return ok


if __name__ == "__main__":
    main()
