"""The reference's tools on several GPUs of one node: ``torchrun ... -m nfact_b200.dropin nfact_decomp ...``.

One process per GPU.  Rank 0 is the only process that runs NFACT's own main -- argument parsing, directory set-up,
reading the subjects' files, writing images and logs happen once, exactly as in the single-process tool.  The other
ranks sit in ``serve()`` and execute the jobs rank 0 broadcasts when its (rebound) hot-path functions are reached:

``decomp``   ``nmf_decomp(parameters, X)`` on a host matrix (NFACT's ``--no_sso`` flow and every direct call,
             NFACT/decomp/decomposition/sso/nmfsso.py:32-59).  X is partitioned by seed rows: rank 0 uploads row
             blocks and sends rank r its rows over NVLink, every rank builds the fp16 planes of its shard, the
             default NNDSVD (or random) initialisation is computed across the ranks (nndsvd.py), the solve runs
             row-sharded inside libnfact_b200 (solver.cu / p2p.cu), and the W shards travel back to rank 0.
``sso_runs`` the N independent random-initialised solves of NMF-sso (nmfsso.py:140-190, where the reference forks
             joblib workers over a shared-memory copy of X): the replicas axis.  X is broadcast once, every rank
             keeps its own resident copy and solves runs rank, rank + world, ... on its own GPU with no collective
             on the data path; the thresholded components go back to rank 0, which clusters them and runs the final
             solve.
``dr``       the subject loop of ``nfact_dr`` (set_up_dual_regression.py:186-235): subjects rank, rank + world, ...
             per GPU, no collective; every rank writes its own subjects' images with the reference's writer.
``stop``     leave ``serve()``.

torch.distributed (NCCL) is the plumbing: the control messages are pickled dictionaries, the matrices travel as
device tensors.
"""
from __future__ import annotations

import os

import numpy as np

_ctx = None


class Context:
    def __init__(self, rank, world, local_rank, handle, solo, dev):
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.handle = handle          # carries the NCCL communicator: row-sharded solves
        self.solo = solo              # same GPU, no communicator: independent (replica) solves
        self.dev = dev


def init_from_env():
    """Bring up the multi-GPU context when launched by torchrun with more than one rank; None otherwise."""
    global _ctx
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if _ctx is not None:
        return _ctx
    import torch
    import torch.distributed as dist
    from . import device as nfdev
    from .sharding import init_nccl_from_torch
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    handle = nfdev.Handle(local)
    init_nccl_from_torch(handle)
    solo = nfdev.Handle(local)
    nfdev.set_default_handle(solo)    # plain calls (post-processing, similarity, ingest) stay on this GPU alone
    _ctx = Context(rank, world, local, handle, solo, dev)
    return _ctx


def context():
    return _ctx


def active() -> bool:
    return _ctx is not None and _ctx.world > 1


def _bcast(obj=None):
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def shutdown() -> None:
    """Rank 0, at the end of the tool: release the workers."""
    global _ctx
    if _ctx is None:
        return
    import torch.distributed as dist
    if _ctx.rank == 0:
        _bcast({"op": "stop"})
    dist.barrier()
    dist.destroy_process_group()
    _ctx = None


def serve() -> None:
    """Ranks > 0: execute rank 0's jobs until it says stop.  A failing job takes the whole launch down (torchrun
    kills the other ranks) instead of leaving rank 0 waiting in a collective."""
    import sys
    import traceback
    while True:
        job = _bcast()
        try:
            if job["op"] == "stop":
                break
            {"decomp": _decomp_job, "sso_runs": _sso_job, "dr": _dr_job}[job["op"]](job, None)
        except BaseException:
            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


# ---- moving the matrix -----------------------------------------------------------------------------------
_CHUNK_BYTES = 1 << 30


def _scatter_rows(x_host, m, n):
    """Rank 0 holds X [m, n] float32 on the host; returns this rank's seed rows as a resident ``DeviceMatrix`` on
    the sharded handle.  Rank 0 uploads <= 1 GiB row blocks and sends each to its owner (NCCL send / recv over
    NVLink); an owner collects its rows in one fp32 device buffer that lives only until its planes are built."""
    import torch
    import torch.distributed as dist
    from .device import DeviceMatrix
    from .sharding import row_range
    ctx = _ctx
    r0, r1 = row_range(m, ctx.rank, ctx.world)
    step = max(1, _CHUNK_BYTES // (4 * n))
    if ctx.rank == 0:
        for dst in range(1, ctx.world):
            d0, d1 = row_range(m, dst, ctx.world)
            for c0 in range(d0, d1, step):
                c1 = min(c0 + step, d1)
                blk = torch.from_numpy(x_host[c0:c1]).to(ctx.dev, non_blocking=False)
                dist.send(blk, dst=dst)
                del blk
        return DeviceMatrix.from_host(x_host[r0:r1], ctx.handle)
    buf = torch.empty((r1 - r0, n), dtype=torch.float32, device=ctx.dev)
    for c0 in range(0, r1 - r0, step):
        dist.recv(buf[c0:min(c0 + step, r1 - r0)], src=0)
    torch.cuda.synchronize()
    xdev = DeviceMatrix.from_device_ptr(buf.data_ptr(), r1 - r0, n, n, handle=ctx.handle)
    del buf
    torch.cuda.empty_cache()
    return xdev


def _replicate(x_host, m, n):
    """Every rank ends up with the whole X resident on its unsharded handle (broadcast of <= 1 GiB row blocks)."""
    import torch
    import torch.distributed as dist
    from .device import DeviceMatrix
    ctx = _ctx
    buf = torch.empty((m, n), dtype=torch.float32, device=ctx.dev)
    step = max(1, _CHUNK_BYTES // (4 * n))
    for c0 in range(0, m, step):
        c1 = min(c0 + step, m)
        if ctx.rank == 0:
            buf[c0:c1].copy_(torch.from_numpy(x_host[c0:c1]))
        dist.broadcast(buf[c0:c1], src=0)
    torch.cuda.synchronize()
    xdev = DeviceMatrix.from_device_ptr(buf.data_ptr(), m, n, n, handle=ctx.solo)
    del buf
    torch.cuda.empty_cache()
    return xdev


def _gather_rows_to_root(local: np.ndarray, m: int, cols: int):
    """W shards [rows_r, cols] -> the full [m, cols] on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from .sharding import row_range
    ctx = _ctx
    if ctx.rank != 0:
        dist.send(torch.from_numpy(np.ascontiguousarray(local)).to(ctx.dev), dst=0)
        return None
    out = np.empty((m, cols), dtype=local.dtype)
    r0, r1 = row_range(m, 0, ctx.world)
    out[r0:r1] = local
    for src in range(1, ctx.world):
        s0, s1 = row_range(m, src, ctx.world)
        buf = torch.empty((s1 - s0, cols), dtype=torch.from_numpy(local[:0]).dtype, device=ctx.dev)
        dist.recv(buf, src=src)
        out[s0:s1] = buf.cpu().numpy()
    return out


# ---- jobs -------------------------------------------------------------------------------------------------
def nmf_fit_sharded(parameters: dict, x_host, W=None, H=None):
    """Rank 0's side of a row-sharded ``NMF(**parameters).fit_transform(X, W=W, H=H)``: returns
    (W [m, k], H [k, n], n_iter, reconstruction_err) like the estimator's attributes."""
    x_host = np.ascontiguousarray(x_host, dtype=np.float32)
    m, n = x_host.shape
    job = {"op": "decomp", "parameters": dict(parameters), "shape": (m, n), "custom": W is not None and H is not None}
    _bcast(job)
    return _decomp_job(job, (x_host, W, H))


def _decomp_job(job, payload):
    import torch
    import torch.distributed as dist
    from .nmf import NMF
    from .sharding import row_range
    ctx = _ctx
    m, n = job["shape"]
    params = job["parameters"]
    k = int(params["n_components"])
    x_host, W, H = payload if payload is not None else (None, None, None)
    xdev = _scatter_rows(x_host, m, n)
    try:
        w_rows = h0 = None
        if job["custom"]:
            r0, r1 = row_range(m, ctx.rank, ctx.world)
            wbuf = torch.empty((m, k), dtype=torch.float32, device=ctx.dev)
            hbuf = torch.empty((k, n), dtype=torch.float32, device=ctx.dev)
            if ctx.rank == 0:
                wbuf.copy_(torch.from_numpy(np.ascontiguousarray(W, dtype=np.float32)))
                hbuf.copy_(torch.from_numpy(np.ascontiguousarray(H, dtype=np.float32)))
            dist.broadcast(wbuf, src=0)
            dist.broadcast(hbuf, src=0)
            w_rows, h0 = wbuf[r0:r1].cpu().numpy(), hbuf.cpu().numpy()
            del wbuf, hbuf
        est = NMF(**params)
        w_local = est.fit_transform(xdev, W=w_rows, H=h0)
    finally:
        xdev.free()
    w_full = _gather_rows_to_root(np.asarray(w_local), m, k)
    if ctx.rank != 0:
        return None
    return w_full, est.components_, est.n_iter_, est.reconstruction_err_


def sso_runs_distributed(parameters: dict, x_host, num_int: int, threshold: float = 3):
    """Rank 0's side of the N random-initialised NMF-sso runs spread over the ranks (the replicas axis); returns
    ({"grey": [W_i], "white": [thresholded H_i]} in run order, this rank's resident X for the final solve)."""
    x_host = np.ascontiguousarray(x_host, dtype=np.float32)
    m, n = x_host.shape
    job = {"op": "sso_runs", "parameters": dict(parameters), "shape": (m, n), "num_int": int(num_int),
           "threshold": float(threshold)}
    _bcast(job)
    return _sso_job(job, x_host)


def _sso_job(job, x_host):
    import torch
    import torch.distributed as dist
    from .matrix_handling import thresholding
    from .nmf import nmf_decomp
    from .utils import colours, nprint
    ctx = _ctx
    m, n = job["shape"]
    params = dict(job["parameters"], init="random", random_state=None)
    k = int(params["n_components"])
    col = colours()
    xdev = _replicate(x_host, m, n)
    mine = list(range(ctx.rank, job["num_int"], ctx.world))
    done = {}
    for run in mine:
        nprint(f"{col['pink']}Run: {col['reset']}{run + 1}/{job['num_int']} (GPU {ctx.rank})", to_flush=True)
        res = nmf_decomp(params, xdev)
        done[run] = (res["grey_components"], thresholding(res["white_components"], job["threshold"]))
    if ctx.rank != 0:
        xdev.free()
        for run in mine:                                   # run order: rank 0 knows who owns which run
            for arr in done[run]:
                dist.send(torch.from_numpy(np.ascontiguousarray(arr)).to(ctx.dev), dst=0)
        return None
    results = {"grey": [], "white": []}
    for run in range(job["num_int"]):
        owner = run % ctx.world
        if owner == 0:
            g, w = done[run]
        else:
            gb = torch.empty((m, k), dtype=torch.float32, device=ctx.dev)
            wb = torch.empty((k, n), dtype=torch.float32, device=ctx.dev)
            dist.recv(gb, src=owner)
            dist.recv(wb, src=owner)
            g, w = gb.cpu().numpy(), wb.cpu().numpy()
            del gb, wb
        results["grey"].append(g)
        results["white"].append(w)
    return results, xdev


def dr_distributed(subject_dirs, components, seeds, threshold, normalise, save_spec, read_ahead=0):
    """Rank 0's side of the ``nfact_dr`` subject loop over all GPUs.  ``save_spec``: picklable description of the
    image writer every rank builds for its own subjects (see dropin.run_locally)."""
    job = {"op": "dr", "subjects": list(subject_dirs), "seeds": seeds, "threshold": int(threshold),
           "normalise": bool(normalise), "save_spec": save_spec, "read_ahead": int(read_ahead),
           "grey": np.asarray(components["grey_components"])}
    _bcast(job)
    return _dr_job(job, None)


def _dr_job(job, _payload):
    import torch.distributed as dist
    from . import device as nfdev
    from .dual_regression import run_locally
    ctx = _ctx
    save = None
    if job["save_spec"] is not None:
        from .dropin import make_dr_saver
        save = make_dr_saver(job["save_spec"])
    nfdev.set_default_handle(ctx.solo)
    out = run_locally(job["subjects"], {"grey_components": job["grey"]}, job["seeds"], threshold=job["threshold"],
                      normalise=job["normalise"], rank=ctx.rank, world=ctx.world, save=save,
                      read_ahead=job["read_ahead"])
    dist.barrier()
    return out
