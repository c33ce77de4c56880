"""NNDSVD initialisation with every product against X on the device, single GPU or seed-row sharded.

Follows sklearn/decomposition/_nmf.py:309-347 on top of sklearn/utils/extmath.py:313-383, 560-630
(``_randomized_svd``: n_oversamples=10, n_iter 7|4, LU-normalised power iterations, transpose='auto',
u-based sign flip).  X never leaves HBM: the 2*n_iter+2 products X.Q / X'.Q go through the
split-precision tensor-core GEMM (``nfb_matrix_matmul``); the skinny LU / QR / SVD factorisations of the
(k+10)-column panels use torch.linalg on the device (off the hot loop).

Row-sharded runs (one process per GPU, X partitioned by seed rows, NFACT's default ``init='nndsvd'``
reached through ``nmf_decomp`` under torchrun): a panel is either *seed-sharded* ``[m_local, r]`` or
*replicated* ``[n, r]``.  ``X q`` of a replicated panel is local; ``X' q`` of a seed-sharded panel is summed
over the ranks (all-reduce of ``n x r`` floats); a seed-sharded panel is orthonormalised by TSQR (local
Householder QR, all-gather of the ``r x r`` triangles, QR of the stack) -- any basis of the same column space
gives the same final subspace, which is all the power iteration needs.  The Gaussian test matrix is drawn in
full from the same RandomState on every rank and sliced by global seed row, so the result does not depend on
the number of ranks beyond rounding.
"""
from __future__ import annotations

import numpy as np


class Comm:
    """The three collectives the sharded initialisation needs, over ``torch.distributed`` (NCCL for device
    tensors, gloo in the CPU tests).  ``world == 1``: every method is the identity."""

    def __init__(self, rank: int = 0, world: int = 1, group=None):
        self.rank, self.world, self.group = int(rank), int(world), group

    def allreduce(self, t):
        if self.world > 1:
            import torch.distributed as dist
            t = t.contiguous()
            dist.all_reduce(t, group=self.group)
        return t

    def allgather(self, t):
        """list of every rank's tensor (same shape everywhere)."""
        if self.world == 1:
            return [t]
        import torch
        import torch.distributed as dist
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t.contiguous(), group=self.group)
        return out

    def row_layout(self, m_local: int, like):
        """(first global seed row of this rank, total seed rows) from every rank's local row count."""
        if self.world == 1:
            return 0, int(m_local)
        import torch
        counts = self.allgather(torch.tensor([int(m_local)], dtype=torch.int64, device=like))
        counts = [int(c.item()) for c in counts]
        return sum(counts[:self.rank]), sum(counts)


def device_operator(xdev):
    """``op(trans, q)`` on a resident ``DeviceMatrix``: q torch [rows_K, kk] float32 on the device ->
    X_local @ q (trans False) or X_local.T @ q (trans True)."""
    import torch

    def op(trans: bool, q):
        m, n = xdev.shape
        out_rows = n if trans else m
        qt = q.t().contiguous()                                  # [kk, K] component-major
        out = torch.empty((qt.shape[0], out_rows), dtype=torch.float32, device=q.device)
        torch.cuda.current_stream().synchronize()
        xdev.matmul_dev(trans, qt.data_ptr(), qt.shape[0], out.data_ptr(), out_rows)
        xdev.handle.synchronize()
        return out.t()                                           # [out_rows, kk]
    return op


def _tsqr(a, comm: Comm):
    """Reduced QR of a seed-sharded panel ``[m_local, r]``: returns this rank's rows of Q (and R)."""
    import torch
    q1, r1 = torch.linalg.qr(a, mode="reduced")                  # [m_l, p], [p, r] with p = min(m_l, r)
    r_cols = a.shape[1]
    pad = torch.zeros((r_cols, r_cols), dtype=a.dtype, device=a.device)
    pad[:r1.shape[0]] = r1                                       # ranks with fewer than r rows contribute zero rows
    stack = torch.cat(comm.allgather(pad), dim=0)                # [P r, r], identical on every rank
    q2, r2 = torch.linalg.qr(stack, mode="reduced")
    mine = q2[comm.rank * r_cols: comm.rank * r_cols + r1.shape[0]]
    return q1 @ mine, r2


def randomized_svd_device(op, m_local: int, n: int, k: int, rng, comm: Comm | None = None, device=None,
                          n_oversamples: int = 10):
    """sklearn ``_randomized_svd(X, k)`` with X given as the operator ``op``; returns
    (U rows of this rank ``[m_local, k]``, s ``[k]``, Vt ``[k, n]``)."""
    import torch
    comm = comm or Comm()
    dev = device if device is not None else torch.device("cpu")
    row0, m = comm.row_layout(m_local, dev)
    n_random = k + n_oversamples
    n_iter = 7 if k < 0.1 * min(m, n) else 4
    transpose = m < n                        # sklearn works on M = X.T in that case
    cols_m = m if transpose else n           # M.shape[1]
    q_full = rng.normal(size=(cols_m, n_random)).astype(np.float32, copy=False)
    if transpose:                            # test matrix indexed by seed row: this rank's rows of it
        q_full = q_full[row0:row0 + m_local]
    q = torch.from_numpy(np.ascontiguousarray(q_full)).to(dev)

    def x_mul(q_rep):                        # X q: replicated [n, r] -> seed-sharded [m_local, r]
        return op(False, q_rep.contiguous())

    def xt_mul(q_sh):                        # X' q: seed-sharded [m_local, r] -> replicated [n, r]
        return comm.allreduce(op(True, q_sh.contiguous()).contiguous())

    def lu_normalise(a):
        # sklearn normalises the power iterations with the P.L factor of an LU decomposition; any basis of the
        # same column space gives the same final subspace, and a tall-skinny LU (a serial panel factorisation)
        # is the slow way to get one on a GPU -- Householder QR from cuSOLVER takes its place for big panels
        if a.shape[0] > 8192:
            return torch.linalg.qr(a, mode="reduced")[0]
        p, l, _ = torch.linalg.lu(a)
        return p @ l

    def normalise_sharded(a):
        return lu_normalise(a) if comm.world == 1 else _tsqr(a, comm)[0]

    def qr_sharded(a):
        return torch.linalg.qr(a, mode="reduced")[0] if comm.world == 1 else _tsqr(a, comm)[0]

    # M q and M' q in terms of X: M = X' when `transpose`
    if transpose:
        m_mul, mt_mul = xt_mul, x_mul
        norm_m, norm_mt = lu_normalise, normalise_sharded
    else:
        m_mul, mt_mul = x_mul, xt_mul
        norm_m, norm_mt = normalise_sharded, lu_normalise
    for _ in range(n_iter):
        q = norm_m(m_mul(q)) if n_iter > 2 else m_mul(q)
        q = norm_mt(mt_mul(q)) if n_iter > 2 else mt_mul(q)
    if transpose:
        q, _ = torch.linalg.qr(m_mul(q), mode="reduced")         # replicated [n, r]
        b_local = mt_mul(q).t()                                  # Q' M restricted to this rank's seeds [r, m_local]
        if comm.world > 1:
            b = torch.zeros((b_local.shape[0], m), dtype=b_local.dtype, device=dev)
            b[:, row0:row0 + m_local] = b_local
            b = comm.allreduce(b)
        else:
            b = b_local
    else:
        q = qr_sharded(m_mul(q))                                 # seed-sharded [m_local, r]
        b = mt_mul(q).t()                                        # replicated [r, n]
    uhat, s, vt = torch.linalg.svd(b, full_matrices=False)
    u = q @ uhat
    if not transpose:
        # svd_flip, u-based: the sign of the entry of largest magnitude in every column of the (sharded) U
        absu = u.abs()
        top, piv = absu.max(dim=0) if u.shape[0] else (torch.zeros(u.shape[1], device=dev), None)
        sign_local = torch.sign(u[piv, torch.arange(u.shape[1], device=dev)]) if piv is not None else top
        if comm.world > 1:
            tops = torch.stack(comm.allgather(top.contiguous()))          # [P, r]
            sgns = torch.stack(comm.allgather(sign_local.contiguous()))
            best = torch.argmax(tops, dim=0)                              # first rank holding the maximum
            signs = sgns[best, torch.arange(u.shape[1], device=dev)]
        else:
            signs = sign_local
    else:
        piv = torch.argmax(vt.abs(), dim=1)
        signs = torch.sign(vt[torch.arange(vt.shape[0], device=dev), piv])
    u = u * signs[None, :]
    vt = vt * signs[:, None]
    if transpose:
        return vt[:k, row0:row0 + m_local].t(), s[:k], u[:, :k].t()
    return u[:, :k], s[:k], vt[:k, :]


def nndsvd_from_operator(op, m_local: int, n: int, k: int, rng, comm: Comm | None = None, device=None,
                         eps: float = 1e-6):
    """(W rows of this rank ``[m_local, k]``, H ``[k, n]``) float32 numpy arrays
    (sklearn _nmf.py:309-347, init='nndsvd')."""
    import torch
    comm = comm or Comm()
    u, s, v = randomized_svd_device(op, m_local, n, k, rng, comm, device)
    u, v, s = u.contiguous(), v.contiguous(), s.to(torch.float64)
    w = torch.zeros_like(u)
    h = torch.zeros_like(v)
    w[:, 0] = torch.sqrt(s[0]).float() * u[:, 0].abs()
    h[0, :] = torch.sqrt(s[0]).float() * v[0, :].abs()
    x_p, x_n = u.clamp(min=0), (-u).clamp(min=0)
    y_p, y_n = v.clamp(min=0), (-v).clamp(min=0)
    if comm.world > 1:       # column norms of the seed-sharded side: squared sums over all ranks
        sq = comm.allreduce(torch.stack([(x_p.double() ** 2).sum(dim=0), (x_n.double() ** 2).sum(dim=0)]))
        xp_n, xn_n = torch.sqrt(sq[0]).float().double(), torch.sqrt(sq[1]).float().double()
    else:
        xp_n, xn_n = torch.linalg.norm(x_p, dim=0).double(), torch.linalg.norm(x_n, dim=0).double()
    yp_n, yn_n = torch.linalg.norm(y_p, dim=1).double(), torch.linalg.norm(y_n, dim=1).double()
    m_p, m_n = xp_n * yp_n, xn_n * yn_n
    use_p = m_p > m_n
    sigma = torch.where(use_p, m_p, m_n)
    lbd = torch.sqrt(s * sigma)
    for j in range(1, k):
        if bool(use_p[j]):
            w[:, j] = (lbd[j] / xp_n[j]).float() * x_p[:, j]
            h[j, :] = (lbd[j] / yp_n[j]).float() * y_p[j, :]
        else:
            w[:, j] = (lbd[j] / xn_n[j]).float() * x_n[:, j]
            h[j, :] = (lbd[j] / yn_n[j]).float() * y_n[j, :]
    w[w < eps] = 0
    h[h < eps] = 0
    return w.cpu().numpy(), h.cpu().numpy()


def nndsvd_device(xdev, k: int, rng, eps: float = 1e-6, comm: Comm | None = None):
    """NNDSVD of a resident ``DeviceMatrix`` (this rank's seed rows of X when ``comm.world > 1``)."""
    import torch
    dev = torch.device("cuda", xdev.handle.device)
    m_local, n = xdev.shape
    return nndsvd_from_operator(device_operator(xdev), m_local, n, k, rng, comm, dev, eps)
