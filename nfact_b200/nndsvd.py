"""NNDSVD initialisation with every product against X on the device.

Follows sklearn/decomposition/_nmf.py:309-347 on top of sklearn/utils/extmath.py:313-383, 560-630
(``_randomized_svd``: n_oversamples=10, n_iter 7|4, LU-normalised power iterations, transpose='auto',
u-based sign flip).  X never leaves HBM: the 2*n_iter+2 products X.Q / X'.Q go through the
split-precision tensor-core GEMM (``nfb_matrix_matmul``); the skinny LU / QR / SVD factorisations of the
(k+10)-column panels use torch.linalg on the device (off the hot loop).
"""
from __future__ import annotations

import numpy as np


def _matmul(xdev, trans: bool, q):
    """q: torch [rows_K, kk] float32 on the device -> X @ q (trans False) or X.T @ q (trans True)."""
    import torch
    m, n = xdev.shape
    out_rows = n if trans else m
    qt = q.t().contiguous()                                  # [kk, K] component-major
    out = torch.empty((qt.shape[0], out_rows), dtype=torch.float32, device=q.device)
    torch.cuda.current_stream().synchronize()
    xdev.matmul_dev(trans, qt.data_ptr(), qt.shape[0], out.data_ptr(), out_rows)
    return out.t()                                           # [out_rows, kk]


def randomized_svd_device(xdev, k: int, rng, n_oversamples: int = 10):
    import torch
    dev = torch.device("cuda", xdev.handle.device)
    m, n = xdev.shape
    n_random = k + n_oversamples
    n_iter = 7 if k < 0.1 * min(m, n) else 4
    transpose = m < n                        # sklearn works on M = X.T in that case
    cols_m = m if transpose else n           # M.shape[1]
    q = torch.from_numpy(rng.normal(size=(cols_m, n_random)).astype(np.float32, copy=False)).to(dev)

    def mul(mat_t: bool, qq):                # M @ qq  (mat_t False) or M.T @ qq (mat_t True)
        return _matmul(xdev, trans=(mat_t != transpose), q=qq.contiguous())

    def lu_normalise(a):
        # sklearn normalises the power iterations with the P.L factor of an LU decomposition; any basis of the
        # same column space gives the same final subspace, and a tall-skinny LU (a serial panel factorisation)
        # is the slow way to get one on a GPU -- Householder QR from cuSOLVER takes its place for big panels
        if a.shape[0] > 8192:
            return torch.linalg.qr(a, mode="reduced")[0]
        p, l, _ = torch.linalg.lu(a)
        return p @ l

    for _ in range(n_iter):
        q = lu_normalise(mul(False, q)) if n_iter > 2 else mul(False, q)
        q = lu_normalise(mul(True, q)) if n_iter > 2 else mul(True, q)
    q, _ = torch.linalg.qr(mul(False, q), mode="reduced")
    b = mul(True, q).t()                     # Q.T @ M  ==  (M.T @ Q).T
    uhat, s, vt = torch.linalg.svd(b, full_matrices=False)
    u = q @ uhat
    if not transpose:
        piv = torch.argmax(u.abs(), dim=0)
        signs = torch.sign(u[piv, torch.arange(u.shape[1], device=dev)])
    else:
        piv = torch.argmax(vt.abs(), dim=1)
        signs = torch.sign(vt[torch.arange(vt.shape[0], device=dev), piv])
    u = u * signs[None, :]
    vt = vt * signs[:, None]
    if transpose:
        return vt[:k, :].t(), s[:k], u[:, :k].t()
    return u[:, :k], s[:k], vt[:k, :]


def nndsvd_device(xdev, k: int, rng, eps: float = 1e-6):
    """Returns (W [m,k], H [k,n]) float32 numpy arrays (sklearn _nmf.py:309-347, init='nndsvd')."""
    import torch
    u, s, v = randomized_svd_device(xdev, k, rng)
    u, v, s = u.contiguous(), v.contiguous(), s.to(torch.float64)
    w = torch.zeros_like(u)
    h = torch.zeros_like(v)
    w[:, 0] = torch.sqrt(s[0]).float() * u[:, 0].abs()
    h[0, :] = torch.sqrt(s[0]).float() * v[0, :].abs()
    x_p, x_n = u.clamp(min=0), (-u).clamp(min=0)
    y_p, y_n = v.clamp(min=0), (-v).clamp(min=0)
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
