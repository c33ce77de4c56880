"""Emulation of the greedy phase of nnls_smem_kernel with the kernel's exact selection logic: truncated high-word
keys, per-lane maxima, up to `batch` steps per key pass (best candidates of distinct lanes), stop at half the tolerance.
Run gen_problems.py first."""
import numpy as np
def hi(x):
    return (np.asarray(x, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)
def solve(G, r, batch=4, budget_mult=8, tol=1e-10):
    k = G.shape[0]; inv = 1.0/np.diag(G)
    h = np.zeros(k); g = -r.copy()
    thresh = tol*np.abs(r).max()
    stop_key = int(hi(0.5*thresh)) & 0x7fffff00
    steps = 0; updates = 0; nsel = 0
    idx = np.arange(k); lane = idx & 31
    left = budget_mult*k
    while left > 0:
        ghi = hi(g); hhi = hi(h)
        take = (hhi > 0) | (ghi < 0)
        key = np.where(take, (ghi & 0x7fffff00) | idx, 0)
        mine = np.zeros(32, dtype=np.int64)
        np.maximum.at(mine, lane, key)
        nsel += 1
        top = int(mine.max())
        if (top & 0x7fffff00) <= stop_key: break
        any_moved = False
        for b in range(batch):
            if b > 0:
                top = int(mine.max())
                if (top & 0x7fffff00) <= stop_key: break
            t = top & 0xff
            mine[t & 31] = 0
            hn = max(h[t] - g[t]*inv[t], 0.0); delta = hn - h[t]
            steps += 1; left -= 1
            if delta != 0.0:
                g += delta*G[:, t]; h[t] = hn; any_moved = True; updates += 1
        if not any_moved: break
    viol = np.where(h > 0, np.abs(g), np.maximum(-g, 0)).max()
    return h, steps, updates, nsel, viol <= thresh
for fname in ("/tmp/nnls_hard.npz",):
    d = np.load(fname)
    for name in ("1", "2"):
        G = d["G"+name]; R = d["R"+name]
        s = 1.0/np.sqrt(np.diag(G)); Gs = G*s[:, None]*s[None, :]; Rs = R*s[:, None]
        for batch in (1, 2, 4, 6, 8):
            res = [solve(Gs, Rs[:, j], batch) for j in range(20)]
            print(fname[-13:], name, "batch", batch, "steps %.0f" % np.mean([x[1] for x in res]), "updates %.0f" % np.mean([x[2] for x in res]), "selections %.0f" % np.mean([x[3] for x in res]), "max steps", max(x[1] for x in res), "conv", all(x[4] for x in res))
