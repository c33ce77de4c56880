"""Column-update counts of nnls_smem_kernel's strategies on C3-like problems (CPU simulation, no GPU):
cyclic Gauss-Seidel sweeps + CG phases (the fallback loop) against a greedy (Gauss-Southwell) phase of a given budget.
Run gen_problems.py first."""
import sys, numpy as np
d = np.load("/tmp/nnls_hard.npz")   # written by gen_problems.py
def solve(G, r, n_init=6, eta=1e-2, tol=1e-10, greedy=0, greedy_between=0):
    kk = G.shape[0]; inv = 1.0/np.diag(G); ax = 0
    h = np.zeros(kk); g = -r.copy()
    thresh = tol*np.abs(r).max(); rr_stop = 0.01*thresh*thresh
    ncg = 0; nsw = 0
    def sweep():
        nonlocal ax, nsw
        nsw += 1
        for blk in range(0, kk, 32):
            idx = [t for t in range(blk, min(blk+32, kk)) if h[t] > 0 or g[t] < 0]
            ax += len(idx)
            for t in idx:
                gt, ht = g[t], h[t]
                hn = max(ht - gt*inv[t], 0.0); delta = hn - ht
                if delta != 0.0:
                    g[:] += delta*G[:, t]; h[t] = hn
    def greedy_steps(budget, only_enter=False):
        nonlocal ax
        for _ in range(budget):
            pg = np.where(h > 0, np.abs(g), np.maximum(-g, 0))
            if only_enter: pg = np.where(h > 0, 0.0, pg)
            t = int(np.argmax(pg))
            if pg[t] <= thresh: break
            hn = max(h[t] - g[t]*inv[t], 0.0); delta = hn - h[t]
            g[:] += delta*G[:, t]; h[t] = hn; ax += 1
    first = True
    for outer in range(400):
        if first and greedy:
            greedy_steps(greedy)
        elif (not first) and greedy_between:
            greedy_steps(greedy_between, only_enter=False)
        else:
            for _ in range(n_init if first else 1): sweep()
        first = False
        viol = np.where(h > 0, np.abs(g), np.maximum(-g, 0)).max()
        if viol <= thresh: return h, ax, nsw, ncg
        in_p = h > 0
        dd = np.where(in_p, -g, 0.0); rr = dd@dd
        stop = max(rr_stop, eta*eta*rr)
        its = 0
        while its < 2*kk and rr > stop:
            its += 1; ncg += 1
            nz = dd != 0; ax += int(nz.sum())
            q = G[:, nz]@dd[nz]; dq = dd@q
            neg = in_p & (dd < 0)
            ratio = np.where(neg, -h/np.where(dd != 0, dd, 1), 1e300)
            amax = ratio.min()
            if not dq > 0: break
            alpha = rr/dq; hit = alpha >= amax
            if hit: alpha = amax
            hn = h + alpha*dd; g = g + alpha*q
            blocked = neg & ((hn <= 0) | (hit & (ratio <= amax)))
            h = np.where(blocked, 0.0, hn)
            if hit:
                in_p &= h > 0
                dd = np.where(in_p, -g, 0.0); rr = dd@dd; continue
            rr_new = (g[in_p]**2).sum(); beta = rr_new/rr
            dd = np.where(in_p, beta*dd - g, 0.0); rr = rr_new
    return h, ax, nsw, ncg
ncol = 12
for name in ("1", "2"):
    G = d["G"+name]; R = d["R"+name]
    s = 1.0/np.sqrt(np.diag(G)); Gs = G*s[:, None]*s[None, :]; Rs = R*s[:, None]
    ref = [solve(Gs, Rs[:, j])[0] for j in range(ncol)]
    for kw in (dict(), dict(greedy=150), dict(greedy=300), dict(greedy=500), dict(greedy=300, greedy_between=40), dict(greedy=300, greedy_between=100), dict(greedy=500, greedy_between=100)):
        tot = cg = 0; err = 0
        for j in range(ncol):
            h, ax, nsw, ncg = solve(Gs, Rs[:, j], **kw); tot += ax; cg += ncg
            err = max(err, np.abs(h-ref[j]).max()/np.abs(ref[j]).max())
        print("stage", name, kw, "cost/col", round(tot/ncol), "cg", round(cg/ncol,1), "err %.1e" % err, flush=True)
