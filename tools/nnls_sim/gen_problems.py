"""C3-like NNLS problems for the simulations: a synthetic 6000 x 9000 connectivity matrix (bench.py's generator at
small scale), 60 CD iterations of sklearn's NMF with NFACT's defaults, then the Gram matrices and right-hand sides of the
two dual-regression stages for 40 sampled columns / rows -> /tmp/nnls_hard.npz."""
import numpy as np, time, warnings
from sklearn.decomposition import NMF
rng = np.random.default_rng(0)
m, n, k = 6000, 9000, 200
h_star = rng.gamma(0.3, 1.0, (k, n)); w_star = rng.gamma(0.3, 1.0, (m, k))
X = (np.floor(300.0*(w_star@h_star))*(rng.random((m, n)) < 0.05)*(1e8/1.5e7)).astype(np.float32)
avg = np.sqrt(X.mean()/k)
W0 = np.abs(avg*rng.standard_normal((m, k))).astype(np.float32); H0 = np.abs(avg*rng.standard_normal((k, n))).astype(np.float32)
t = time.time()
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    est = NMF(n_components=k, init="custom", solver="cd", tol=0.0, max_iter=60, alpha_W=0.1, alpha_H="same", l1_ratio=1.0)
    W = est.fit_transform(X, W=W0, H=H0); H = est.components_
print("nmf", time.time()-t, "W nnz", (W > 0).mean(), "H nnz", (H > 0).mean())
W64, H64, X64 = W.astype(np.float64), H.astype(np.float64), X.astype(np.float64)
cols = rng.choice(n, 40, replace=False); rows = rng.choice(m, 40, replace=False)
G1 = W64.T@W64; R1 = W64.T@X64[:, cols]
G2 = H64@H64.T; R2 = H64@X64[rows, :].T
live1 = np.diag(G1) > 0; live2 = np.diag(G2) > 0
print("dead comps", (~live1).sum(), (~live2).sum(), "diag spread", np.diag(G1)[live1].max()/np.diag(G1)[live1].min())
np.savez("/tmp/nnls_hard.npz", G1=G1[np.ix_(live1, live1)], R1=R1[live1], G2=G2[np.ix_(live2, live2)], R2=R2[live2])
