"""Pin the CPU oracle against the third-party code NFACT delegates its arithmetic to, run live (sklearn's NMF and
scipy's nnls are installed on both boxes): random small problems beyond the committed golden vectors -- ragged shapes,
zero rows / columns (the `HHt[t,t] == 0` and `den == 0 -> eps` branches), every regularisation mix, both solvers."""
import warnings

import numpy as np
import pytest

from oracle import nfact_oracle as orc
from conftest import rel_fro


def _problem(rng, m, n, k, zero_cols=0, zero_rows=0):
    W = rng.gamma(0.5, 1.0, (m, k))
    H = rng.gamma(0.5, 1.0, (k, n))
    X = (np.floor(40.0 * (W @ H)) * (rng.random((m, n)) < 0.4)).astype(np.float32)
    X[:, :zero_cols] = 0.0
    X[:zero_rows, :] = 0.0
    avg = np.sqrt(X.mean() / k)
    W0 = np.abs(avg * rng.standard_normal((m, k))).astype(np.float32)
    H0 = np.abs(avg * rng.standard_normal((k, n))).astype(np.float32)
    return X, W0, H0


def _sklearn_fit(X, W0, H0, solver, alpha_W, alpha_H, l1_ratio, max_iter, tol=0.0):
    from sklearn.decomposition import NMF
    est = NMF(n_components=W0.shape[1], init="custom", solver=solver, beta_loss="frobenius", tol=tol, max_iter=max_iter,
              random_state=None, alpha_W=alpha_W, alpha_H=alpha_H, l1_ratio=l1_ratio, shuffle=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        W = est.fit_transform(X, W=W0.copy(), H=H0.copy())
    return W, est.components_, est.n_iter_


CASES = [  # m, n, k, zero_cols, zero_rows, alpha_W, alpha_H, l1_ratio
    (37, 53, 4, 0, 0, 0.0, "same", 0.0),
    (61, 47, 5, 6, 0, 0.01, "same", 1.0),          # NFACT's mix (pure L1), leading zero columns like the fixtures
    (50, 80, 7, 0, 3, 0.02, 0.005, 0.5),
    (33, 29, 3, 4, 2, 0.0, 0.03, 0.0),             # pure L2 on H only
    (90, 41, 11, 0, 0, 0.005, "same", 0.3),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("solver", ["cd", "mu"])
def test_fit_matches_sklearn(case, solver):
    m, n, k, zc, zr, aW, aH, l1r = case
    rng = np.random.default_rng(1000 + m * n + k)
    X, W0, H0 = _problem(rng, m, n, k, zc, zr)
    regs = orc.compute_regularization(m, n, aW, aH, l1r)
    for iters in (1, 4, 12):
        Ws, Hs, _ = _sklearn_fit(X, W0, H0, solver, aW, aH, l1r, iters)
        fit = orc.fit_cd if solver == "cd" else orc.fit_mu
        Wo, Ho, _ = fit(X, W0, H0, *regs, tol=0, max_iter=iters)
        assert rel_fro(Wo, Ws) < 2e-5, (case, solver, iters)
        assert rel_fro(Ho, Hs) < 2e-5, (case, solver, iters)


@pytest.mark.parametrize("solver", ["cd", "mu"])
def test_stopping_iteration_matches_sklearn(solver):
    rng = np.random.default_rng(77)
    X, W0, H0 = _problem(rng, 64, 70, 6)
    regs = orc.compute_regularization(64, 70, 0.01, "same", 1.0)
    _, _, n_iter = _sklearn_fit(X, W0, H0, solver, 0.01, "same", 1.0, 200, tol=1e-4)
    fit = orc.fit_cd if solver == "cd" else orc.fit_mu
    _, _, n_iter_oracle = fit(X, W0, H0, *regs, tol=1e-4, max_iter=200)
    assert abs(n_iter_oracle - n_iter) <= (2 if solver == "cd" else 0), (n_iter_oracle, n_iter)


def test_nnls_matches_scipy_on_random_problems():
    from scipy.optimize import nnls
    rng = np.random.default_rng(5)
    for trial in range(40):
        m, k = int(rng.integers(5, 120)), int(rng.integers(1, 25))
        A = rng.gamma(0.4, 1.0, (m, k)) * (rng.random((m, k)) < 0.6)
        if trial % 5 == 0:
            A[:, 0] = 0.0                                   # a dead component
        b = rng.gamma(0.5, 1.0, m) * (rng.random(m) < 0.5) * 100
        if trial % 7 == 0:
            b[:] = 0.0
        ref = nnls(A, b)[0]
        got = orc.nnls(A, b)
        assert np.allclose(got, ref, rtol=1e-8, atol=1e-10 * max(1.0, np.abs(ref).max())), trial


def test_init_random_matches_sklearn():
    from sklearn.decomposition._nmf import _initialize_nmf
    rng = np.random.default_rng(8)
    X = rng.gamma(0.5, 1.0, (45, 31)).astype(np.float32)
    for seed in (0, 3, 12345):
        W, H = _initialize_nmf(X, 6, init="random", random_state=seed)
        Wo, Ho = orc.init_random(X, 6, seed)
        assert np.array_equal(W, Wo) and np.array_equal(H, Ho)
