"""Host-side glue of the NMF-sso flow that needs no GPU: the post-processing of the device correlations and the seeds of
the device random starts."""
import numpy as np


def test_finish_similarity_symmetric_clipped_dead_rows():
    """`_finish_similarity` turns the float32 device correlations into what `compute_similairty_matrix`
    (NFACT sso_functions.py:49-68) returns: float64, symmetric, unit diagonal, clipped to [0, 1]; a constant (dead)
    component is dissimilar to everything instead of NaN."""
    from nfact_b200.nmfsso import _finish_similarity, sim2dis
    rng = np.random.default_rng(0)
    z = rng.standard_normal((6, 400))
    z[4] = 0.0                                            # a dead component
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = np.corrcoef(z)
    sim32 = np.nan_to_num(ref, nan=0.0).astype(np.float32)
    sim32[0, 1] += 3e-6                                   # the two triangles come from different tile passes
    dead = np.zeros(6, np.int32)
    dead[4] = 1
    sim = _finish_similarity(sim32, dead)
    assert sim.dtype == np.float64 and np.array_equal(sim, sim.T)
    assert np.all(np.diag(sim) == 1.0) and sim.min() >= 0.0 and sim.max() <= 1.0
    assert np.all(sim[4, [0, 1, 2, 3, 5]] == 0.0) and np.all(sim[[0, 1, 2, 3, 5], 4] == 0.0)
    live = [0, 1, 2, 3, 5]
    want = np.clip(ref[np.ix_(live, live)], 0, 1)
    assert np.abs(sim[np.ix_(live, live)] - want).max() < 1e-5
    dis = sim2dis(sim)
    assert np.all(np.diag(dis) == 0.0) and dis.max() <= 1.0


def test_run_seeds_are_pinned_by_the_environment_and_random_otherwise(monkeypatch):
    from nfact_b200.nmfsso import _run_seed
    monkeypatch.setenv("NFB_SSO_SEED", "7")
    a = [_run_seed(r) for r in range(5)]
    assert a == [_run_seed(r) for r in range(5)] and len(set(a)) == 5
    assert all(0 <= s < 2 ** 64 for s in a)
    monkeypatch.setenv("NFB_SSO_SEED", "8")
    assert _run_seed(0) != a[0]
    monkeypatch.delenv("NFB_SSO_SEED")
    assert _run_seed(0) != _run_seed(0)                   # OS entropy, like random_state=None in the reference
