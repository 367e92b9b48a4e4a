"""Known-answer tests of the nnmf restatement (oracle/nmf_bg.py): rank-1 ALS converges to the dominant singular pair of a
non-negative matrix; exact non-negative rank-k data is reproduced; output conventions (unit rows of h, energy order)."""
import numpy as np


def test_rank1_matches_dominant_singular_pair():
    from oracle.nmf_bg import nnmf_als
    rng = np.random.default_rng(0)
    a = np.outer(rng.uniform(1, 2, 60), rng.uniform(1, 2, 400)) * 50 + rng.uniform(0, 1, (60, 400))
    w, h, it = nnmf_als(a, 1)
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    assert it < 100 and abs(np.linalg.norm(h) - 1) < 1e-12
    best = s[0] * np.outer(u[:, 0], vt[0])
    assert np.abs(np.outer(w[:, 0], h[0]) - best).max() < 1e-2 * np.abs(best).max()      # stopping tolerance of nnmf is 1e-4 (loose)


def test_exact_rank2_reproduced_and_ordered():
    from oracle.nmf_bg import nnmf_als, hash_uniform
    rng = np.random.default_rng(1)
    W = rng.uniform(0, 1, (40, 2)); W[:20, 0] = 0; W[20:, 1] = 0        # disjoint supports: the factorisation is unique
    H = rng.uniform(0, 1, (2, 300)); H[1] *= 3
    w, h, it = nnmf_als(W @ H, 2)
    assert np.abs(w @ h - W @ H).max() < 1e-2 * np.abs(W @ H).max()
    assert np.allclose(np.sum(h * h, axis=1), 1.0) and np.sum(w[:, 0] ** 2) >= np.sum(w[:, 1] ** 2)
    u = hash_uniform(np.arange(1000))
    assert u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.05
