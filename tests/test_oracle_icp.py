"""Known-answer tests pinning the ICP oracle."""
import numpy as np
from scipy.linalg import orthogonal_procrustes

from followmyhold_b200.synthetic import random_similarity
from oracle import icp_oracle as O


def test_procrustes_recovers_similarity_and_matches_scipy_rotation():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(500, 3))
    T = random_similarity(3, (0.7, 3.0), 0.5)
    b = O.transform_points(a, T)
    M = O.procrustes(a, b, reflection=False, scale=True)
    assert np.allclose(M, T, atol=1e-10)
    # rotation part agrees with scipy's orthogonal Procrustes on the centred, scale-normalised sets
    ac = a - a.mean(0); bc = b - b.mean(0)
    R, _ = orthogonal_procrustes(ac / np.sqrt((ac ** 2).sum() / len(a)), bc / np.sqrt((bc ** 2).sum() / len(b)))
    assert np.allclose(R.T, M[:3, :3] / np.linalg.norm(M[:3, 0]), atol=1e-10)
    # scale is the ratio of RMS radii even with noise (not the Umeyama trace form)
    bn = b + 0.05 * rng.normal(size=b.shape)
    Mn = O.procrustes(a, bn)
    bnc = bn - bn.mean(0)
    assert abs(np.linalg.norm(Mn[:3, 0]) - np.sqrt((bnc ** 2).sum() / (ac ** 2).sum())) < 1e-12
    # no reflection even for mirrored data
    Mm = O.procrustes(a, a * np.array([1, 1, -1.0]), reflection=False)
    assert np.linalg.det(Mm[:3, :3]) > 0


def test_icp_recovers_known_similarity_with_outliers():
    rng = np.random.default_rng(1)
    tgt = rng.normal(size=(4000, 3)) * np.array([1.0, 0.6, 0.3])
    T = random_similarity(11, (0.9, 1.2), 0.05)
    T[:3, :3] = 0.1 * T[:3, :3] + 0.9 * np.linalg.norm(T[:3, 0]) * np.eye(3)
    u, s, vh = np.linalg.svd(T[:3, :3]); T[:3, :3] = (u @ vh) * s.mean()
    src = O.transform_points(tgt[:1500], np.linalg.inv(T))
    src[:300] += rng.normal(scale=1.0, size=(300, 3))           # 20 % outliers
    Tb, cost = O.icp_points(src, tgt, 60, int(0.2 * 1500), False, 0.7, 3.0)
    assert np.allclose(Tb, T, atol=1e-6), np.abs(Tb - T).max()
    assert cost < 1e-6


def test_trim_semantics_and_best_cost_pairing():
    rng = np.random.default_rng(2)
    tgt = rng.normal(size=(800, 3)); src = tgt[:200] * 1.05 + 0.02
    Tb, cb, hist, qi = O.icp_points(src, tgt, 8, 40, False, 0.5, 2.0, return_history=True)
    # cost_k is the trimmed mean *before* update k; the best transform is the one produced in the
    # iteration whose pre-update cost was lowest (mesh_align.py:140-142)
    assert cb == hist.min()
    T = np.eye(4)
    from scipy.spatial import cKDTree
    tree = cKDTree(tgt)
    for k in range(int(hist.argmin()) + 1):
        p = O.transform_points(src, T)
        d, q = tree.query(p)
        order = np.argsort(d)[:-40]
        assert abs(d[order].mean() - hist[k]) < 1e-14
        T = O.procrustes(p[order], tgt[q][order]) @ T
        sc = np.linalg.norm(T[:3, 0]); T[:3, :3] *= np.clip(sc, 0.5, 2.0) / sc
    assert np.allclose(T, Tb, atol=1e-12)
