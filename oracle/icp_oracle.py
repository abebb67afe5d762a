"""CPU ORACLE for the alignment path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates ``icp`` of the reference's src/foho/alignment/mesh_align.py:56-175 on plain
point sets, using ``scipy.spatial.cKDTree`` exactly as the reference does (:89,111) and a
restatement of ``trimesh.registration.procrustes`` (trimesh is absent offline; semantics
from memory -- SURVEY.md Appendix C: centre on means, scale = ratio of RMS radii,
R = U diag(1,1,det(U V^T)) V^T, t = b_mean - s R a_mean).

PARITY UNPINNED by the reference (no tests / fixtures); pinned here by known-answer tests
(tests/test_oracle_icp.py): recovery of a known similarity with 20 % outliers,
``scipy.linalg.orthogonal_procrustes`` cross-check, argsort trim semantics.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def transform_points(points, T):
    return points @ T[:3, :3].T + T[:3, 3]


def procrustes(a, b, reflection=False, scale=True):
    """trimesh.registration.procrustes(a, b, reflection, translation=True, scale, return_cost=False)."""
    a = np.asanyarray(a, dtype=np.float64)
    b = np.asanyarray(b, dtype=np.float64)
    acenter = a.mean(axis=0)
    bcenter = b.mean(axis=0)
    ac = a - acenter
    bc = b - bcenter
    if scale:
        ascale = np.sqrt((ac ** 2).sum() / len(a))
        bscale = np.sqrt((bc ** 2).sum() / len(b))
    else:
        ascale = bscale = 1.0
    u, s, vh = np.linalg.svd(np.dot((bc / bscale).T, ac / ascale))
    if reflection:
        R = u @ vh
    else:
        R = u @ np.diag([1, 1, np.linalg.det(u @ vh)]) @ vh
    t = bcenter - (bscale / ascale) * (R @ acenter)
    M = np.eye(4)
    M[:3, :3] = (bscale / ascale) * R
    M[:3, 3] = t
    return M


def icp_points(source_points, target_points, n_iter, n_outliers, fixed_scale=False, min_scale=0.5, max_scale=2.0,
               start=None, return_history=False):
    """mesh_align.py:89-142 for one start transform ("cube")."""
    source_points = np.asarray(source_points, dtype=np.float64)
    target_points = np.asarray(target_points, dtype=np.float64)
    kdtree = cKDTree(target_points)
    transform = np.eye(4) if start is None else start.copy()
    best_cost = np.inf
    best_transform = transform.copy()
    hist = []
    qi = None
    for _ in range(n_iter):
        p = transform_points(source_points, transform)
        dist, qi = kdtree.query(p)
        q = target_points[qi]
        if n_outliers > 0:
            order = np.argsort(dist)
            inl = order[:-n_outliers]
            cost = dist[inl].mean()
            p_in, q_in = p[inl], q[inl]
        else:
            p_in, q_in = p, q
            cost = dist.mean()
        nxt = procrustes(p_in, q_in, reflection=False, scale=not fixed_scale)
        transform = nxt @ transform
        if not fixed_scale:
            sc = np.linalg.norm(transform[:3, 0])
            transform[:3, :3] /= sc
            sc = np.clip(sc, min_scale, max_scale)
            transform[:3, :3] *= sc
        hist.append(cost)
        if cost < best_cost:
            best_cost = cost
            best_transform = transform
    if return_history:
        return best_transform, best_cost, np.asarray(hist), qi
    return best_transform, best_cost


def init_transform_points(src, tgt, fixed_scale=False):
    """mesh_align.py:18-35 for two point clouds: vertex-mean centroids, bbox-diagonal scales,
    T(translation) @ S(scale about the source centroid)."""
    sc, tc = src.mean(axis=0), tgt.mean(axis=0)
    ss = np.linalg.norm(src.max(axis=0) - src.min(axis=0))
    ts = np.linalg.norm(tgt.max(axis=0) - tgt.min(axis=0))
    T = np.eye(4)
    T[:3, 3] = tc - sc
    if fixed_scale:
        return T
    f = ts / ss
    S = np.diag([f, f, f, 1.0])
    S[:3, 3] = sc * (1.0 - f)
    return T @ S


def align_points(src, tgt, outliers=0.2, iterations_coarse=50, iterations_fine=100, min_scale=0.7, max_scale=3.0,
                 fixed_scale=False):
    """mesh_align.py:178-217 for two point clouds (the sample counts collapse to the cloud sizes,
    :75-83): init, coarse ICP, fine ICP (which keeps icp()'s own fixed_scale=False default, :201-204),
    final = fine @ coarse @ init.  Returns (final [4,4], transformed source points)."""
    src = np.asarray(src, dtype=np.float64)
    tgt = np.asarray(tgt, dtype=np.float64)
    n_out = int(outliers * len(src))
    init = init_transform_points(src, tgt, fixed_scale)
    p = transform_points(src, init)
    coarse, _ = icp_points(p, tgt, iterations_coarse, n_out, fixed_scale, min_scale, max_scale)
    p = transform_points(p, coarse)
    fine, _ = icp_points(p, tgt, iterations_fine, n_out, False, min_scale, max_scale)
    p = transform_points(p, fine)
    return fine @ coarse @ init, p
