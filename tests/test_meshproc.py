"""Mesh post-processing of the guidance stage (src/foho/guidance/run.py:158-161: FloaterRemover,
DegenerateFaceRemover, FaceReducer).  hy3dgen / MeshLab are not in the reference tree, so these are
known-answer tests of the published behaviour of those filters (manifoldness, boundary, shape, counts)."""
import numpy as np
import pytest

from followmyhold_b200 import meshproc as MP
from followmyhold_b200.meshio import TriMesh
from followmyhold_b200.synthetic import boundary_edges, icosphere, standin_hand_mesh


def edge_face_counts(faces):
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1)
    return np.unique(e, axis=0, return_counts=True)


def signed_volume(v, f):
    t = v[f]
    return float(np.einsum("ij,ij->i", t[:, 0], np.cross(t[:, 1], t[:, 2])).sum() / 6.0)


def test_decimated_sphere_is_a_closed_manifold_sphere():
    v, f = icosphere(5, 1.0)                                   # 20 480 faces
    v = v.astype(np.float64); f = f.astype(np.int32)
    for target in (5000, 1000, 200):
        ov, of = MP.reduce_faces(v, f, target)
        assert target - 1 <= len(of) <= target                 # closed mesh: face count stays even
        assert of.min() == 0 and of.max() == len(ov) - 1 and len(np.unique(of)) == len(ov)
        edges, cnt = edge_face_counts(of)
        assert (cnt == 2).all()                                # closed 2-manifold
        assert len(ov) - len(edges) + len(of) == 2             # Euler characteristic of a sphere
        r = np.linalg.norm(ov, axis=1)
        assert np.abs(r - 1.0).max() < (0.01 if target >= 1000 else 0.06)
        vol = signed_volume(ov, of)
        assert vol > 0 and abs(vol - 4.0 / 3.0 * np.pi) < (0.03 if target >= 1000 else 0.3)   # orientation kept
        t = ov[of]
        n = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0])
        assert (np.einsum("ij,ij->i", n, t.mean(1)) > 0).all()  # no flipped face
        assert np.linalg.norm(n, axis=1).min() > 0
    # determinism
    a = MP.reduce_faces(v, f, 1000); b = MP.reduce_faces(v, f, 1000)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_decimation_keeps_flat_regions_exact_and_the_boundary_loop():
    # a flat square grid: every collapse has zero quadric error, the outline must stay the outline
    n = 41
    x, y = np.meshgrid(np.linspace(0, 1, n), np.linspace(0, 1, n), indexing="ij")
    v = np.stack([x.ravel(), y.ravel(), np.zeros(n * n)], 1)
    idx = np.arange(n * n).reshape(n, n)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    f = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    ov, of = MP.reduce_faces(v, f, 400)
    assert len(of) <= 400
    assert np.abs(ov[:, 2]).max() < 1e-12
    t = ov[of]
    area = 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1).sum()
    assert abs(area - 1.0) < 1e-9                               # the square is still covered exactly
    be = boundary_edges(of)
    on_outline = np.isclose(ov[be.ravel()], 0).any(1) | np.isclose(ov[be.ravel()], 1)[:, :2].any(1)
    assert on_outline.all()
    for corner in ([0, 0], [0, 1], [1, 0], [1, 1]):
        assert np.isclose(ov[:, :2], corner).all(1).any()       # corners survive (boundary planes of two directions)


def test_decimation_of_an_open_hand_mesh_keeps_one_boundary_loop():
    v, f = standin_hand_mesh(0.35)
    v = v.astype(np.float64); f = f.astype(np.int32)
    nb0 = len(boundary_edges(f))
    ov, of = MP.reduce_faces(v, f, 600)
    assert len(of) <= 600 and len(of) >= 590
    edges, cnt = edge_face_counts(of)
    assert cnt.max() == 2 and 3 <= (cnt == 1).sum() <= nb0      # still manifold with a (coarser) wrist loop
    be = boundary_edges(of)
    deg = np.bincount(be.ravel(), minlength=len(ov))
    assert set(np.unique(deg)) <= {0, 2}                        # one simple loop: every boundary vertex has 2 edges
    lo, hi = v.min(0), v.max(0)
    assert (ov >= lo - 0.02).all() and (ov <= hi + 0.02).all()


def test_small_meshes_and_bad_input():
    v, f = icosphere(1, 1.0)
    ov, of = MP.reduce_faces(v, f, 40000)
    assert ov is v and np.array_equal(of, f)                    # already small enough: untouched
    from followmyhold_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    bad = np.array([[0, 1, 9]], np.int32); vv = np.zeros((3, 3))
    nv, nf = C.c_int32(), C.c_int32()
    out_v = np.zeros((3, 3)); out_f = np.zeros((1, 3), np.int32)
    assert lib.foho_mesh_decimate(vv.ctypes.data, 3, bad.ctypes.data, 1, 0, 3.0, out_v.ctypes.data, C.byref(nv),
                                  out_f.ctypes.data, C.byref(nf)) == -4           # FOHO_E_ARG: index out of range
    assert lib.foho_mesh_decimate(None, 3, bad.ctypes.data, 1, 0, 3.0, out_v.ctypes.data, C.byref(nv),
                                  out_f.ctypes.data, C.byref(nf)) == -1


def test_floater_and_degenerate_removers():
    v, f = icosphere(3, 1.0)                                    # 1280 faces
    sv, sf = icosphere(0, 0.05)                                 # a 20-face crumb: 1.6 % of the big one -> stays at 0.5 %
    tv, tf = icosphere(0, 0.02)
    tf = tf[:5]                                                 # a 5-face crumb: 0.4 % -> removed
    V = np.concatenate([v, sv + 3.0, tv - 3.0, [[9.0, 9.0, 9.0]]])
    F = np.concatenate([f, sf + len(v), tf + len(v) + len(sv)])
    ov, of = MP.remove_floaters(V, F)
    assert len(of) == 1280 + 20
    assert len(ov) == len(v) + len(sv)                          # crumb's and the loose vertex are gone
    assert np.allclose(ov[of].reshape(-1, 3).max(0), [3.05, 3.05, 3.05], atol=0.01)
    # two spheres touching in ONE vertex are two components (edge connectivity)
    v2 = np.concatenate([v, v + np.array([2.0, 0, 0])])
    f2 = np.concatenate([f, f + len(v)])
    assert len(np.unique(MP.face_components(f2))) == 2
    # degenerate faces: repeated index, zero area (collinear), plus unreferenced vertices
    dv = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 0, 0], [5, 5, 5.0]])
    df = np.array([[0, 1, 2], [0, 1, 1], [0, 1, 3]])
    cv, cf = MP.remove_degenerate_faces(dv, df)
    assert np.array_equal(cf, [[0, 1, 2]]) and len(cv) == 3
    # the reference's call sequence (run.py:158-161) on TriMesh objects
    m = MP.FaceReducer()(MP.DegenerateFaceRemover()(MP.FloaterRemover()(TriMesh(V, F))), max_facenum=500)
    assert isinstance(m, TriMesh) and len(m.faces) <= 500


def test_decimator_survives_triangle_soups_and_keeps_the_genus():
    """Host C++ inside the stage must not fall over on bad input: random soups (index-degenerate, duplicate and
    non-manifold faces, duplicate positions, planar sets) come back with valid indices; a torus keeps genus 1."""
    rng = np.random.default_rng(0)
    for trial in range(200):
        V = int(rng.integers(4, 60)); F = int(rng.integers(1, 200))
        v = rng.standard_normal((V, 3))
        if trial % 5 == 0:
            v[:, 2] = 0
        if trial % 7 == 0:
            v[rng.integers(V)] = v[rng.integers(V)]
        f = rng.integers(0, V, (F, 3)).astype(np.int32)
        ov, of = MP.reduce_faces(v, f, int(rng.integers(0, F + 1)))
        assert of.ndim == 2 and len(of) <= F and np.isfinite(ov).all()
        assert len(of) == 0 or (of.min() >= 0 and of.max() < len(ov))
        MP.remove_floaters(v, f.astype(np.int64)); MP.remove_degenerate_faces(v, f.astype(np.int64))
    n, m = 48, 24
    u, w = np.meshgrid(np.linspace(0, 2 * np.pi, n, endpoint=False), np.linspace(0, 2 * np.pi, m, endpoint=False), indexing="ij")
    tv = np.stack([(1 + 0.35 * np.cos(w)) * np.cos(u), (1 + 0.35 * np.cos(w)) * np.sin(u), 0.35 * np.sin(w)], -1).reshape(-1, 3)
    idx = np.arange(n * m).reshape(n, m)
    a, b, c, d = idx, np.roll(idx, -1, 0), np.roll(np.roll(idx, -1, 0), -1, 1), np.roll(idx, -1, 1)
    tf = np.concatenate([np.stack([a, b, c], -1).reshape(-1, 3), np.stack([a, c, d], -1).reshape(-1, 3)]).astype(np.int32)
    ov, of = MP.reduce_faces(tv, tf, 600)
    edges, cnt = edge_face_counts(of)
    assert len(of) <= 600 and (cnt == 2).all() and len(ov) - len(edges) + len(of) == 0
