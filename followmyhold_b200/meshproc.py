"""Mesh post-processing of the guidance stage: the three hy3dgen post-processors the reference applies to
the extracted object mesh before export (src/foho/guidance/run.py:158-161)::

    obj_mesh = FloaterRemover()(obj_mesh)
    obj_mesh = DegenerateFaceRemover()(obj_mesh)
    obj_mesh = FaceReducer()(obj_mesh)

hy3dgen (Hunyuan3D-2 @ e664e74, ``hy3dgen/shapegen/postprocessors.py``) and the MeshLab filters it calls are
not in the reference tree, so the behaviour restated here is the published one of those filters ("parity
unpinned"):

* ``FloaterRemover``  -> ``compute_selection_by_small_disconnected_components_per_face(nbfaceratio=0.005)`` +
  ``meshing_remove_selected_vertices_and_faces``: drop every edge-connected component whose face count is
  below 0.5 % of the largest component's;
* ``DegenerateFaceRemover`` -> a save / re-load round trip through a PLY file; here: faces with a repeated
  vertex index or zero area, and vertices no face references, are dropped;
* ``FaceReducer`` -> ``meshing_decimation_quadric_edge_collapse(targetfacenum=40000, preserveboundary,
  boundaryweight=3, preservenormal, preservetopology)``: ``foho_mesh_decimate`` (csrc/mesh_decimate.cpp, host
  code in the C-ABI library).
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _lib
from .meshio import TriMesh


def _compact(verts: np.ndarray, faces: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Drop unreferenced vertices, keep the order of the rest."""
    used = np.zeros(len(verts), dtype=bool)
    used[faces.reshape(-1)] = True
    remap = np.cumsum(used) - 1
    return verts[used], remap[faces].astype(faces.dtype)


def face_components(faces: np.ndarray) -> np.ndarray:
    """Label of the edge-connected component of every face (faces sharing an EDGE are connected -- the
    face-face adjacency MeshLab's component filters use; a shared vertex alone does not connect)."""
    F = len(faces)
    if F == 0:
        return np.zeros(0, dtype=np.int64)
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    e = np.sort(np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]), axis=1).astype(np.int64)
    key = e[:, 0] * (int(faces.max()) + 1) + e[:, 1]
    fid = np.tile(np.arange(F), 3)
    order = np.argsort(key, kind="stable")
    key, fid = key[order], fid[order]
    same = key[1:] == key[:-1]                      # consecutive entries of one edge: link their faces
    a, b = fid[:-1][same], fid[1:][same]
    g = coo_matrix((np.ones(len(a), dtype=np.int8), (a, b)), shape=(F, F))
    return connected_components(g, directed=False)[1]


def remove_floaters(verts: np.ndarray, faces: np.ndarray, nbfaceratio: float = 0.005) -> Tuple[np.ndarray, np.ndarray]:
    faces = np.asarray(faces)
    if len(faces) == 0:
        return verts, faces
    lab = face_components(faces)
    cnt = np.bincount(lab)
    small = cnt < nbfaceratio * cnt.max()
    keep = ~small[lab]
    return _compact(np.asarray(verts), faces[keep])


def remove_degenerate_faces(verts: np.ndarray, faces: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    verts, faces = np.asarray(verts), np.asarray(faces)
    if len(faces) == 0:
        return verts[:0], faces
    ok = (faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])
    tri = verts[faces].astype(np.float64)
    area2 = np.linalg.norm(np.cross(tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]), axis=1)
    ok &= area2 > 0.0
    return _compact(verts, faces[ok])


def reduce_faces(verts: np.ndarray, faces: np.ndarray, max_facenum: int = 40000,
                 boundary_weight: float = 3.0) -> Tuple[np.ndarray, np.ndarray]:
    """Quadric edge-collapse decimation to at most ``max_facenum`` faces (a mesh that is already small enough
    comes back as it is, like hy3dgen's ``reduce_face``)."""
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    if len(faces) <= max_facenum:
        return verts, faces
    lib = _lib.load()
    v = np.ascontiguousarray(verts, dtype=np.float64)
    ov = np.empty_like(v); of = np.empty_like(faces)
    nv, nf = C.c_int32(0), C.c_int32(0)
    _lib.check("foho_mesh_decimate", lib.foho_mesh_decimate(
        v.ctypes.data, len(v), faces.ctypes.data, len(faces), int(max_facenum), float(boundary_weight),
        ov.ctypes.data, C.byref(nv), of.ctypes.data, C.byref(nf)))
    return ov[:nv.value].astype(np.asarray(verts).dtype, copy=False), of[:nf.value]


class FloaterRemover:
    def __call__(self, mesh: TriMesh) -> TriMesh:
        return TriMesh(*remove_floaters(mesh.vertices, mesh.faces))


class DegenerateFaceRemover:
    def __call__(self, mesh: TriMesh) -> TriMesh:
        return TriMesh(*remove_degenerate_faces(mesh.vertices, mesh.faces))


class FaceReducer:
    def __call__(self, mesh: TriMesh, max_facenum: int = 40000) -> TriMesh:
        return TriMesh(*reduce_faces(mesh.vertices, mesh.faces, max_facenum))
