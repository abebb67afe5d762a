"""Dependency-free readers / writers for the files the stages exchange:
PLY (ascii + binary little/big endian; meshes and face-less point clouds), OBJ, binary glTF
(geometry of ``mesh.glb``), and the ``.npy`` 4x4 transforms.  Stands in for ``trimesh.load(process=False)`` /
``mesh.export`` as used by the reference's src/foho/alignment/mesh_align.py:186-187,214.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Union

import numpy as np

_PLY_TYPES = {
    "char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2",
    "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4",
    "float": "f4", "float32": "f4", "double": "f8", "float64": "f8",
}


@dataclass
class TriMesh:
    """Triangle mesh (float64 vertices, int64 faces) -- the role of ``trimesh.Trimesh``."""
    vertices: np.ndarray
    faces: np.ndarray

    def copy(self) -> "TriMesh":
        return TriMesh(self.vertices.copy(), self.faces.copy())

    def apply_transform(self, T: np.ndarray) -> "TriMesh":
        self.vertices = transform_points(self.vertices, T)
        return self

    @property
    def triangles(self) -> np.ndarray:
        return self.vertices[self.faces]

    @property
    def area_faces(self) -> np.ndarray:
        t = self.triangles
        return 0.5 * np.linalg.norm(np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]), axis=1)

    @property
    def area(self) -> float:
        return float(self.area_faces.sum())

    @property
    def centroid(self) -> np.ndarray:
        """Area-weighted mean of the triangle centroids (``trimesh.Trimesh.centroid``)."""
        a = self.area_faces
        c = self.triangles.mean(axis=1)
        return (c * a[:, None]).sum(0) / a.sum()

    @property
    def scale(self) -> float:
        """Length of the AABB diagonal (``trimesh.Trimesh.scale``)."""
        return float(np.linalg.norm(self.vertices.max(0) - self.vertices.min(0)))


@dataclass
class PointCloud:
    """Face-less vertex set -- the role of ``trimesh.PointCloud`` (mesh_align.py:19-22)."""
    vertices: np.ndarray

    def copy(self) -> "PointCloud":
        return PointCloud(self.vertices.copy())

    def apply_transform(self, T: np.ndarray) -> "PointCloud":
        self.vertices = transform_points(self.vertices, T)
        return self


Geometry = Union[TriMesh, PointCloud]


def transform_points(points: np.ndarray, T: np.ndarray) -> np.ndarray:
    """``trimesh.transform_points``: dot(T[:3,:3], p) + T[:3,3]."""
    points = np.asarray(points, dtype=np.float64)
    return points @ T[:3, :3].T + T[:3, 3]


def _read_ply(path: str) -> Geometry:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None
        elements = []          # [(name, count, [(prop_name, dtype | ('list', count_t, item_t))])]
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").strip().split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append((tok[1], int(tok[2]), []))
            elif tok[0] == "property":
                if tok[1] == "list":
                    elements[-1][2].append((tok[4], ("list", _PLY_TYPES[tok[2]], _PLY_TYPES[tok[3]])))
                else:
                    elements[-1][2].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        verts = None
        faces = None
        if fmt == "ascii":
            rest = f.read().decode("ascii", "replace").split()
            pos = 0
            for name, count, props in elements:
                if name == "vertex":
                    n = len(props)
                    arr = np.array(rest[pos:pos + n * count], dtype=np.float64).reshape(count, n)
                    pos += n * count
                    names = [p[0] for p in props]
                    verts = arr[:, [names.index("x"), names.index("y"), names.index("z")]]
                elif name == "face":
                    fl = []
                    for _ in range(count):
                        k = int(rest[pos]); idx = [int(v) for v in rest[pos + 1:pos + 1 + k]]; pos += 1 + k
                        for t in range(1, k - 1):
                            fl.append((idx[0], idx[t], idx[t + 1]))
                    faces = np.asarray(fl, dtype=np.int64).reshape(-1, 3)
                else:
                    for _ in range(count):
                        for pn, pt in props:
                            if isinstance(pt, tuple):
                                k = int(rest[pos]); pos += 1 + k
                            else:
                                pos += 1
        else:
            end = "<" if fmt == "binary_little_endian" else ">"
            for name, count, props in elements:
                if all(not isinstance(pt, tuple) for _, pt in props):
                    dt = np.dtype([(pn, end + pt) for pn, pt in props])
                    arr = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count)
                    if name == "vertex":
                        verts = np.stack([arr["x"], arr["y"], arr["z"]], 1).astype(np.float64)
                elif name == "face" and len(props) == 1:
                    _, (_, ct, it) = props[0]
                    cdt, idt = np.dtype(end + ct), np.dtype(end + it)
                    raw = f.read()
                    # fast path: all triangles
                    rec = np.dtype([("n", cdt), ("v", idt, (3,))])
                    if len(raw) >= rec.itemsize * count:
                        arr = np.frombuffer(raw[:rec.itemsize * count], dtype=rec, count=count)
                        if (arr["n"] == 3).all():
                            faces = arr["v"].astype(np.int64)
                            f.seek(-(len(raw) - rec.itemsize * count), os.SEEK_CUR)
                            continue
                    fl = []
                    pos = 0
                    for _ in range(count):
                        k = int(np.frombuffer(raw, dtype=cdt, count=1, offset=pos)[0]); pos += cdt.itemsize
                        idx = np.frombuffer(raw, dtype=idt, count=k, offset=pos); pos += idt.itemsize * k
                        for t in range(1, k - 1):
                            fl.append((int(idx[0]), int(idx[t]), int(idx[t + 1])))
                    faces = np.asarray(fl, dtype=np.int64).reshape(-1, 3)
                    f.seek(-(len(raw) - pos), os.SEEK_CUR)
                else:
                    raise ValueError(f"{path}: unsupported PLY element '{name}' with list properties")
    if verts is None:
        raise ValueError(f"{path}: PLY without vertices")
    if faces is None or len(faces) == 0:
        return PointCloud(verts)
    return TriMesh(verts, faces)


def _read_obj(path: str) -> Geometry:
    vs, fs = [], []
    with open(path, "r") as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                vs.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = [int(tok.split("/")[0]) for tok in line.split()[1:]]
                idx = [i - 1 if i > 0 else len(vs) + i for i in idx]
                for t in range(1, len(idx) - 1):
                    fs.append((idx[0], idx[t], idx[t + 1]))
    v = np.asarray(vs, dtype=np.float64).reshape(-1, 3)
    if not fs:
        return PointCloud(v)
    return TriMesh(v, np.asarray(fs, dtype=np.int64))


_GLTF_DTYPES = {5120: "i1", 5121: "u1", 5122: "<i2", 5123: "<u2", 5125: "<u4", 5126: "<f4"}
_GLTF_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT2": 4, "MAT3": 9, "MAT4": 16}


def _read_glb(path: str) -> Geometry:
    """Binary glTF 2.0 (what MoGe's ``save_glb`` writes as ``mesh.glb``, src/foho/geometry/moge.py:159-160;
    the reference reads it with trimesh / pytorch3d's MeshGlbFormat, alignment/h2m.py:27-31,
    pipelines.py:1247-1250): every triangle primitive of the default scene, node transforms applied,
    concatenated into one mesh (like ``trimesh.Scene.dump(concatenate=True)``).  Geometry only: positions
    and indices; normals, colours, textures are skipped."""
    import json
    import struct
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 20 or data[:4] != b"glTF":
        raise ValueError(f"{path}: not a binary glTF file")
    version, length = struct.unpack_from("<II", data, 4)
    if version != 2:
        raise ValueError(f"{path}: glTF version {version} is not supported")
    off, gltf, blob = 12, None, b""
    while off + 8 <= min(length, len(data)):
        clen, ctype = struct.unpack_from("<II", data, off)
        chunk = data[off + 8: off + 8 + clen]
        if ctype == 0x4E4F534A:
            gltf = json.loads(chunk.decode("utf-8"))
        elif ctype == 0x004E4942 and not blob:
            blob = chunk
        off += 8 + clen + (-clen) % 4
    if gltf is None:
        raise ValueError(f"{path}: no JSON chunk")

    def accessor(i: int) -> np.ndarray:
        a = gltf["accessors"][i]
        ncomp, dt = _GLTF_NCOMP[a["type"]], np.dtype(_GLTF_DTYPES[a["componentType"]])
        count = int(a["count"])
        if "bufferView" not in a:
            return np.zeros((count, ncomp), dtype=dt)
        bv = gltf["bufferViews"][a["bufferView"]]
        if bv.get("buffer", 0) != 0:
            raise ValueError(f"{path}: external buffers are not supported")
        start = int(bv.get("byteOffset", 0)) + int(a.get("byteOffset", 0))
        stride = int(bv.get("byteStride", 0)) or ncomp * dt.itemsize
        raw = np.frombuffer(blob, dtype=np.uint8, count=(count - 1) * stride + ncomp * dt.itemsize, offset=start) if count else np.zeros(0, np.uint8)
        out = np.lib.stride_tricks.as_strided(raw, shape=(count, ncomp * dt.itemsize), strides=(stride, 1)) if count else raw.reshape(0, ncomp * dt.itemsize)
        return np.ascontiguousarray(out).view(dt).reshape(count, ncomp)

    def node_matrix(n: dict) -> np.ndarray:
        if "matrix" in n:
            return np.asarray(n["matrix"], dtype=np.float64).reshape(4, 4).T          # column-major
        M = np.eye(4)
        if "scale" in n:
            M = np.diag(list(n["scale"]) + [1.0]) @ M
        if "rotation" in n:                                                             # xyzw
            x, y, z, w = [float(c) for c in n["rotation"]]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            Rm = np.eye(4); Rm[:3, :3] = R
            M = Rm @ M
        if "translation" in n:
            Tm = np.eye(4); Tm[:3, 3] = n["translation"]
            M = Tm @ M
        return M

    verts, faces, nv = [], [], 0
    nodes = gltf.get("nodes", [])
    scenes = gltf.get("scenes", [])
    roots = scenes[gltf.get("scene", 0)].get("nodes", []) if scenes else list(range(len(nodes)))
    stack = [(r, np.eye(4)) for r in roots]
    if not nodes:                                    # no scene graph: take the meshes as they are
        stack = []
        for m in gltf.get("meshes", []):
            nodes.append({"mesh": gltf["meshes"].index(m)})
            stack.append((len(nodes) - 1, np.eye(4)))
    while stack:
        i, parent = stack.pop()
        n = nodes[i]
        M = parent @ node_matrix(n)
        stack.extend((c, M) for c in n.get("children", []))
        if "mesh" not in n:
            continue
        for prim in gltf["meshes"][n["mesh"]].get("primitives", []):
            if "POSITION" not in prim.get("attributes", {}):
                continue
            v = transform_points(accessor(prim["attributes"]["POSITION"]).astype(np.float64), M)
            mode = prim.get("mode", 4)
            if mode == 4:
                idx = accessor(prim["indices"]).astype(np.int64).reshape(-1) if "indices" in prim else np.arange(len(v))
                faces.append(idx[: len(idx) // 3 * 3].reshape(-1, 3) + nv)
            elif mode != 0:
                raise ValueError(f"{path}: primitive mode {mode} is not supported (triangles and points only)")
            verts.append(v)
            nv += len(v)
    if not verts:
        raise ValueError(f"{path}: no geometry")
    V = np.concatenate(verts, 0)
    F = np.concatenate(faces, 0) if faces else np.zeros((0, 3), dtype=np.int64)
    return TriMesh(V, F) if len(F) else PointCloud(V)


def load(path: str) -> Geometry:
    """``trimesh.load(path, process=False)`` for the formats the stages exchange."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".ply":
        return _read_ply(path)
    if ext == ".obj":
        return _read_obj(path)
    if ext == ".glb":
        return _read_glb(path)
    raise ValueError(f"unsupported mesh format '{ext}' ({path})")


def write_ply(path: str, vertices: np.ndarray, faces: Optional[np.ndarray] = None, double: bool = False) -> None:
    """Binary little-endian PLY (float32 vertices, int32 faces) like ``trimesh`` exports;
    ``double=True`` keeps float64 coordinates (``property double``)."""
    v = np.ascontiguousarray(vertices, dtype="<f8" if double else "<f4")
    t = "double" if double else "float"
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {len(v)}",
              f"property {t} x", f"property {t} y", f"property {t} z"]
    if faces is not None and len(faces):
        header += [f"element face {len(faces)}", "property list uchar int vertex_indices"]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(v.tobytes())
        if faces is not None and len(faces):
            rec = np.zeros(len(faces), dtype=[("n", "u1"), ("v", "<i4", (3,))])
            rec["n"] = 3
            rec["v"] = np.asarray(faces, dtype=np.int32)
            f.write(rec.tobytes())


def write_obj(path: str, vertices: np.ndarray, faces: Optional[np.ndarray] = None) -> None:
    with open(path, "w") as f:
        for p in np.asarray(vertices):
            f.write(f"v {p[0]:.8f} {p[1]:.8f} {p[2]:.8f}\n")
        if faces is not None:
            for t in np.asarray(faces):
                f.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")


def export(geom: Geometry, path: str) -> None:
    ext = os.path.splitext(path)[1].lower()
    faces = geom.faces if isinstance(geom, TriMesh) else None
    if ext == ".ply":
        write_ply(path, geom.vertices, faces)
    elif ext == ".obj":
        write_obj(path, geom.vertices, faces)
    else:
        raise ValueError(f"unsupported export format '{ext}'")
