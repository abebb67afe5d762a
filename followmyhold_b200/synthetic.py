"""Synthetic stand-ins for the inputs of the guidance / alignment hot path.

Nothing here is on the hot path: these builders only manufacture inputs with the
shapes, dtypes and coordinate frames the reference's stages exchange, because the real
assets are absent from the build box (SURVEY.md §8c "Assets absent"):

* ``MANO_RIGHT.pkl`` is licensed (reference ``README.md:82-86``) -> ``standin_hand_mesh``
  builds an open triangle mesh with MANO's exact topology counts (778 vertices,
  1538 faces, one boundary loop of 16 edges -- a topological disk, like MANO's
  open wrist).
* ``J_regressor_hamer.pt`` is generated at run time by the reference
  (``src/foho/hand/hamer.py:102-104``) -> ``standin_j_regressor``.
* Hunyuan3D decoder output (65^3 logits, negated so negative = inside,
  ``third_party_patches/hy3dgen/shapegen/pipelines.py:309-312``) -> ``ellipsoid_volume``.
* MoGe partial point cloud (``src/foho/geometry/moge.py:156-158``) -> ``partial_cloud``.

Everything is seeded and generated with CPU generators so the same seed gives the
same inputs on the build box and on the GPU box.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

GRID_BOUND = 1.10  # pipelines.py:1127 (bounds = 1.10)
MANO_NUM_VERTS = 778
MANO_NUM_FACES = 1538
MANO_BOUNDARY_EDGES = 16
MANO_FINGERTIP_VERTS = (744, 320, 443, 554, 671)  # pipelines.py:127
MANO_TO_OPENPOSE = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)  # pipelines.py:128


def standin_hand_mesh(length: float = 0.35) -> Tuple[np.ndarray, np.ndarray]:
    """Open 'mitten' mesh with MANO's counts: 778 verts, 1538 faces, 16 boundary edges.

    Construction: 48 rings of 16 vertices along a gently bent axis (ring 0 is the
    open wrist loop), one ring of 8 vertices and 2 tip vertices closing the far end.
    Faces: 47*32 (ring to ring) + 24 (16-ring to 8-ring) + 10 (octagon with two
    interior vertices) = 1538.  Returned verts are float32 [778,3] centred near the
    origin with the long axis along +x and total length ``length``; faces int32
    [1538,3] with outward orientation.
    """
    n_rings, n_seg = 48, 16
    verts = []
    ts = np.linspace(0.0, 0.93, n_rings)

    def axis_point(t):
        # gentle bend in the x-z plane so the shape is not a surface of revolution
        return np.array([t, 0.04 * math.sin(2.2 * t), 0.10 * t * t])

    def radii(t):
        # wrist narrow, palm wide and flat, fingers tapering
        wide = 0.10 + 0.16 * math.exp(-((t - 0.38) / 0.30) ** 2) - 0.05 * t
        thick = 0.055 + 0.035 * math.exp(-((t - 0.30) / 0.35) ** 2) - 0.03 * t
        return wide, thick

    for t in ts:
        c = axis_point(t)
        a, b = radii(t)
        for k in range(n_seg):
            ang = 2.0 * math.pi * k / n_seg
            # a little ripple so no two triangles are coplanar / axis aligned
            rip = 1.0 + 0.06 * math.sin(3.0 * ang + 9.0 * t)
            verts.append(c + np.array([0.0, a * rip * math.cos(ang), b * rip * math.sin(ang)]))
    # ring of 8
    t8 = 0.975
    c = axis_point(t8)
    a, b = radii(t8)
    for k in range(8):
        ang = 2.0 * math.pi * (k + 0.25) / 8
        verts.append(c + 0.55 * np.array([0.0, a * math.cos(ang), b * math.sin(ang)]))
    # two tip verts
    c = axis_point(1.0)
    verts.append(c + np.array([0.0, 0.012, 0.002]))
    verts.append(c + np.array([0.0, -0.012, -0.002]))
    verts = np.asarray(verts, dtype=np.float64)
    assert verts.shape[0] == MANO_NUM_VERTS

    faces = []
    for r in range(n_rings - 1):
        for k in range(n_seg):
            a0 = r * n_seg + k
            a1 = r * n_seg + (k + 1) % n_seg
            b0 = (r + 1) * n_seg + k
            b1 = (r + 1) * n_seg + (k + 1) % n_seg
            faces.append((a0, a1, b1))
            faces.append((a0, b1, b0))
    base16 = (n_rings - 1) * n_seg
    base8 = n_rings * n_seg
    # 16-ring -> 8-ring: 24 triangles
    for k in range(8):
        o0 = base16 + 2 * k
        o1 = base16 + (2 * k + 1) % 16
        o2 = base16 + (2 * k + 2) % 16
        i0 = base8 + k
        i1 = base8 + (k + 1) % 8
        faces.append((o0, o1, i0))
        faces.append((o1, o2, i1))
        faces.append((o1, i1, i0))
    # octagon + two interior verts p (near k=0..3 side) and q: 10 triangles
    p, q = base8 + 8, base8 + 9
    ring = [base8 + k for k in range(8)]
    # p takes ring 0..4, q takes ring 4..8(=0)
    for k in range(0, 4):
        faces.append((ring[k], ring[k + 1], p))
    for k in range(4, 8):
        faces.append((ring[k], ring[(k + 1) % 8], q))
    faces.append((ring[4], q, p))
    faces.append((ring[0], p, q))
    faces = np.asarray(faces, dtype=np.int32)
    assert faces.shape[0] == MANO_NUM_FACES, faces.shape

    # scale to requested length, centre on bbox centre
    lo, hi = verts.min(0), verts.max(0)
    verts = (verts - 0.5 * (lo + hi)) * (length / (hi - lo).max())
    return verts.astype(np.float32), faces


def boundary_edges(faces: np.ndarray) -> np.ndarray:
    """Edges used by exactly one face (MANO's wrist loop has 16)."""
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    e = np.sort(e, axis=1)
    uniq, cnt = np.unique(e, axis=0, return_counts=True)
    return uniq[cnt == 1]


def cap_boundary_loops(faces: np.ndarray) -> np.ndarray:
    """Close every boundary loop with a triangle fan over its own vertices (no new
    vertices): MANO's 16-edge wrist loop gains 14 faces (1538 -> 1552).

    The inside/outside rule is ray parity; on an open mesh every column through the hole
    is 'inside' all the way down, so the guidance engine voxelises the *capped* topology
    (DESIGN.md "sign rule").  Orientation of the fan is opposite to the loop so the cap
    is consistently oriented with the rest of the surface."""
    faces = np.asarray(faces)
    directed = {}
    for a, b, c in faces:
        for u, v in ((a, b), (b, c), (c, a)):
            directed[(int(u), int(v))] = True
    nxt = {u: v for (u, v) in directed if (v, u) not in directed}   # boundary edges follow face orientation
    extra = []
    seen = set()
    for start in list(nxt):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        cur = nxt[start]
        while cur != start and cur in nxt and cur not in seen:
            loop.append(cur); seen.add(cur); cur = nxt[cur]
        if cur != start or len(loop) < 3:
            continue
        for i in range(1, len(loop) - 1):
            extra.append((loop[0], loop[i + 1], loop[i]))
    if not extra:
        return faces.astype(np.int32)
    return np.concatenate([faces, np.asarray(extra, dtype=faces.dtype)], 0).astype(np.int32)


def standin_j_regressor(seed: int = 0) -> np.ndarray:
    """Sparse convex 16x778 joint regressor (rows sum to 1), like MANO's."""
    rng = np.random.default_rng(seed)
    J = np.zeros((16, MANO_NUM_VERTS), dtype=np.float32)
    for j in range(16):
        idx = rng.choice(MANO_NUM_VERTS, size=24, replace=False)
        w = rng.random(24).astype(np.float32)
        J[j, idx] = w / w.sum()
    return J


def random_quaternion(gen: torch.Generator) -> torch.Tensor:
    q = torch.randn(4, generator=gen, dtype=torch.float64)
    return (q / q.norm()).to(torch.float32)


def quat_to_mat_np(q) -> np.ndarray:
    r, i, j, k = [float(x) for x in q]
    two_s = 2.0 / (r * r + i * i + j * j + k * k)
    return np.array([
        [1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r)],
        [two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r)],
        [two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)],
    ], dtype=np.float64)


def ellipsoid_volume(D: int, seed: int, device="cpu", noise: float = 0.02,
                     radii: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[D,D,D] fp32 field ``|x/r| - 1`` + smooth noise on linspace(-1.10,1.10,D)^3.

    Negative inside (the sign convention after pipelines.py:311-312); index order
    [ix,iy,iz], z fastest (pipelines.py:352-357 with indexing="ij").  Returns
    (volume, radii).
    """
    gen = torch.Generator().manual_seed(seed)
    if radii is None:
        radii = 0.3 + 0.4 * torch.rand(3, generator=gen)
    low = torch.randn(1, 1, 12, 12, 12, generator=gen)
    lin = torch.linspace(-GRID_BOUND, GRID_BOUND, D, dtype=torch.float32, device=device)
    r = radii.to(device)
    x = (lin / r[0]).view(D, 1, 1)
    y = (lin / r[1]).view(1, D, 1)
    z = (lin / r[2]).view(1, 1, D)
    vol = torch.sqrt(x * x + y * y + z * z) - 1.0
    up = torch.nn.functional.interpolate(low.to(device), size=(D, D, D), mode="trilinear", align_corners=True)
    vol = vol + noise * up[0, 0]
    return vol.contiguous(), radii


@dataclass
class GuidanceSample:
    """One synthetic guidance problem (all float32 CPU tensors unless noted)."""
    sdf: torch.Tensor          # [D,D,D]
    hand_rest: torch.Tensor    # [778,3]  aligned-MANO verts in MoGe space (pipelines.py:1241)
    hand_faces: torch.Tensor   # [1538,3] int32
    cloud: torch.Tensor        # [P,3]    MoGe-space partial cloud
    T_h2m: torch.Tensor        # [4,4]    Hunyuan -> MoGe similarity (alignment/h2m.py output)
    obj_center: torch.Tensor   # [3]      centre the object similarity acts about (MoGe space)
    theta_h: torch.Tensor      # [8]      s(1) t(3) q(4, wxyz)   hand leaves
    theta_o: torch.Tensor      # [8]      object leaves
    j_regressor: torch.Tensor  # [16,778]
    kps_2d: torch.Tensor       # [21,2]   pixel coordinates in the 512x512 crop
    fov_deg: float
    image_hw: Tuple[int, int]


def make_guidance_sample(D: int, P: int, seed: int, perturb: float = 1.0,
                         device="cpu") -> GuidanceSample:
    """Synthetic sample following SURVEY.md §8d config 1/3.

    The hand is placed (in Hunyuan space) with a random similarity (scale
    U(0.8,1.2), uniform quaternion, translation U(-0.2,0.2)^3) so that it grazes /
    partly penetrates the ellipsoid, then carried to MoGe space by ``T_h2m``.  The
    leaves ``theta_h``/``theta_o`` start at identity + ``perturb`` * small noise so
    every derivative path is exercised.
    """
    gen = torch.Generator().manual_seed(1000 + seed)
    sdf, radii = ellipsoid_volume(D, seed, device="cpu")
    hv, hf = standin_hand_mesh(0.35)
    # base placement in Hunyuan space
    sc = 0.8 + 0.4 * torch.rand(1, generator=gen).item()
    q = random_quaternion(gen)
    R = quat_to_mat_np(q.numpy())
    tr = (torch.rand(3, generator=gen, dtype=torch.float64) * 0.4 - 0.2).numpy()
    # push the hand to the ellipsoid surface along a random direction
    d = torch.randn(3, generator=gen, dtype=torch.float64).numpy()
    d /= np.linalg.norm(d)
    surf = d * radii.double().numpy() * 0.92
    hand_hun = (sc * hv.astype(np.float64)) @ R.T + surf + 0.25 * tr
    # T_h2m: similarity, scale 0.3, modest rotation, object about 1.5 units in front of the camera
    qh = torch.tensor([1.0, 0.0, 0.0, 0.0]) + 0.25 * torch.randn(4, generator=gen)
    Rh = quat_to_mat_np(qh.numpy())
    s_h2m = 0.3 * (0.9 + 0.2 * torch.rand(1, generator=gen).item())
    T = np.eye(4)
    T[:3, :3] = s_h2m * Rh
    T[:3, 3] = np.array([0.05, -0.03, -1.5]) + 0.05 * torch.randn(3, generator=gen, dtype=torch.float64).numpy()
    hand_moge = hand_hun @ T[:3, :3].T + T[:3, 3]
    # cloud: half hand-surface samples, half camera-facing ellipsoid samples (+noise), in MoGe space
    n_hand = P // 2
    fi = torch.randint(0, hf.shape[0], (n_hand,), generator=gen).numpy()
    bc = torch.rand(n_hand, 2, generator=gen, dtype=torch.float64).numpy()
    flip = bc.sum(1) > 1
    bc[flip] = 1 - bc[flip]
    tri = hand_moge[hf[fi]]
    hp = tri[:, 0] + bc[:, :1] * (tri[:, 1] - tri[:, 0]) + bc[:, 1:] * (tri[:, 2] - tri[:, 0])
    n_obj = P - n_hand
    dirs = torch.randn(n_obj, 3, generator=gen, dtype=torch.float64).numpy()
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    op_h = dirs * radii.double().numpy()
    op = op_h @ T[:3, :3].T + T[:3, 3]
    # camera sits at the MoGe origin: mirror samples whose outward normal points away from it
    nrm = (dirs / radii.double().numpy()) @ Rh.T
    away = (nrm * (-op)).sum(1) <= 0
    op[away] = ((-dirs[away]) * radii.double().numpy()) @ T[:3, :3].T + T[:3, 3]
    cloud = np.concatenate([hp, op], 0)
    cloud = cloud + 0.005 * s_h2m / 0.3 * torch.randn(P, 3, generator=gen, dtype=torch.float64).numpy()
    perm = torch.randperm(P, generator=gen).numpy()
    cloud = cloud[perm]

    obj_center = T[:3, 3].copy()  # centre of the Hunyuan cube mapped to MoGe space

    def leaves():
        s = 1.0 + perturb * 0.05 * torch.randn(1, generator=gen)
        t = perturb * 0.01 * torch.randn(3, generator=gen)
        qq = torch.tensor([1.0, 0.0, 0.0, 0.0]) + perturb * 0.05 * torch.randn(4, generator=gen)
        return torch.cat([s, t, qq]).float()

    theta_h = leaves()
    theta_o = leaves()
    J = torch.from_numpy(standin_j_regressor(0))
    kps = 256.0 + 60.0 * torch.randn(21, 2, generator=gen)
    return GuidanceSample(
        sdf=sdf.to(device),
        hand_rest=torch.from_numpy(hand_moge.astype(np.float32)).to(device),
        hand_faces=torch.from_numpy(hf).to(device),
        cloud=torch.from_numpy(cloud.astype(np.float32)).to(device),
        T_h2m=torch.from_numpy(T.astype(np.float32)).to(device),
        obj_center=torch.from_numpy(obj_center.astype(np.float32)).to(device),
        theta_h=theta_h.to(device), theta_o=theta_o.to(device),
        j_regressor=J.to(device), kps_2d=kps.float().to(device),
        fov_deg=41.0, image_hw=(512, 512),
    )


def icosphere(subdiv: int = 3, radius: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """Closed triangle sphere (used as a watertight known-answer mesh)."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
         (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4),
         (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8),
         (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}
        nf = []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return (np.asarray(v) * radius).astype(np.float32), np.asarray(f, dtype=np.int32)


def random_similarity(seed: int, scale_range=(0.8, 1.5), trans=0.3) -> np.ndarray:
    """4x4 float64 similarity with a generic rotation (ICP known-answer tests)."""
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    R = quat_to_mat_np(q / np.linalg.norm(q))
    s = rng.uniform(*scale_range)
    T = np.eye(4)
    T[:3, :3] = s * R
    T[:3, 3] = rng.uniform(-trans, trans, size=3)
    return T


def stack_samples(samples, device="cuda:0", with_kp: bool = True, cap: bool = False):
    """Batch ``GuidanceSample``s into the engine's inputs: (sdf [B,D,D,D], theta [B,16], GuidanceStatics).
    ``cap=True`` closes the wrist loop of the (shared) hand topology first."""
    from .guidance.engine import GuidanceStatics
    dev = torch.device(device)
    sdf = torch.stack([s.sdf for s in samples]).to(dev).contiguous()
    theta = torch.stack([torch.cat([s.theta_h, s.theta_o]) for s in samples]).to(dev).contiguous()
    st = GuidanceStatics(
        hand_rest=torch.stack([s.hand_rest for s in samples]).to(dev).contiguous(),
        hand_faces=(torch.from_numpy(cap_boundary_loops(samples[0].hand_faces.cpu().numpy())) if cap
                    else samples[0].hand_faces).to(torch.int32).to(dev).contiguous(),
        cloud=torch.stack([s.cloud for s in samples]).to(dev).contiguous(),
        T_h2m=torch.stack([s.T_h2m for s in samples]).to(dev).contiguous(),
        obj_center=torch.stack([s.obj_center for s in samples]).to(dev).contiguous(),
        j_regressor=samples[0].j_regressor.to(dev).contiguous() if with_kp else None,
        kps_2d=torch.stack([s.kps_2d for s in samples]).to(dev).contiguous() if with_kp else None,
        fov_deg=samples[0].fov_deg, image_hw=samples[0].image_hw)
    return sdf, theta, st
