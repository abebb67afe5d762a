"""CPU ORACLE for the guidance hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product
(``followmyhold_b200``) never does; it fails loudly when its CUDA library is missing.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
 and none of its hot-path modules can be imported offline
(pytorch3d / kaolin / trimesh / hy3dgen are absent).  This file therefore *restates*
the reference arithmetic term by term (rows marked REF, each citing file:line under
/root/reference) and *defines* in plain torch autograd the volume-sampling terms that
BASELINE.json's north_star adds (rows marked NS).  It is pinned only by the analytic
known-answer tests in ``tests/test_oracle_*.py``.  Library semantics asserted from
memory (pytorch3d ``knn_points`` / ``quaternion_to_matrix`` / ``mesh_edge_loss`` /
``FoVPerspectiveCameras``) are flagged "APPENDIX-C" below.

Conventions
-----------
* Volume ``sdf`` [D,D,D]: index [ix,iy,iz], z fastest, sample positions
  linspace(-1.10, 1.10, D) per axis in *Hunyuan space*, negative inside
  (pipelines.py:341-360, 309-312).
* ``theta`` = [s, tx,ty,tz, qw,qx,qy,qz] (8 floats) for hand and object leaves
  (third_party/utilz/code_utils.py:57-78; identity init pipelines.py:1208-1215).
* MoGe space is where all losses live (pipelines.py:1241, 1520).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Dict, Optional

import numpy as np
import torch

GRID_BOUND = 1.10
FINGERTIPS = (744, 320, 443, 554, 671)
MANO_TO_OPENPOSE = (0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20)


@dataclass
class Weights:
    """Loss weights.  REF values are the literals of pipelines.py:1499-1504,1578-1588."""
    # REF
    w_dist: float = 10.0       # 10 * distance_loss                      (:1580)
    w_vreg: float = 1e-3       # 1e-3 * obj_verts_loss_3                 (:1584)
    w_edge: float = 1.0        # 1 * obj_loss_3                          (:1585)
    w_treg_o: float = 1e-3     # 1e-3 * loss_obj_reg                     (:1586)
    w_hand: float = 1e-3       # 1e-3 * hand_loss                        (:1587)
    w_kp: float = 1e-4         # inside hand_loss: 1e-4 * loss_2d_kps    (:1500)
    w_treg_h: float = 1e-2     # inside hand_loss: 1e-2 * loss_hand_trans(:1503)
    w_int_lo: float = 1e-9     # w_intersection default                  (:1564)
    w_int_hi: float = 1e-5     # w_intersection when close & late        (:1562)
    dist_margin: float = 0.01  # margin on the *squared* distance        (:1539)
    # NS (north_star volume terms; weights mirror the REF "10 *" data terms)
    w_pen: float = 10.0
    w_con: float = 10.0
    w_ivol: float = 10.0
    w_ch: float = 10.0
    w_mom: float = 1e-3
    con_margin: float = 0.01


# --------------------------------------------------------------------------- a6
def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """REF/APPENDIX-C: pytorch3d.transforms.quaternion_to_matrix (real-first, scaled by
    2/(q.q), *not* normalised by the caller: pipelines.py:1484,1524)."""
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack([
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def bbox_center(verts: torch.Tensor) -> torch.Tensor:
    """REF pipelines.py:111: (min+max)/2 of the *current* verts (autograd routes the
    gradient to the arg-min / arg-max vertices)."""
    return (verts.min(dim=0)[0] + verts.max(dim=0)[0]) / 2.0


def transform_around_center_w_scale(verts, theta, center=None):
    """REF pipelines.py:108-118 with RT built as in :1482-1486."""
    s, t, q = theta[0], theta[1:4], theta[4:8]
    R = quaternion_to_matrix(q)
    c = bbox_center(verts) if center is None else center
    return (s * (verts - c)) @ R.T + c + t


def transform_hunyuan2moge(verts, T):
    """REF pipelines.py:242-250."""
    return verts @ T[:3, :3].T + T[:3, 3]


# --------------------------------------------------------------------------- a2
def set_timesteps_sigmas(num_inference_steps: int, shift: float = 1.0) -> torch.Tensor:
    """REF schedulers.py:171-211 called with sigmas=linspace(0,1,N) (pipelines.py:1187):
    sigma <- shift*sigma/(1+(shift-1)*sigma); sigmas = cat(sigma,[1])."""
    s = np.linspace(0, 1, num_inference_steps)
    s = shift * s / (1 + (shift - 1) * s)
    s = torch.from_numpy(s).to(torch.float32)
    return torch.cat([s, torch.ones(1)])


def scheduler_step(sample, model_output, sigma, sigma_next):
    """REF schedulers.py:294-309: fp32 up-cast, prev = x + (s'-s) v, x1 = x + (1-s) v,
    both cast back to model_output.dtype."""
    x = sample.to(torch.float32)
    prev = x + (sigma_next - sigma) * model_output
    x1 = x + (1 - sigma) * model_output
    return prev.to(model_output.dtype), x1.to(model_output.dtype)


def scheduler_step_final(sample, model_output, sigma):
    """REF schedulers.py:470-484."""
    x = sample.to(torch.float32)
    return (x + (1 - sigma) * model_output).to(model_output.dtype)


# --------------------------------------------------------------------------- a13
def world_to_grid(x_hun: torch.Tensor, D: int) -> torch.Tensor:
    """Continuous grid index of a Hunyuan-space point for linspace(-b,b,D) sampling."""
    return (x_hun + GRID_BOUND) * ((D - 1) / (2.0 * GRID_BOUND))


def trilinear_sample(vol: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """NS a13: trilinear lookup at continuous index g [V,3] (order ix,iy,iz), border
    clamp (== F.grid_sample(bilinear, align_corners=True, padding_mode='border') with
    the axis flip; checked in tests/test_oracle_guidance.py)."""
    D = vol.shape[0]
    gc = torch.minimum(torch.maximum(g, torch.zeros_like(g)), torch.full_like(g, D - 1))
    i0 = torch.clamp(torch.floor(gc.detach()), max=D - 2).long()
    f = gc - i0.to(gc.dtype)
    x0, y0, z0 = i0.unbind(-1)
    fx, fy, fz = f.unbind(-1)
    out = 0
    for dx in (0, 1):
        wx = fx if dx else 1 - fx
        for dy in (0, 1):
            wy = fy if dy else 1 - fy
            for dz in (0, 1):
                wz = fz if dz else 1 - fz
                out = out + wx * wy * wz * vol[x0 + dx, y0 + dy, z0 + dz]
    return out


# --------------------------------------------------------------------------- a8 (sign rule)
def _edge_sign(lo, hi, X, Y):
    """Canonical (index-ordered) edge function sign with simulation-of-simplicity ties.

    All arithmetic is IEEE float32 with separately rounded products, exactly what the
    CUDA kernel does with __fmul_rn/__fsub_rn (no FMA contraction)."""
    dx = hi[0] - lo[0]
    dy = hi[1] - lo[1]
    e = dx * (Y - lo[1]) - dy * (X - lo[0])          # float32 ops, products rounded first
    s = np.sign(e).astype(np.int32)
    tie = s == 0
    if np.any(tie):
        t = np.where(dy != 0, -np.sign(dy), np.sign(dx)).astype(np.int32)
        s = np.where(tie, t, s)
    return s, e


def raster_parity_inside(verts_grid: np.ndarray, faces: np.ndarray, D: int) -> np.ndarray:
    """Inside/outside of every voxel centre of a D^3 lattice by ray parity along +z.

    DEFINED RULE (SURVEY.md §7 "inside/outside for an open mesh"; stands in for
    kaolin.ops.mesh.check_sign, third_party/utilz/kaolin_sdf_ops.py:104, whose source is
    not in the tree -- APPENDIX-C):  voxel (X,Y,Z) is inside iff the number of
    triangles whose xy-projection contains (X,Y) -- consistent half-open edge rule,
    shared edges counted once -- and whose plane crosses the column at zc with
    float32(Z) < zc is odd.  verts_grid float32 [V,3] in lattice units.
    Returns bool [D,D,D].
    """
    v = np.ascontiguousarray(verts_grid, dtype=np.float32)
    par = np.zeros((D, D, D), dtype=bool)
    zs = np.arange(D, dtype=np.float32)
    for (ia, ib, ic) in faces:
        a, b, c = v[ia], v[ib], v[ic]
        xmin = max(int(np.ceil(min(a[0], b[0], c[0]))), 0)
        xmax = min(int(np.floor(max(a[0], b[0], c[0]))), D - 1)
        ymin = max(int(np.ceil(min(a[1], b[1], c[1]))), 0)
        ymax = min(int(np.floor(max(a[1], b[1], c[1]))), D - 1)
        if xmin > xmax or ymin > ymax:
            continue
        X, Y = np.meshgrid(np.arange(xmin, xmax + 1, dtype=np.float32),
                           np.arange(ymin, ymax + 1, dtype=np.float32), indexing="ij")

        def oriented(iu, iv, pu, pv):
            if iu < iv:
                s, e = _edge_sign(pu, pv, X, Y)
                return s, e
            s, e = _edge_sign(pv, pu, X, Y)
            return -s, -e

        s_ab, e_ab = oriented(ia, ib, a, b)
        s_bc, e_bc = oriented(ib, ic, b, c)
        s_ca, e_ca = oriented(ic, ia, c, a)
        hit = (s_ab == s_bc) & (s_bc == s_ca)
        if not hit.any():
            continue
        wa, wb, wc = e_bc, e_ca, e_ab                     # barycentric weights (unnormalised)
        den = (wa + wb) + wc
        hit &= den != 0
        with np.errstate(divide="ignore", invalid="ignore"):
            zc = ((wa * a[2] + wb * b[2]) + wc * c[2]) / den
        xi, yi = np.nonzero(hit)
        for k in range(xi.size):
            par[xmin + xi[k], ymin + yi[k], :] ^= zs < zc[xi[k], yi[k]]
    return par


# --------------------------------------------------------------------------- a8 (distance)
def closest_point_barycentric(p: torch.Tensor, tri: torch.Tensor):
    """Closest point of triangles tri [F,3,3] to points p [M,3] (Ericson, Real-Time
    Collision Detection §5.1.5).  Returns (d2 [M], face [M], bary [M,3]) of the nearest
    face.  Restates the semantics of kaolin point_to_mesh_distance (squared distance to
    the nearest face/edge/vertex; kaolin_sdf_ops.py:100) -- APPENDIX-C."""
    a, b, c = tri[:, 0][None], tri[:, 1][None], tri[:, 2][None]
    P = p[:, None, :]
    ab, ac, ap = b - a, c - a, P - a
    d1 = (ab * ap).sum(-1); d2 = (ac * ap).sum(-1)
    bp = P - b
    d3 = (ab * bp).sum(-1); d4 = (ac * bp).sum(-1)
    cp = P - c
    d5 = (ab * cp).sum(-1); d6 = (ac * cp).sum(-1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    M, F = d1.shape
    wa = torch.zeros(M, F, dtype=p.dtype); wb = torch.zeros_like(wa); wc = torch.zeros_like(wa)
    done = torch.zeros(M, F, dtype=torch.bool)

    def put(mask, A, B, C):
        nonlocal wa, wb, wc, done
        m = mask & ~done
        wa = torch.where(m, A, wa); wb = torch.where(m, B, wb); wc = torch.where(m, C, wc)
        done = done | m

    one = torch.ones_like(wa); zero = torch.zeros_like(wa)
    put((d1 <= 0) & (d2 <= 0), one, zero, zero)                      # vertex a
    put((d3 >= 0) & (d4 <= d3), zero, one, zero)                     # vertex b
    v = d1 / torch.where((d1 - d3) != 0, d1 - d3, one)
    put((vc <= 0) & (d1 >= 0) & (d3 <= 0), 1 - v, v, zero)           # edge ab
    put((d6 >= 0) & (d5 <= d6), zero, zero, one)                     # vertex c
    w = d2 / torch.where((d2 - d6) != 0, d2 - d6, one)
    put((vb <= 0) & (d2 >= 0) & (d6 <= 0), 1 - w, zero, w)           # edge ac
    den = (d4 - d3) + (d5 - d6)
    w2 = (d4 - d3) / torch.where(den != 0, den, one)
    put((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), zero, 1 - w2, w2)  # edge bc
    s = va + vb + vc
    s = torch.where(s != 0, s, one)
    put(torch.ones_like(done), 1 - vb / s - vc / s, vb / s, vc / s)  # face interior
    q = wa[..., None] * a + wb[..., None] * b + wc[..., None] * c
    dd = ((P - q) ** 2).sum(-1)
    d2min, fidx = dd.min(dim=1)
    ar = torch.arange(M)
    bary = torch.stack([wa[ar, fidx], wb[ar, fidx], wc[ar, fidx]], -1)
    return d2min, fidx, bary


def point_mesh_distance(p: torch.Tensor, verts: torch.Tensor, faces: torch.Tensor, chunk: int = 512):
    """Differentiable unsigned distance of p [M,3] to mesh (verts [V,3], faces [F,3]).

    The nearest face and barycentric weights are found without grad; the distance is
    then |p - sum_k w_k v_k| with w constant -- exact gradient by the envelope theorem
    (the closest point is a constrained minimiser)."""
    faces = faces.long()
    fidx_all, bary_all = [], []
    with torch.no_grad():
        tri = verts.detach()[faces]
        for s in range(0, p.shape[0], chunk):
            _, fi, ba = closest_point_barycentric(p.detach()[s:s + chunk], tri)
            fidx_all.append(fi); bary_all.append(ba)
    if not fidx_all:
        return p.new_zeros(0), torch.zeros(0, dtype=torch.long), p.new_zeros(0, 3)
    fidx = torch.cat(fidx_all); bary = torch.cat(bary_all)
    tv = verts[faces[fidx]]                              # [M,3,3]
    q = (bary[..., None] * tv).sum(1)
    d = torch.sqrt(((p - q) ** 2).sum(-1))
    return d, fidx, bary


def mesh2sdf(verts: torch.Tensor, faces: torch.Tensor, grid_points: torch.Tensor,
             inside: torch.Tensor) -> torch.Tensor:
    """REF kaolin_sdf_ops.py:88-109: sdf = sqrt(point->mesh d^2) * (inside ? -1 : +1)."""
    d, _, _ = point_mesh_distance(grid_points, verts, faces)
    return d * torch.where(inside, -torch.ones_like(d), torch.ones_like(d))


def honerf_intersection_loss(sdf_hand, sdf_obj):
    """REF pipelines.py:231-239 (integer count / 1000; no gradient)."""
    obj_inner = sdf_obj < 0
    return (sdf_hand[obj_inner] < 0).sum() / 1000


# --------------------------------------------------------------------------- a7 / a15
def knn1_sq(p1: torch.Tensor, p2: torch.Tensor, chunk: int = 2048):
    """REF/APPENDIX-C pytorch3d.ops.knn_points(K=1): squared L2 to the nearest p2 for each
    p1, differentiable w.r.t. both clouds (pipelines.py:1529-1532)."""
    idx = []
    with torch.no_grad():
        for s in range(0, p1.shape[0], chunk):
            d = torch.cdist(p1[s:s + chunk].double(), p2.double())
            idx.append(d.argmin(dim=1))
    idx = torch.cat(idx) if idx else torch.zeros(0, dtype=torch.long)
    diff = p1 - p2[idx]
    return (diff * diff).sum(-1), idx


# --------------------------------------------------------------------------- a10
def mesh_edge_loss(verts: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """REF/APPENDIX-C pytorch3d.loss.mesh_edge_loss(target_length=0) for one mesh:
    mean over unique edges of |v0 - v1|^2 (pipelines.py:1575)."""
    v0, v1 = verts[edges[:, 0].long()], verts[edges[:, 1].long()]
    return ((v0 - v1).norm(dim=1) ** 2).mean()


def unique_edges(faces: np.ndarray) -> np.ndarray:
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
    return np.unique(np.sort(e, axis=1), axis=0).astype(np.int32)


# --------------------------------------------------------------------------- a11
def mano_vert_to_3dkps(verts, J_regressor):
    """REF pipelines.py:121-135."""
    tips = verts[list(FINGERTIPS)]
    kps = torch.cat([J_regressor @ verts, tips], 0)
    return kps[list(MANO_TO_OPENPOSE)]


def fov_project_screen(points, fov_deg: float, H: int, W: int):
    """REF/APPENDIX-C FoVPerspectiveCameras(R=diag(-1,1,-1), T=0, fov=deg)
    .transform_points_screen (guidance/run.py:84-90, pipelines.py:1491): view = X R + T
    (row vectors); ndc = (x,y)/(z tan(fov/2)); screen = size/2 - min(H,W)/2 * ndc."""
    xv, yv, zv = -points[:, 0], points[:, 1], -points[:, 2]
    th = np.tan(np.deg2rad(fov_deg) / 2.0)
    xn = xv / (zv * th)
    yn = yv / (zv * th)
    sc = min(H, W) / 2.0
    return torch.stack([W / 2.0 - sc * xn, H / 2.0 - sc * yn], -1)


# --------------------------------------------------------------------------- the energy
def object_frame(theta_o, T_h2m, obj_center, D):
    """Affine lattice -> MoGe map  y = A g + b  of the object under (a5) T_h2m then (a6)
    the similarity theta_o about ``obj_center``.  Returns (A [3,3], b [3], kappa) with
    kappa = MoGe length of one lattice step (A^T A = kappa^2 I)."""
    s, t, q = theta_o[0], theta_o[1:4], theta_o[4:8]
    R = quaternion_to_matrix(q)
    A_h = T_h2m[:3, :3]
    t_h = T_h2m[:3, 3]
    step = (2.0 * GRID_BOUND) / (D - 1)
    # x_hun = g*step - bound ; x_m = A_h x_hun + t_h ; y = s R (x_m - c) + c + t
    A = s * (R @ A_h) * step
    b = s * (R @ (A_h @ (-GRID_BOUND * torch.ones(3, dtype=theta_o.dtype)) + t_h - obj_center)) + obj_center + t
    s_h2m = torch.linalg.norm(A_h[:, 0])
    kappa = s * s_h2m * step
    return A, b, kappa


def guidance_energy(sdf, hand_rest, hand_faces, cloud, theta_h, theta_o, T_h2m, obj_center,
                    weights: Weights = Weights(), *, j_regressor=None, kps_2d=None, fov_deg=41.0,
                    image_hw=(512, 512), obj_verts=None, obj_faces=None, obj_edges=None,
                    late_step: bool = False, hand_grid_verts_override: Optional[np.ndarray] = None,
                    ) -> Dict[str, torch.Tensor]:
    """One guidance evaluation (forward).  Call ``.backward()`` on ``out['total']``.

    Differentiable inputs: ``sdf`` (dense dE/dSDF), ``theta_h``, ``theta_o`` and, when an
    explicit object mesh is supplied (the FlexiCubes output of pipelines.py:1509, in
    Hunyuan space), ``obj_verts``.

    REF terms (need obj_verts): a7 distance_loss, a10 obj_verts_loss / mesh_edge_loss.
    REF terms (always): a6 transforms, a10 translation regularisers, a11 key-points,
    a9 count (computed here on the *Hunyuan lattice*, see NS a14), a12 assembly.
    NS terms: a13 pen/con, a14 L_int, a15 chamfer, a10v L_mom.

    ``hand_grid_verts_override`` lets a test feed the kernel's own float32 lattice-space
    hand vertices to the (non-differentiable) sign rule, making the integer count
    bit-comparable.
    """
    W = weights
    D = sdf.shape[0]
    N = float(D) ** 3
    dt = sdf.dtype
    out: Dict[str, torch.Tensor] = {}

    # a6 hand
    hm = transform_around_center_w_scale(hand_rest, theta_h)
    # object frame (a5 + a6 with a fixed centre)
    A, b, kappa = object_frame(theta_o, T_h2m, obj_center, D)
    Ainv = torch.linalg.inv(A)
    hg = (hm - b) @ Ainv.T                                    # lattice coordinates of hand verts

    # a13
    s_v = trilinear_sample(sdf, hg)
    out["L_pen"] = torch.relu(-s_v).mean()
    out["L_con"] = torch.clamp(s_v.abs() - W.con_margin, min=0).mean()

    # a14 + a9: sign by the parity rule on the lattice, distance in MoGe space
    hg32 = hg.detach().to(torch.float32).numpy() if hand_grid_verts_override is None else hand_grid_verts_override
    inside_h = torch.from_numpy(raster_parity_inside(hg32, hand_faces.cpu().numpy(), D))
    both = inside_h & (sdf.detach() < 0)
    out["count"] = both.sum().to(dt) / 1000.0                # a9 (no gradient)
    idx = both.nonzero()
    if idx.shape[0] > 0:
        y = idx.to(dt) @ A.T + b                             # voxel centres in MoGe space
        d_h, _, _ = point_mesh_distance(y, hm, hand_faces)   # = -SDF_h there
        s_o = sdf[idx[:, 0], idx[:, 1], idx[:, 2]]
        out["L_int"] = (torch.relu(-s_o) * d_h).sum() / N
    else:
        out["L_int"] = sdf.sum() * 0.0

    # a10v: occupancy-weighted second moment in MoGe space
    w_occ = torch.relu(-sdf)
    ar = torch.arange(D, dtype=dt)
    gx, gy, gz = ar.view(D, 1, 1), ar.view(1, D, 1), ar.view(1, 1, D)
    yx = A[0, 0] * gx + A[0, 1] * gy + A[0, 2] * gz + b[0]
    yy = A[1, 0] * gx + A[1, 1] * gy + A[1, 2] * gz + b[1]
    yz = A[2, 0] * gx + A[2, 1] * gy + A[2, 2] * gz + b[2]
    out["L_mom"] = (w_occ * (yx * yx + yy * yy + yz * yz)).sum() / N

    # a15 chamfer (squared, K=1, both directions)
    d_hc, _ = knn1_sq(hm, cloud)
    d_ch, _ = knn1_sq(cloud, hm)
    out["L_ch"] = d_hc.mean() + d_ch.mean()

    # a10 translation regularisers
    out["L_treg_h"] = (theta_h[1:4] ** 2).mean()
    out["L_treg_o"] = (theta_o[1:4] ** 2).mean()

    # a11
    if j_regressor is not None and kps_2d is not None:
        k3 = mano_vert_to_3dkps(hm, j_regressor)
        k2 = fov_project_screen(k3, fov_deg, image_hw[0], image_hw[1])
        out["L_kp"] = ((k2 - kps_2d) ** 2).mean()
    else:
        out["L_kp"] = hm.sum() * 0.0

    # REF mesh terms
    if obj_verts is not None:
        om = transform_hunyuan2moge(obj_verts, T_h2m)        # a5
        ot = transform_around_center_w_scale(om, theta_o)    # a6 (centre from current verts)
        d_ho, _ = knn1_sq(hm, ot)                            # a7
        out["mean_d2"] = d_ho.mean().detach()
        out["L_dist"] = torch.clamp(d_ho - W.dist_margin, min=0).mean()
        out["L_vreg"] = (ot ** 2).mean()
        if obj_edges is None:
            obj_edges = torch.from_numpy(unique_edges(obj_faces.cpu().numpy()))
        out["L_edge"] = mesh_edge_loss(ot, obj_edges)
        close = bool(out["mean_d2"] < 0.001)
    else:
        z = hm.sum() * 0.0
        out["L_dist"] = z; out["L_vreg"] = z; out["L_edge"] = z
        close = False
    w_int = W.w_int_hi if (close and late_step) else W.w_int_lo   # pipelines.py:1561-1564

    hand_loss = W.w_kp * out["L_kp"] + W.w_treg_h * out["L_treg_h"]
    out["total"] = (w_int * out["count"] + W.w_dist * out["L_dist"] + W.w_vreg * out["L_vreg"]
                    + W.w_edge * out["L_edge"] + W.w_treg_o * out["L_treg_o"] + W.w_hand * hand_loss
                    + W.w_pen * out["L_pen"] + W.w_con * out["L_con"] + W.w_ivol * out["L_int"]
                    + W.w_ch * out["L_ch"] + W.w_mom * out["L_mom"])
    out["hand_grid"] = hg.detach()
    out["hand_moge"] = hm.detach()
    return out


def guidance_energy_and_grads(sample, weights: Weights = Weights(), dtype=torch.float32, **kw):
    """Convenience for tests/bench: run forward + backward on a ``GuidanceSample``-like
    object and return (terms, dE/dsdf, dE/dtheta_h, dE/dtheta_o)."""
    sdf = sample.sdf.detach().to(dtype).clone().requires_grad_(True)
    th = sample.theta_h.detach().to(dtype).clone().requires_grad_(True)
    to = sample.theta_o.detach().to(dtype).clone().requires_grad_(True)
    out = guidance_energy(sdf, sample.hand_rest.to(dtype), sample.hand_faces, sample.cloud.to(dtype), th, to,
                          sample.T_h2m.to(dtype), sample.obj_center.to(dtype), weights,
                          j_regressor=sample.j_regressor.to(dtype), kps_2d=sample.kps_2d.to(dtype),
                          fov_deg=sample.fov_deg, image_hw=sample.image_hw, **kw)
    out["total"].backward()
    return out, sdf.grad, th.grad, to.grad


# --------------------------------------------------------------------------- a4
def adamw_step(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-4, weight_decay=0.01):
    """REF torch.optim.AdamW(eps=1e-4) single-tensor update (pipelines.py:1384,1478;
    weight_decay=0 reproduces torch.optim.Adam of :1318).  ``step`` is 1-based.
    Checked against torch.optim.AdamW in tests/test_oracle_update.py."""
    p = p * (1 - lr * weight_decay)
    m = m + (g - m) * (1 - beta1)
    v = v * beta2 + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / (bc2 ** 0.5) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


def _fma32(a, b, c):
    """One correctly rounded float32 fused multiply-add (numpy has none): the product of two float32 is
    exact in the 64-bit mantissa of x87 long double, so only the final rounding to float32 matters."""
    ld = np.longdouble
    return (np.asarray(a, np.float32).astype(ld) * np.asarray(b, np.float32).astype(ld)
            + np.asarray(c, np.float32).astype(ld)).astype(np.float32)


def adamw_step_torch_ops(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-4, weight_decay=0.01,
                         order: str = "cuda"):
    """REF a4, rounding for rounding: one torch.optim.Adam/AdamW step on numpy float32 or float16 arrays as
    the sequence of elementwise ops torch launches (torch/optim/adam.py; call sites pipelines.py:1318,
    1384,1478), each op computed in float32 and rounded to the tensors' dtype -- for a half leaf (the
    reference's velocity, code_utils.py:43-78) the moments are half and every op rounds to half:

        p.mul_(1-lr*wd); m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2)
        den = (v.sqrt() / bc2**0.5).add_(eps); p.addcdiv_(m, den, -lr/bc1)

    ``order`` picks how torch's kernels group the three `a + s*x` ops:
      "cuda": lerp = fma(w, g-m, m); addcmul = fma(s, g*g, v); addcdiv = fma(s, m/den, p)  (ATen CUDA
              functors `a + alpha * (b * c)` / `a + alpha * (b / c)`, contracted by nvcc; the path the
              reference runs, and the one the kernels follow),
      "cpu":  lerp = fma(w, g-m, m); addcmul = fma(s*g, g, v); addcdiv = p + (s*m)/den  (ATen CPU kernels
              `self + alpha * t1 * t2`, `self + alpha * t1 / t2`) -- pinned bit for bit by torch.optim on
              CPU in tests/test_oracle_update.py.
    Scalars are python doubles cast to float32, as torch passes them.  Returns (p, m, v) in the input dtype."""
    dt = np.asarray(p).dtype
    assert dt in (np.float32, np.float16)
    f32 = np.float32
    R = (lambda x: x.astype(f32)) if dt == np.float32 else (lambda x: x.astype(np.float16).astype(f32))
    p, g, m, v = (np.asarray(a).astype(f32) for a in (p, g, m, v))
    w1, w2, b2 = f32(1 - beta1), f32(1 - beta2), f32(beta2)
    bc1 = 1 - beta1 ** step
    bc2_sqrt = f32((1 - beta2 ** step) ** 0.5)
    neg_step = f32((lr / bc1) * -1)
    p = R(p * f32(1 - lr * weight_decay))
    m = R(_fma32(w1, g - m, m))
    v = R(v * b2)
    v = R(_fma32(w2, g * g, v)) if order == "cuda" else R(_fma32(w2 * g, g, v))
    den = R(np.sqrt(v))
    den = R(den / bc2_sqrt)
    den = R(den + f32(eps))
    p = R(_fma32(neg_step, m / den, p)) if order == "cuda" else R(p + (neg_step * m) / den)
    return p.astype(dt), m.astype(dt), v.astype(dt)


def step_final_torch_ops(x_t, v, sigma: float):
    """REF a2 ``step_final`` on numpy float32 / float16 arrays with torch's rounding (schedulers.py:470-484):
    float32: separately rounded product and sum; float16: the fp32 0-dim factor is cast to half, the product
    rounded to half, the sum with the upcast sample taken in fp32 and cast back to half."""
    dt = np.asarray(v).dtype
    f32 = np.float32
    oms = f32(1.0) - f32(sigma)
    if dt == np.float32:
        return (x_t.astype(f32) + oms * v.astype(f32)).astype(f32)
    prod = (oms.astype(np.float16).astype(f32) * v.astype(f32)).astype(np.float16).astype(f32)
    return (x_t.astype(f32) + prod).astype(np.float16)
