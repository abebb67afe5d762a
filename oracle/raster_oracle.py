"""CPU ORACLE for the renderer of the guidance loop -- TEST INFRASTRUCTURE, NOT PRODUCT
CODE.  Only ``tests/`` may import it.

What the reference runs every inner iteration (third_party_patches/hy3dgen/shapegen/pipelines.py:272-275):

    norms = renderer(mesh)                              # MeshRasterizer(naive, faces_per_pixel=1) + PhongNormalShader
    depth = renderer.rasterizer(mesh).zbuf              # the same rasterisation again

with the camera and settings of src/foho/guidance/run.py:84-105 (``FoVPerspectiveCameras(R = diag(-1, 1, -1), T = 0,
znear 0.01, zfar 100, fov = MoGe fov_x``, ``blur_radius = log(1/1e-4 - 1) * 1e-8``, ``bin_size = -1``) and the shader
of pipelines.py:74-92 (pixel colour = SUM of the top face's three vertex normals -- the barycentrics are replaced by
ones -- blended by ``softmax_rgb_blend`` with sigma = gamma = 1e-8 over a white background).

PARITY UNPINNED: pytorch3d is not vendored (git HEAD, unpinned: scripts/create_env_foho.sh:74-80) and cannot be
imported offline.  The functions below restate FROM MEMORY what its kernels compute
(``pytorch3d/csrc/rasterize_meshes/rasterize_meshes.cu``: ``CheckPixelInsideFace``; ``utils/geometry_utils.cuh``:
``EdgeFunctionForward``, ``BarycentricCoordsForward``, ``BarycentricPerspectiveCorrectionForward``;
``renderer/cameras.py``: ``FoVPerspectiveCameras.compute_projection_matrix``; ``structures/meshes.py``:
``_compute_vertex_normals``; ``renderer/blending.py``: ``softmax_rgb_blend``) in plain torch, so that autograd
defines the gradients the rasteriser's hand-written backward returns for ``zbuf`` and the vertex normals.
Deliberate simplification, stated once: the blend probability ``sigmoid(-dist / 1e-8)`` is taken as exactly 1 on
covered pixels (it cancels out of the blended colour; it differs from 1 only for pixel centres within 1e-4 NDC
units of a face edge) -- so alpha is 0 / 1 and no gradient flows through the edge distances.

Conventions: NDC +x points LEFT and +y UP; pixel (row i, column j) has its centre at
``(x, y) = (1 - (2 j + 1) / W, 1 - (2 i + 1) / H)``; z is view-space depth; a face is a candidate for a pixel when
all three (perspective-corrected) barycentrics are > 0, the interpolated depth is >= 0 and |2 area| > 1e-8; the
nearest candidate wins, ties go to the smaller face index.
"""
from __future__ import annotations

import math

import torch

K_EPS = 1e-8
ZNEAR, ZFAR = 0.01, 100.0
BLEND_EPS = 1e-10


def camera_rotation(dtype=torch.float64) -> torch.Tensor:
    """``rotation_y_180`` of src/foho/guidance/run.py:84-86."""
    return torch.tensor([[-1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, -1.0]], dtype=dtype)


def project(verts: torch.Tensor, fov_deg: float, R: torch.Tensor = None):
    """World -> (NDC xy, view depth).  pytorch3d: view = X R + T (row vectors, T = 0 here), then the FoV projection
    x_ndc = x_v / (z_v tan(fov / 2)), y_ndc = y_v / (z_v tan(fov / 2)) (aspect ratio 1)."""
    R = camera_rotation(verts.dtype) if R is None else R.to(verts.dtype)
    v = verts @ R
    t = math.tan(math.radians(fov_deg) / 2.0)
    z = v[:, 2]
    return torch.stack([v[:, 0] / (z * t), v[:, 1] / (z * t)], -1), z


def vertex_normals(verts: torch.Tensor, faces: torch.Tensor) -> torch.Tensor:
    """``Meshes.verts_normals_packed``: every vertex accumulates the (area-weighted) normals of its faces, then
    ``F.normalize(eps=1e-6)``."""
    f = faces.long()
    v0, v1, v2 = verts[f[:, 0]], verts[f[:, 1]], verts[f[:, 2]]
    n = torch.zeros_like(verts)
    n = n.index_add(0, f[:, 1], torch.cross(v2 - v1, v0 - v1, dim=1))
    n = n.index_add(0, f[:, 2], torch.cross(v0 - v2, v1 - v2, dim=1))
    n = n.index_add(0, f[:, 0], torch.cross(v1 - v0, v2 - v0, dim=1))
    return n / n.norm(dim=1, keepdim=True).clamp_min(1e-6)


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def pixel_centres(H: int, W: int, dtype=torch.float64):
    ys = 1.0 - (2.0 * torch.arange(H, dtype=dtype) + 1.0) / H
    xs = 1.0 - (2.0 * torch.arange(W, dtype=dtype) + 1.0) / W
    return xs, ys


def _bary(px, py, xy, z, f, perspective_correct=True):
    """Barycentrics (pytorch3d formulas) of points (px, py) w.r.t. faces f [..., 3]; broadcasting over leading dims."""
    x0, y0 = xy[f[..., 0], 0], xy[f[..., 0], 1]
    x1, y1 = xy[f[..., 1], 0], xy[f[..., 1], 1]
    x2, y2 = xy[f[..., 2], 0], xy[f[..., 2], 1]
    z0, z1, z2 = z[f[..., 0]], z[f[..., 1]], z[f[..., 2]]
    area = _edge(x2, y2, x0, y0, x1, y1)
    a = area + K_EPS
    w0 = _edge(px, py, x1, y1, x2, y2) / a
    w1 = _edge(px, py, x2, y2, x0, y0) / a
    w2 = _edge(px, py, x0, y0, x1, y1) / a
    if perspective_correct:
        t0, t1, t2 = w0 * z1 * z2, z0 * w1 * z2, z0 * z1 * w2
        den = (t0 + t1 + t2).clamp_min(K_EPS)
        w0, w1, w2 = t0 / den, t1 / den, t2 / den
    pz = w0 * z0 + w1 * z1 + w2 * z2
    return w0, w1, w2, pz, area, torch.maximum(torch.maximum(z0, z1), z2)


def rasterize(xy: torch.Tensor, z: torch.Tensor, faces: torch.Tensor, H: int, W: int, chunk: int = 4096):
    """Naive rasterisation, one face per pixel.  Returns (pix_to_face [H, W] int64 (-1 = none), zbuf [H, W]
    (-1 = none, differentiable w.r.t. xy and z), bary [H, W, 3])."""
    f = faces.long()
    xs, ys = pixel_centres(H, W, xy.dtype)
    px = xs.view(1, W).expand(H, W).reshape(-1)
    py = ys.view(H, 1).expand(H, W).reshape(-1)
    best_z = torch.full((H * W,), float("inf"), dtype=xy.dtype)
    best_f = torch.full((H * W,), -1, dtype=torch.long)
    with torch.no_grad():
        for s in range(0, H * W, chunk):
            w0, w1, w2, pz, area, zmax = _bary(px[s:s + chunk, None], py[s:s + chunk, None], xy, z, f[None])
            ok = (w0 > 0) & (w1 > 0) & (w2 > 0) & (pz >= 0) & (area.abs() > K_EPS) & (zmax >= 0)
            pzm = torch.where(ok, pz, torch.full_like(pz, float("inf")))
            m = pzm.min(dim=1).values
            first = (pzm == m[:, None]).to(torch.int8).argmax(dim=1)          # smallest face index among the nearest
            hit = torch.isfinite(m)
            best_z[s:s + chunk] = m
            best_f[s:s + chunk] = torch.where(hit, first, torch.full_like(first, -1))
    hit = best_f >= 0
    fi = best_f.clamp_min(0)
    w0, w1, w2, pz, _, _ = _bary(px, py, xy, z, f[fi])
    zbuf = torch.where(hit, pz, torch.full_like(pz, -1.0))
    bary = torch.stack([w0, w1, w2], -1) * hit[:, None]
    return best_f.view(H, W), zbuf.view(H, W), bary.view(H, W, 3)


def render_normals_and_depth(verts: torch.Tensor, faces: torch.Tensor, fov_deg: float, H: int, W: int):
    """``renderer(mesh)`` and ``renderer.rasterizer(mesh).zbuf`` (pipelines.py:273-274).  Returns
    (norms4 [H, W, 4] = blended normal colour + alpha, zbuf [H, W, 1], pix_to_face [H, W])."""
    xy, z = project(verts, fov_deg)
    p2f, zbuf, _ = rasterize(xy, z, faces, H, W)
    vn = vertex_normals(verts, faces)
    f = faces.long()
    hit = p2f >= 0
    fi = p2f.clamp_min(0)
    col = vn[f[fi, 0]] + vn[f[fi, 1]] + vn[f[fi, 2]]                       # barycentrics replaced by ones (:86-89)
    # softmax_rgb_blend with K = 1, prob = 1 on covered pixels, background (1, 1, 1)
    zi = ((ZFAR - zbuf) / (ZFAR - ZNEAR)) * hit
    zi_max = zi.clamp_min(BLEND_EPS)
    wnum = hit.to(verts.dtype) * torch.exp((zi - zi_max) / 1e-8)
    delta = torch.exp((BLEND_EPS - zi_max) / 1e-8).clamp_min(BLEND_EPS)
    den = wnum + delta
    rgb = (wnum[..., None] * col + delta[..., None] * 1.0) / den[..., None]
    alpha = hit.to(verts.dtype)
    return torch.cat([rgb, alpha[..., None]], -1), zbuf[..., None], p2f
