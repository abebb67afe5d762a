"""CPU ORACLE for the surface extraction of the guidance loop (second part) -- TEST
INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/`` may import it.

What the reference runs (third_party_patches/hy3dgen/shapegen/pipelines.py:1142-1143,1393,1509,1642):

    fc = kaolin.non_commercial.FlexiCubes(device); x_nx3, cube_fx8 = fc.construct_voxel_grid(octree_res)
    verts, faces, _ = fc(x_nx3 * 2.2 ..., sdf.view(-1), cube_fx8, octree_res)        # no weights passed

PARITY UNPINNED, and deliberately NOT a restatement: kaolin is not vendored (0.17.0, scripts/create_env_foho.sh:68),
FlexiCubes rests on lookup tables (``dmc_table``, ``num_vd_table``, ``tet_table``) that cannot be reproduced from
memory, and its ambiguous-configuration handling (up to four dual vertices per cube) and quad-splitting rule come out
of those tables.  What this file DEFINES instead is the scheme FlexiCubes reduces to when no weights are passed and no
cube is ambiguous -- Dual Marching Cubes:

  * one dual vertex per lattice cube whose eight corners are not all of one sign (inside = SDF < 0), placed at the
    MEAN of the zero crossings of the cube's sign-changing edges, each crossing by linear interpolation
    ``p = a + (b - a) * s_a / (s_a - s_b)``;
  * one quad per interior lattice edge with a sign change, joining the dual vertices of the four cubes around it,
    wound so that the normal points from inside to outside, split along the diagonal (0, 2);
  * orders: vertices by cube index (x-major, z fastest, like the lattice), faces by (edge axis, lattice point index),
    the two triangles of a quad adjacent.

Vertex positions are differentiable functions of the SDF (torch ops), so autograd defines the gradient the CUDA
backward is held against.
"""
from __future__ import annotations

import torch

_CORNERS = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]      # index = 4 dx + 2 dy + dz
_EDGES = [(a, b) for a in range(8) for b in range(a + 1, 8) if bin(a ^ b).count("1") == 1]                # 12 cube edges


def extract(sdf: torch.Tensor, bound: float = 1.10):
    """``sdf`` [D, D, D] (negative inside).  Returns (verts [Nv, 3] in lattice world units, faces [Nf, 3] int64,
    edges [Ne, 2] int64 unique)."""
    D = sdf.shape[0]
    n = D - 1
    step = 2.0 * bound / (D - 1)
    s8 = torch.stack([sdf[dx:dx + n, dy:dy + n, dz:dz + n] for dx, dy, dz in _CORNERS], -1).reshape(-1, 8)
    inside = s8 < 0
    active = inside.any(1) & ~inside.all(1)
    cube_ids = active.nonzero().reshape(-1)
    sa = s8[cube_ids]
    ia = inside[cube_ids]
    num = torch.zeros(cube_ids.shape[0], 3, dtype=sdf.dtype)
    cnt = torch.zeros(cube_ids.shape[0], dtype=sdf.dtype)
    corner = torch.tensor(_CORNERS, dtype=sdf.dtype)
    for a, b in _EDGES:
        cross = ia[:, a] != ia[:, b]
        den = sa[:, a] - sa[:, b]
        t = torch.where(cross, sa[:, a] / torch.where(cross, den, torch.ones_like(den)), torch.zeros_like(den))
        p = corner[a][None] + (corner[b] - corner[a])[None] * t[:, None]
        num = num + p * cross[:, None]
        cnt = cnt + cross
    local = num / cnt[:, None]
    ci = torch.stack([cube_ids // (n * n), (cube_ids // n) % n, cube_ids % n], -1).to(sdf.dtype)
    verts = -bound + step * (ci + local)
    vidx = torch.full((n * n * n,), -1, dtype=torch.long)
    vidx[cube_ids] = torch.arange(cube_ids.shape[0])
    vid3 = vidx.view(n, n, n)
    neg = sdf < 0
    faces, diag = [], []
    for d in range(3):
        u, w = (d + 1) % 3, (d + 2) % 3
        # lattice edges from point q to q + e_d with both other coordinates interior (four cubes around the edge exist)
        rng = [torch.arange(D)] * 3
        rng[d] = torch.arange(D - 1)
        rng[u] = torch.arange(1, D - 1)
        rng[w] = torch.arange(1, D - 1)
        g = torch.stack(torch.meshgrid(*rng, indexing="ij"), -1).reshape(-1, 3)
        q2 = g.clone(); q2[:, d] += 1
        s0 = neg[g[:, 0], g[:, 1], g[:, 2]]
        s1 = neg[q2[:, 0], q2[:, 1], q2[:, 2]]
        sel = s0 != s1
        g, s0 = g[sel], s0[sel]
        quad = []
        for du, dw in ((-1, -1), (0, -1), (0, 0), (-1, 0)):                      # counter-clockwise around +e_d
            c = g.clone(); c[:, u] += du; c[:, w] += dw
            quad.append(vid3[c[:, 0], c[:, 1], c[:, 2]])
        quad = torch.stack(quad, -1)
        quad = torch.where(s0[:, None], quad, quad[:, [0, 3, 2, 1]])              # inside at the low end: normal along +e_d
        # order by lattice point index of q (x-major), which meshgrid(ij) over ascending ranges already gives
        faces.append(torch.stack([quad[:, [0, 1, 2]], quad[:, [0, 2, 3]]], 1).reshape(-1, 3))
        diag.append(quad[:, [0, 2]])
    faces = torch.cat(faces) if faces else torch.zeros(0, 3, dtype=torch.long)
    # unique edges: one per pair of face-adjacent active cubes that share a sign-changing lattice edge, plus the diagonals
    e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    e = torch.unique(torch.sort(e, 1).values, dim=0)
    return verts, faces, e
