"""CPU checks of the renderer oracle (oracle/raster_oracle.py; pytorch3d semantics restated from memory, PARITY
UNPINNED): analytic known answers for projection, coverage and depth, and a float64 finite-difference check of the
gradient autograd returns -- the gradient the CUDA rasteriser's backward is then held against."""
import math

import numpy as np
import torch

from oracle import raster_oracle as RO


def test_projection_and_pixel_convention():
    # camera looks down -z of the MoGe frame (R = diag(-1, 1, -1)): a point at (0, 0, -2) projects to the image centre
    xy, z = RO.project(torch.tensor([[0.0, 0.0, -2.0], [0.5, 0.25, -2.0]], dtype=torch.float64), 60.0)
    t = math.tan(math.radians(30.0))
    assert torch.allclose(z, torch.tensor([2.0, 2.0], dtype=torch.float64))
    assert torch.allclose(xy[0], torch.zeros(2, dtype=torch.float64))
    # view x = -world x; NDC +x points left, so world +x lands in the RIGHT half of the image (column > W/2)
    assert torch.allclose(xy[1], torch.tensor([-0.5 / (2 * t), 0.25 / (2 * t)], dtype=torch.float64))
    xs, ys = RO.pixel_centres(4, 4)
    assert torch.allclose(xs, torch.tensor([0.75, 0.25, -0.25, -0.75], dtype=torch.float64)) and torch.allclose(xs, ys)


def test_fronto_parallel_triangle_depth_and_coverage():
    H = W = 64
    z0 = 1.7
    t = math.tan(math.radians(41.0 / 2))
    s = 0.6 * z0 * t                                        # NDC half-size 0.6
    verts = torch.tensor([[-s, -s, -z0], [s, -s, -z0], [0.0, s, -z0]], dtype=torch.float64)
    faces = torch.tensor([[0, 1, 2]])
    n4, zb, p2f = RO.render_normals_and_depth(verts, faces, 41.0, H, W)
    hit = p2f >= 0
    assert abs(int(hit.sum()) - 0.5 * 1.2 * 1.2 / 4 * H * W) <= 0.03 * H * W          # area of the NDC triangle / 4 of the image
    assert torch.allclose(zb[..., 0][hit], torch.full((int(hit.sum()),), z0, dtype=torch.float64), atol=1e-6)
    assert (zb[..., 0][~hit] == -1).all()
    # one face: every vertex normal is the face normal (0, 0, +-1); the shader SUMS the three vertex normals
    assert torch.allclose(n4[..., :3][hit].abs(), torch.tensor([0.0, 0.0, 3.0], dtype=torch.float64).expand(int(hit.sum()), 3), atol=1e-5)
    assert torch.allclose(n4[..., :3][~hit], torch.ones(3, dtype=torch.float64).expand(int((~hit).sum()), 3))      # white background
    assert set(n4[..., 3].unique().tolist()) == {0.0, 1.0}


def test_nearest_face_wins_and_ties_go_to_the_smaller_index():
    verts = torch.tensor([[-1, -1, -2.0], [1, -1, -2.0], [0, 1, -2.0], [-1, -1, -1.5], [1, -1, -1.5], [0, 1, -1.5],
                          [-1, -1, -1.5], [1, -1, -1.5], [0, 1, -1.5]], dtype=torch.float64) * torch.tensor([0.3, 0.3, 1.0], dtype=torch.float64)
    faces = torch.tensor([[0, 1, 2], [3, 4, 5], [6, 7, 8]])
    _, zb, p2f = RO.render_normals_and_depth(verts, faces, 41.0, 32, 32)
    centre = p2f[16, 16]
    assert int(centre) == 1 and abs(float(zb[16, 16, 0]) - 1.5) < 1e-9


def test_autograd_gradient_matches_finite_differences():
    torch.manual_seed(0)
    from followmyhold_b200.synthetic import standin_hand_mesh
    hv, hf = standin_hand_mesh(0.35)
    v0 = torch.as_tensor(hv, dtype=torch.float64)[:200].clone()
    f = torch.as_tensor(hf)
    f = f[(f < 200).all(1)]
    v0[:, 2] -= 1.2
    H = W = 24

    def loss(v):
        n4, zb, _ = RO.render_normals_and_depth(v, f, 41.0, H, W)
        hit = n4[..., 3] > 0
        return (zb[..., 0] * hit).sum() * 0.1 + (n4[..., :3] * hit[..., None] * torch.tensor([0.3, -0.2, 0.5], dtype=torch.float64)).sum()

    v = v0.clone().requires_grad_(True)
    L = loss(v)
    L.backward()
    g = v.grad
    # central differences on the vertices with the largest gradient (coverage is piecewise constant: eps small enough
    # that no pixel changes face)
    idx = g.abs().sum(1).argsort(descending=True)[:4]
    for i in idx.tolist():
        for a in range(3):
            e = torch.zeros_like(v0); e[i, a] = 1e-7
            fd = (loss(v0 + e) - loss(v0 - e)) / 2e-7
            assert abs(float(fd) - float(g[i, a])) <= 1e-4 * max(1.0, abs(float(g[i, a]))), (i, a, float(fd), float(g[i, a]))
