"""Row f2 (second part) on the GPU: Dual Marching Cubes extraction and its backward against oracle/surface_oracle.py
(which DEFINES the scheme: parity unpinned against kaolin's FlexiCubes, see its header)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _volumes(B, D, seed=0):
    g = torch.Generator().manual_seed(seed)
    ax = torch.linspace(-1.1, 1.1, D)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    vols = []
    for b in range(B):
        r = 0.35 + 0.3 * torch.rand(3, generator=g)
        c = 0.2 * (torch.rand(3, generator=g) - 0.5)
        s = torch.sqrt(((X - c[0]) / r[0]) ** 2 + ((Y - c[1]) / r[1]) ** 2 + ((Z - c[2]) / r[2]) ** 2) - 1.0
        vols.append(s + 0.03 * torch.sin(7 * X + b) * torch.cos(5 * Y) + 0.02 * torch.randn(D, D, D, generator=g))
    return torch.stack(vols).float().contiguous()


@pytest.mark.parametrize("B,D", [(1, 17), (3, 33), (2, 65)])
def test_extraction_equals_the_oracle(B, D):
    from followmyhold_b200.guidance.surface import SurfaceExtractor
    from oracle import surface_oracle as SO
    sdf = _volumes(B, D, seed=D)
    ex = SurfaceExtractor(B, D, index_base=7)
    ex.extract(sdf.cuda())
    torch.cuda.synchronize()
    ex.check_flags()
    for b, (v, f, e) in enumerate(ex.meshes()):
        ov, of, oe = SO.extract(sdf[b].double())
        assert v.shape[0] == ov.shape[0] and f.shape[0] == of.shape[0]
        assert torch.allclose(v.double(), ov, atol=2e-5)                       # same vertices in the same (cube) order
        assert torch.equal(f, of)                                              # same triangles in the same order
        es = torch.unique(torch.sort(e, 1).values, dim=0)
        assert es.shape[0] == e.shape[0] and torch.equal(es, oe)               # unique edges: the oracle's set, no duplicates
        # closed manifold on noisy ellipsoids well inside the lattice
        ee = torch.sort(torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), 1).values
        _, cnt = torch.unique(ee, dim=0, return_counts=True)
        assert (cnt == 2).float().mean() > 0.99
    ex2 = SurfaceExtractor(B, D, index_base=7)
    ex2.extract(sdf.cuda())
    assert torch.equal(ex.verts, ex2.verts) and torch.equal(ex.faces, ex2.faces) and torch.equal(ex.edges, ex2.edges)     # deterministic


def test_backward_equals_autograd_through_the_oracle():
    from followmyhold_b200.guidance.surface import SurfaceExtractor
    from oracle import surface_oracle as SO
    B, D = 2, 25
    sdf = _volumes(B, D, seed=3)
    ex = SurfaceExtractor(B, D)
    ex.extract(sdf.cuda())
    vo = ex.vert_offsets.tolist()
    g = torch.Generator().manual_seed(1)
    gv = torch.zeros(ex.cap_verts, 3)
    gv[:vo[B]] = torch.randn(vo[B], 3, generator=g)
    gs = torch.zeros(B, D, D, D, device="cuda")
    gs[0, 0, 0, 0] = 5.0                                                       # the backward ACCUMULATES into grad_sdf
    ex.backward(gv.cuda(), gs)
    torch.cuda.synchronize()
    for b in range(B):
        s = sdf[b].double().requires_grad_(True)
        ov, _, _ = SO.extract(s)
        (ov * gv[vo[b]:vo[b + 1]].double()).sum().backward()
        ref = s.grad.clone()
        if b == 0:
            ref[0, 0, 0] += 5.0
        assert (gs[b].cpu().double() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item()
    gs2 = torch.zeros(B, D, D, D, device="cuda"); gs2[0, 0, 0, 0] = 5.0
    ex.backward(gv.cuda(), gs2)
    assert torch.equal(gs, gs2)                                                # fixed-point scatter: bit-identical


def test_capacity_overflow_is_flagged():
    from followmyhold_b200 import _lib
    from followmyhold_b200.guidance.surface import SurfaceExtractor
    sdf = _volumes(1, 33, seed=9).cuda()
    ex = SurfaceExtractor(1, 33, cap_verts=100)
    ex.extract(sdf)
    with pytest.raises(_lib.FohoStatusError):
        ex.check_flags()
