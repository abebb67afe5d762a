"""Row f2 (first part) on the GPU: the rasteriser + fused image losses + backward to the vertices
(``foho_raster_losses_fwd_bwd``) against the float64 torch oracle (oracle/raster_oracle.py for the renderer, the
reference's own normalisation / loss formulas of pipelines.py:276-287,178-187,1567-1569 written with torch ops, so
autograd defines every gradient).  pytorch3d's semantics are restated from memory: PARITY UNPINNED for the renderer;
the loss arithmetic is pinned separately by tests/golden/ref_golden_image_losses.npz (test_oracle_image_losses.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _image_losses_torch(n4, zbuf, gt_n, gt_mask, gt_disp, gt_sil):
    """pipelines.py:276-287 + the three losses, torch ops only (autograd: full-reduction min / max share their gradient
    among ties)."""
    n = n4[..., :3]
    mask = n4[..., 3] > 0
    rn = (n - n.min()) / (n.max() - n.min() + 1e-6)
    rn = rn * mask[..., None]
    z = torch.where(zbuf[..., 0] < 0, torch.full_like(zbuf[..., 0], 10.0), zbuf[..., 0])
    d = 1.0 / (z + 1e-6)
    rd = (d - d.min()) / (d.max() - d.min() + 1e-6)
    cos = (F.normalize(rn, dim=-1) * F.normalize(gt_n, dim=-1)).sum(-1)
    l_n = (1 - cos)[gt_mask].mean()
    l_d = F.l1_loss(rd, gt_disp)
    l_s = F.binary_cross_entropy(mask.to(rd.dtype), gt_sil)
    return l_n, l_d, l_s, rn, rd


def _scene(seed, H, W, fov):
    """A posed stand-in hand and a bumpy sphere in front of the camera (MoGe frame: the camera looks down -z)."""
    from followmyhold_b200.synthetic import standin_hand_mesh
    from oracle import raster_oracle as RO
    g = torch.Generator().manual_seed(seed)
    hv, hf = standin_hand_mesh(0.35)
    hv = torch.as_tensor(hv, dtype=torch.float64); hf = torch.as_tensor(hf, dtype=torch.int64)
    # icosphere-like object: lat-long sphere
    nu, nv = 24, 12
    u = torch.arange(nu, dtype=torch.float64) / nu * 2 * np.pi
    v = (torch.arange(1, nv, dtype=torch.float64)) / nv * np.pi
    r = 0.18 * (1 + 0.1 * torch.randn(nv - 1, nu, generator=g).double())
    sv = torch.stack([(r * torch.sin(v)[:, None] * torch.cos(u)[None]).reshape(-1), (r * torch.cos(v)[:, None].expand(-1, nu)).reshape(-1),
                      (r * torch.sin(v)[:, None] * torch.sin(u)[None]).reshape(-1)], -1)
    fs = []
    for i in range(nv - 2):
        for j in range(nu):
            a, b = i * nu + j, i * nu + (j + 1) % nu
            c, d = a + nu, b + nu
            fs += [[a, c, b], [b, c, d]]
    sf = torch.tensor(fs, dtype=torch.int64)
    verts = torch.cat([hv + torch.tensor([-0.05, 0.02, -1.1], dtype=torch.float64), sv + torch.tensor([0.12, -0.03, -1.25], dtype=torch.float64)])
    verts = verts + 0.002 * torch.randn(verts.shape, generator=g).double()
    faces = torch.cat([hf, sf + hv.shape[0]])
    # targets: the oracle's render of a perturbed copy (what the MoGe render is to the reference, pipelines.py:1247-1256)
    tv = verts + torch.tensor([0.03, -0.02, 0.05], dtype=torch.float64) + 0.004 * torch.randn(verts.shape, generator=g).double()
    with torch.no_grad():
        n4, zb, _ = RO.render_normals_and_depth(tv, faces, fov, H, W)
        _, _, _, rn, rd = _image_losses_torch(n4, zb, torch.ones(H, W, 3, dtype=torch.float64), n4[..., 3] > 0,
                                              torch.zeros(H, W, dtype=torch.float64), torch.zeros(H, W, dtype=torch.float64))
    return verts, faces, rn, n4[..., 3] > 0, rd, (n4[..., 3] > 0).double()


@pytest.mark.parametrize("B,H,W", [(1, 64, 64), (2, 96, 80)])
def test_rasteriser_losses_and_vertex_gradients_match_the_oracle(B, H, W):
    from followmyhold_b200.guidance.render import ImageLossRenderer, ImageTargets, pack_meshes
    from oracle import raster_oracle as RO
    fovs = [41.0, 55.0][:B]
    scenes = [_scene(10 + b, H, W, fovs[b]) for b in range(B)]
    verts, faces, vo, fo = pack_meshes([(s[0], s[1]) for s in scenes])
    R = ImageLossRenderer(B, H, W, verts.shape[0], faces.shape[0], tile_cap=2304)   # tiny image: a whole mesh per tile
    R.set_targets(ImageTargets(gt_normals=torch.stack([s[2] for s in scenes]), gt_mask=torch.stack([s[3] for s in scenes]),
                               gt_disp=torch.stack([s[4] for s in scenes]), gt_sil=torch.stack([s[5] for s in scenes]),
                               fov_deg=torch.tensor(fovs)))
    losses, gv, dbg = R(verts, faces, vo, fo, debug=True)
    torch.cuda.synchronize()
    losses = losses.cpu().double(); gv = gv.cpu().double()
    assert (losses[:, 7] == 0).all()                        # no tile overflow
    R_small = ImageLossRenderer(B, H, W, verts.shape[0], faces.shape[0], tile_cap=64)
    R_small.targets, R_small._n_valid = R.targets, R._n_valid
    assert (R_small(verts, faces, vo, fo, backward=False)[0][:, 7] == 1).all()          # ... and an overflow is reported, not silent
    off = 0
    for b, (v, f, gt_n, gt_m, gt_d, gt_s) in enumerate(scenes):
        vq = v.float().double().clone().requires_grad_(True)            # the kernel sees float32 vertices
        n4, zb, p2f = RO.render_normals_and_depth(vq, f, fovs[b], H, W)
        l_n, l_d, l_s, _, _ = _image_losses_torch(n4, zb, gt_n, gt_m, gt_d, gt_s)
        total = 10 * l_n + 10 * l_d + 10 * l_s
        (10 * l_n + 10 * l_d).backward()                    # the hard silhouette carries no gradient
        # coverage: identical up to pixel centres within float32 rounding of an edge
        k_p2f = dbg["p2f"][b].cpu().long()
        k_p2f = torch.where(k_p2f >= 0, k_p2f - int(fo[b]), k_p2f)
        assert (k_p2f != p2f).float().mean().item() <= 2e-3
        same = k_p2f == p2f
        assert torch.allclose(dbg["zbuf"][b].cpu().double()[same], zb[..., 0][same].detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(dbg["nraw"][b].cpu().double()[same], n4[..., :3][same].detach(), atol=2e-4)
        for got, ref in ((losses[b, 0], l_n), (losses[b, 1], l_d), (losses[b, 2], l_s), (losses[b, 3], total)):
            assert abs(float(got) - float(ref)) <= 3e-3 * abs(float(ref)) + 1e-6, (b, float(got), float(ref))
        g_ref = vq.grad
        g_got = gv[off:off + v.shape[0]]
        cos = F.cosine_similarity(g_got.reshape(1, -1), g_ref.reshape(1, -1)).item()
        assert cos > 0.999, cos
        assert (g_got - g_ref).abs().max().item() <= 2e-2 * g_ref.abs().max().item()
        off += v.shape[0]


def test_forward_only_and_determinism():
    from followmyhold_b200.guidance.render import ImageLossRenderer, ImageTargets, pack_meshes
    H = W = 64
    s = _scene(3, H, W, 41.0)
    verts, faces, vo, fo = pack_meshes([(s[0], s[1])])
    R = ImageLossRenderer(1, H, W, verts.shape[0], faces.shape[0])
    R.set_targets(ImageTargets(gt_normals=s[2][None], gt_mask=s[3][None], gt_disp=s[4][None], gt_sil=s[5][None], fov_deg=torch.tensor([41.0])))
    l0, g0 = R(verts, faces, vo, fo)
    l0, g0 = l0.clone(), g0.clone()
    for _ in range(3):                                       # fixed-point accumulation: bit-identical run to run
        l1, g1 = R(verts, faces, vo, fo)
        assert torch.equal(l0, l1) and torch.equal(g0, g1)
    l2, g2 = R(verts, faces, vo, fo, backward=False)
    assert g2 is None and torch.equal(l2, l0)


def test_rendered_hand_terms_reach_the_leaves_like_autograd():
    """Phase-1 hand terms of the reference (1 * normal + 10 * disparity + 1 * silhouette, pipelines.py:1327-1349) wired
    into the fused evaluation through ``grad_hand_ext``: the change they make to dE/d(s_h, t_h, q_h) equals autograd
    through the oracle's similarity transform (a6), renderer and losses."""
    from followmyhold_b200.guidance.loop import GuidanceLoop
    from followmyhold_b200.guidance.render import ImageTargets
    from followmyhold_b200.synthetic import make_guidance_sample, stack_samples
    from oracle import guidance_oracle as O
    from oracle import raster_oracle as RO
    B, D, P, H, W = 2, 32, 512, 128, 128
    fovs = [20.0, 24.0]                      # narrow: the 0.1-unit hand at 1.5 units covers a few hundred pixels
    samples = [make_guidance_sample(D, P, 30 + i) for i in range(B)]
    sdf, theta, st = stack_samples(samples, cap=True)
    raw_faces = samples[0].hand_faces.to(torch.int64)
    assert float(st.hand_rest[..., 2].max()) < 0          # MoGe frame: the scene sits in front of a camera looking down -z
    loop = GuidanceLoop(B, D, st, P, micro_batches=1, mock_decoder=False)
    loop.theta.copy_(theta); loop.sdf.copy_(sdf)
    tg = []
    for b in range(B):
        th = theta[b, :8].cpu().double()
        hm = O.transform_around_center_w_scale(st.hand_rest[b].cpu().double(), th)
        tv = hm + torch.tensor([0.01, -0.015, 0.02], dtype=torch.float64)
        with torch.no_grad():
            n4, zb, _ = RO.render_normals_and_depth(tv, raw_faces, fovs[b], H, W)
            m = n4[..., 3] > 0
            _, _, _, rn, rd = _image_losses_torch(n4, zb, torch.ones(H, W, 3, dtype=torch.float64), m, torch.zeros(H, W, dtype=torch.float64),
                                                  torch.zeros(H, W, dtype=torch.float64))
        tg.append((rn, m, rd * m, m.double()))
    assert all(int(t[1].sum()) > 50 for t in tg), "the synthetic hand must be visible"
    loop.enable_image_terms(ImageTargets(gt_normals=torch.stack([t[0] for t in tg]), gt_mask=torch.stack([t[1] for t in tg]),
                                         gt_disp=torch.stack([t[2] for t in tg]), gt_sil=torch.stack([t[3] for t in tg]),
                                         fov_deg=torch.tensor(fovs)), hand_faces_render=raw_faces.to(torch.int32), tile_cap=2048)
    ln = loop.lanes[0]
    eng = ln.engine
    s = torch.cuda.current_stream()
    w1 = loop.phase_weights(1)
    res = []
    for with_img in (False, True):
        g = loop._hand_image_grad(ln, 1, s, w1) if with_img else None
        desc = eng.make_desc(loop.sdf, loop.theta, st, grad_hand_ext=g)
        desc.w = w1
        eng.launch(desc, s)
        torch.cuda.synchronize()
        res.append(eng.grad_theta.clone().cpu().double())
    diff = (res[1] - res[0])[:, :8]
    assert (res[1] - res[0])[:, 8:].abs().max() == 0            # the object leaves do not see the hand render
    for b in range(B):
        th = theta[b, :8].cpu().double().clone().requires_grad_(True)
        hm = O.transform_around_center_w_scale(st.hand_rest[b].cpu().double(), th)
        n4, zb, _ = RO.render_normals_and_depth(hm, raw_faces, fovs[b], H, W)
        l_n, l_d, l_s, _, _ = _image_losses_torch(n4, zb, *tg[b])
        (float(w1.w_hand) * (1.0 * l_n + 10.0 * l_d)).backward()
        ref = th.grad
        assert abs(float(loop.image_terms[b, 0]) - float(l_n)) <= 5e-3 * abs(float(l_n)) + 1e-6
        assert abs(float(loop.image_terms[b, 1]) - float(l_d)) <= 5e-3 * abs(float(l_d)) + 1e-6
        assert (diff[b] - ref).abs().max().item() <= 3e-2 * ref.abs().max().item(), (diff[b], ref)
