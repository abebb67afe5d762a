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
    # consumers trust the offsets: they never point beyond the capacities, and no face names a dropped vertex
    assert ex.vert_offsets.tolist() == [0, 100] and ex.face_offsets[-1] <= ex.cap_faces and ex.edge_offsets[-1] <= ex.cap_edges
    assert int(ex.faces.max()) < 100 and int(ex.edges.max()) < 100


# --------------------------------------------------------------------------- the extracted mesh inside the guidance evaluation
def _hoi_setup(B=1, D=25, P=512, H=96, W=96, fov=20.0, seed=40):
    import torch.nn.functional as F
    from followmyhold_b200.guidance.loop import GuidanceLoop
    from followmyhold_b200.guidance.render import ImageTargets
    from followmyhold_b200.synthetic import make_guidance_sample, stack_samples
    from oracle import guidance_oracle as O
    from oracle import raster_oracle as RO
    from oracle import surface_oracle as SO
    from tests.test_gpu_raster import _image_losses_torch
    samples = [make_guidance_sample(D, P, seed + i) for i in range(B)]
    sdf, theta, st = stack_samples(samples, cap=True)
    raw_faces = samples[0].hand_faces.to(torch.int64)
    cap = B * 12 * D * D                  # room for the noisy surfaces a random-weight decoder produces
    loop = GuidanceLoop(B, D, st, P, micro_batches=1, mock_decoder=False, max_obj_verts=cap)
    loop.theta.copy_(theta); loop.sdf.copy_(sdf)
    fovs = [fov + 2 * b for b in range(B)]
    tg_hand, tg_hoi = [], []
    for b, s in enumerate(samples):
        with torch.no_grad():
            hm = O.transform_around_center_w_scale(s.hand_rest.double(), s.theta_h.double()) + torch.tensor([0.01, -0.01, 0.02], dtype=torch.float64)
            ov, of, _ = SO.extract(s.sdf.double() + 0.03)
            ot = O.transform_around_center_w_scale(O.transform_hunyuan2moge(ov, s.T_h2m.double()), s.theta_o.double())
            for verts, faces, dst in ((hm, raw_faces, tg_hand), (torch.cat([hm, ot]), torch.cat([raw_faces, of + hm.shape[0]]), tg_hoi)):
                n4, zb, _ = RO.render_normals_and_depth(verts, faces, fovs[b], H, W)
                m = n4[..., 3] > 0
                _, _, _, rn, rd = _image_losses_torch(n4, zb, torch.ones(H, W, 3, dtype=torch.float64), m, torch.zeros(H, W, dtype=torch.float64),
                                                      torch.zeros(H, W, dtype=torch.float64))
                dst.append((rn, m, rd, m.double()))
    mk = lambda tg: ImageTargets(gt_normals=torch.stack([t[0] for t in tg]), gt_mask=torch.stack([t[1] for t in tg]),
                                 gt_disp=torch.stack([t[2] for t in tg]), gt_sil=torch.stack([t[3] for t in tg]), fov_deg=torch.tensor(fovs))
    loop.enable_image_terms(mk(tg_hand), hand_faces_render=raw_faces.to(torch.int32), tile_cap=4096)
    loop.enable_object_terms(hoi_targets=mk(tg_hoi), obj_targets=mk(tg_hoi), tile_cap=4096, cap_verts=cap)
    return loop, samples, st, raw_faces, fovs, tg_hoi, (H, W)


def test_joined_image_terms_reach_leaves_and_volume_like_autograd():
    """The joined hand + object terms of the joint phase (10 nrm + 10 disp + 10 sil, pipelines.py:1544-1569,1580-1583) through
    the whole new chain -- extraction, T_h2m + object similarity about the bbox centre (a5, a6), joint render, losses, and
    back: to theta_h, theta_o and dE/dSDF -- against autograd through the oracles of every link."""
    import ctypes as C
    from followmyhold_b200 import _lib
    from oracle import guidance_oracle as O
    from oracle import raster_oracle as RO
    from oracle import surface_oracle as SO
    from tests.test_gpu_raster import _image_losses_torch
    loop, samples, st, raw_faces, fovs, tg_hoi, (H, W) = _hoi_setup()
    ln = loop.lanes[0]
    w = _lib.Weights()                                           # every kernel-side weight zero: the image terms alone
    s = torch.cuda.current_stream()
    loop._object_eval(ln, 2, False, w, s)
    torch.cuda.synchronize()
    loop._obj.ex.check_flags()
    gt = ln.engine.grad_theta.cpu().double()
    gs = ln.engine.grad_sdf.cpu().double()
    for b, smp in enumerate(samples):
        sdf = smp.sdf.double().clone().requires_grad_(True)
        th = smp.theta_h.double().clone().requires_grad_(True)
        to = smp.theta_o.double().clone().requires_grad_(True)
        hm = O.transform_around_center_w_scale(smp.hand_rest.double(), th)
        ov, of, _ = SO.extract(sdf)
        ot = O.transform_around_center_w_scale(O.transform_hunyuan2moge(ov, smp.T_h2m.double()), to)
        n4, zb, _ = RO.render_normals_and_depth(torch.cat([hm, ot]), torch.cat([raw_faces, of + hm.shape[0]]), fovs[b], H, W)
        l_n, l_d, l_s, _, _ = _image_losses_torch(n4, zb, *tg_hoi[b])
        (10 * l_n + 10 * l_d).backward()
        for got, ref in ((loop._obj.terms_hoi[b, 0], l_n), (loop._obj.terms_hoi[b, 1], l_d), (loop._obj.terms_hoi[b, 2], l_s)):
            assert abs(float(got) - float(ref)) <= 5e-3 * abs(float(ref)) + 1e-6, (float(got), float(ref))
        ref_t = torch.cat([th.grad, to.grad])
        assert (gt[b] - ref_t).abs().max().item() <= 3e-2 * ref_t.abs().max().item(), (gt[b], ref_t)
        ref_s = sdf.grad
        cos = torch.nn.functional.cosine_similarity(gs[b].reshape(1, -1), ref_s.reshape(1, -1)).item()
        assert cos > 0.995, cos
        assert (gs[b] - ref_s).abs().max().item() <= 5e-2 * ref_s.abs().max().item()


def test_schedule_with_extracted_mesh_and_all_image_terms():
    """The whole schedule with the tensor-core decoder, the extracted object mesh, the REF mesh terms and every image term."""
    from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder, lattice_points
    from followmyhold_b200.guidance.config import OptimizationConfig
    from tests.test_gpu_decoder import _vae
    loop, samples, st, raw_faces, fovs, tg_hoi, _ = _hoi_setup(B=2, D=17, seed=60)
    cfg = loop.cfg
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 2, 2, 2
    cfg.with_steps(6)
    loop.sigmas = __import__("followmyhold_b200.guidance.loop", fromlist=["x"]).set_timesteps_sigmas(6)
    loop.nan_steps = torch.zeros(6, 2, dtype=torch.int32, device="cuda")
    vae = _vae(1, seed=21)
    with torch.no_grad():
        vae.geo_decoder.output_proj.weight.mul_(3.0)
    dec = LatentDecoder(DecoderWeights(vae.state_dict(), "cuda:0"), 2, query_chunk=2048, active_chunk=512)
    dec.set_queries(lattice_points(17))
    g = torch.Generator().manual_seed(2)
    loop.x_t.copy_(torch.randn(2, loop.L, generator=g))
    vel = (0.5 * torch.randn(2, loop.L, generator=g)).cuda()
    theta0 = loop.theta.clone()
    loop.run_schedule_tc_decoder(lambda i, x: vel / (1.0 + i), dec, last_step=4)
    torch.cuda.synchronize()
    loop.check_flags()
    assert torch.isfinite(loop.theta).all() and torch.isfinite(loop.x_t).all() and torch.isfinite(loop.grad_velocity).all()
    assert not torch.equal(loop.theta, theta0)
    assert torch.isfinite(loop._obj.terms_hoi).all() and torch.isfinite(loop.image_terms).all()
