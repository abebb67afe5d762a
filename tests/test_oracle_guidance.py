"""Known-answer tests pinning the CPU oracle (the reference ships none: SURVEY.md §4/§8c)."""
import numpy as np
import pytest
import torch

from followmyhold_b200.synthetic import (cap_boundary_loops, boundary_edges, icosphere, make_guidance_sample,
                                         standin_hand_mesh)
from oracle import guidance_oracle as O


def test_standin_hand_has_mano_counts():
    v, f = standin_hand_mesh()
    assert v.shape == (778, 3) and f.shape == (1538, 3)
    assert len(boundary_edges(f)) == 16
    fc = cap_boundary_loops(f)
    assert fc.shape == (1552, 3) and len(boundary_edges(fc)) == 0


def test_scheduler_closed_form():
    # (i) sigma = linspace(0,1,N), x1 = x + (1-s) v, prev = x + (s'-s) v; last step: s_N = s_{N+1} = 1 => prev == x
    N = 20
    sig = O.set_timesteps_sigmas(N)
    assert sig.shape == (N + 1,) and sig[0] == 0 and sig[-1] == 1 and sig[-2] == 1
    x = torch.randn(4, 7); v = torch.randn(4, 7)
    for k in (0, 5, N - 1):
        prev, x1 = O.scheduler_step(x, v, sig[k], sig[k + 1])
        assert torch.allclose(x1, x + (1 - sig[k]) * v)
        assert torch.allclose(prev, x + (sig[k + 1] - sig[k]) * v)
    prev, _ = O.scheduler_step(x, v, sig[N - 1], sig[N])
    assert torch.equal(prev, x)
    assert torch.equal(O.scheduler_step_final(x, v, sig[3]), O.scheduler_step(x, v, sig[3], sig[4])[1])
    # fp16 model output: math in fp32, cast back (schedulers.py:294-309)
    p16, x16 = O.scheduler_step(x.half(), v.half(), sig[3], sig[4])
    assert p16.dtype == torch.float16 and x16.dtype == torch.float16


def test_transforms_identity_and_composition():
    v = torch.randn(50, 3, dtype=torch.float64)
    ident = torch.tensor([1.0, 0, 0, 0, 1, 0, 0, 0], dtype=torch.float64)
    assert torch.allclose(O.transform_around_center_w_scale(v, ident), v)
    th = torch.tensor([1.3, 0.1, -0.2, 0.05, 0.9, 0.1, -0.3, 0.2], dtype=torch.float64)
    R = O.quaternion_to_matrix(th[4:])
    assert torch.allclose(R @ R.T, torch.eye(3, dtype=torch.float64), atol=1e-12)      # invariant to |q|
    c = (v.min(0)[0] + v.max(0)[0]) / 2
    M = torch.eye(4, dtype=torch.float64)
    M[:3, :3] = th[0] * R
    M[:3, 3] = c + th[1:4] - th[0] * R @ c
    out = O.transform_around_center_w_scale(v, th)
    assert torch.allclose(out, v @ M[:3, :3].T + M[:3, 3], atol=1e-12)
    T = torch.eye(4, dtype=torch.float64); T[:3, :3] = 0.3 * R; T[:3, 3] = torch.tensor([0.1, 0.2, -1.5])
    assert torch.allclose(O.transform_hunyuan2moge(v, T), v @ T[:3, :3].T + T[:3, 3])


def test_trilinear_matches_grid_sample_and_is_exact_on_linear_fields():
    D = 17
    vol = torch.randn(D, D, D, dtype=torch.float64)
    g = torch.rand(200, 3, dtype=torch.float64) * (D + 3) - 2          # includes out-of-range -> border clamp
    ours = O.trilinear_sample(vol, g)
    # grid_sample: x = last dim (iz), y = iy, z = ix; align_corners=True
    norm = g / (D - 1) * 2 - 1
    grid = torch.stack([norm[:, 2], norm[:, 1], norm[:, 0]], -1).view(1, -1, 1, 1, 3)
    ref = torch.nn.functional.grid_sample(vol.view(1, 1, D, D, D), grid, mode="bilinear", padding_mode="border",
                                          align_corners=True).view(-1)
    assert torch.allclose(ours, ref, atol=1e-12)
    ar = torch.arange(D, dtype=torch.float64)
    lin = 0.3 * ar.view(D, 1, 1) - 1.2 * ar.view(1, D, 1) + 0.7 * ar.view(1, 1, D) + 2.0
    gi = torch.rand(100, 3, dtype=torch.float64) * (D - 1)
    assert torch.allclose(O.trilinear_sample(lin, gi), 0.3 * gi[:, 0] - 1.2 * gi[:, 1] + 0.7 * gi[:, 2] + 2.0, atol=1e-12)


def test_trilinear_of_sphere_sdf_second_order():
    errs = []
    for D in (17, 33, 65):
        lin = torch.linspace(-1.1, 1.1, D, dtype=torch.float64)
        X, Y, Z = torch.meshgrid(lin, lin, lin, indexing="ij")
        vol = torch.sqrt(X * X + Y * Y + Z * Z) - 0.6
        p = torch.tensor([[0.31, -0.22, 0.41], [0.5, 0.5, 0.1], [-0.7, 0.2, 0.3]], dtype=torch.float64)
        s = O.trilinear_sample(vol, O.world_to_grid(p, D))
        errs.append((s - (p.norm(dim=1) - 0.6)).abs().max().item())
    assert errs[1] < errs[0] / 2.5 and errs[2] < errs[1] / 2.5


def test_trilinear_grad_wrt_volume_sums_to_upstream():
    D = 9
    vol = torch.randn(D, D, D, dtype=torch.float64, requires_grad=True)
    g = torch.rand(30, 3, dtype=torch.float64) * (D - 1)
    up = torch.randn(30, dtype=torch.float64)
    (O.trilinear_sample(vol, g) * up).sum().backward()
    assert abs(vol.grad.sum().item() - up.sum().item()) < 1e-10       # corner weights sum to one


def test_parity_rule_matches_analytic_sphere_and_open_mesh_rule_is_deterministic():
    v, f = icosphere(3, 10.0)
    c = np.float32(20.3)
    D = 41
    ins = O.raster_parity_inside(v + c, f, D)
    ar = np.arange(D)
    X, Y, Z = np.meshgrid(ar, ar, ar, indexing="ij")
    r = np.sqrt((X - c) ** 2 + (Y - c) ** 2 + (Z - c) ** 2)
    assert ins[r < 9.8].all() and not ins[r > 10.01].any()
    # vertices exactly on lattice coordinates (ties): octahedron with integer vertices -> still a clean solid
    ov = np.array([[10, 10, 4], [10, 10, 16], [4, 10, 10], [16, 10, 10], [10, 4, 10], [10, 16, 10]], np.float32)
    of = np.array([[0, 2, 4], [0, 4, 3], [0, 3, 5], [0, 5, 2], [1, 4, 2], [1, 3, 4], [1, 5, 3], [1, 2, 5]], np.int32)
    ins = O.raster_parity_inside(ov, of, 21)
    l1 = np.abs(np.stack(np.meshgrid(*[np.arange(21)] * 3, indexing="ij"), -1) - 10).sum(-1)
    assert ins[l1 < 6].all() and not ins[l1 > 6].any()


def test_mesh2sdf_matches_analytic_sphere():
    v, f = icosphere(4, 0.5)
    lin = np.linspace(-0.8, 0.8, 9, dtype=np.float32)
    P = np.stack(np.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)
    # lattice units for the sign rule: (p - lo)/step
    step = lin[1] - lin[0]
    ins = O.raster_parity_inside(((v - lin[0]) / step).astype(np.float32), f, 9).reshape(-1)
    sdf = O.mesh2sdf(torch.from_numpy(v).double(), torch.from_numpy(f), torch.from_numpy(P).double(), torch.from_numpy(ins))
    ref = np.linalg.norm(P, axis=1) - 0.5
    assert np.abs(sdf.numpy() - ref).max() < 4e-3        # tessellation error of a subdiv-4 icosphere
    assert ((sdf.numpy() < 0) == (ref < -4e-3))[np.abs(ref) > 4e-3].all()


def test_count_loss_has_zero_gradient_and_matches_reference_form():
    sh = torch.randn(500, requires_grad=True); so = torch.randn(500, requires_grad=True)
    c = O.honerf_intersection_loss(sh, so)
    assert not c.requires_grad
    assert abs(float(c) - float(((sh < 0) & (so < 0)).sum()) / 1000) < 1e-7


@pytest.mark.parametrize("term", ["w_pen", "w_con", "w_ivol", "w_ch", "w_mom", "kp", "all"])
def test_energy_gradients_finite_difference_fp64(term):
    s = make_guidance_sample(24, 512, 3)
    s.hand_faces = torch.from_numpy(cap_boundary_loops(s.hand_faces.numpy()))
    W = O.Weights()
    if term != "all":
        for k in list(vars(W)):
            if k.startswith("w_"):
                setattr(W, k, 0.0)
        if term == "kp":
            W.w_hand, W.w_kp = 1.0, 1.0
        else:
            setattr(W, term, 1.0)
    W.w_int_lo = 0.0
    dt = torch.float64
    hg_fixed = {}

    def E(th, to, sdf):
        o = O.guidance_energy(sdf, s.hand_rest.to(dt), s.hand_faces, s.cloud.to(dt), th, to, s.T_h2m.to(dt),
                              s.obj_center.to(dt), W, j_regressor=s.j_regressor.to(dt), kps_2d=s.kps_2d.to(dt),
                              hand_grid_verts_override=hg_fixed.get("hg"))
        return o

    th0 = s.theta_h.to(dt).clone().requires_grad_(True)
    to0 = s.theta_o.to(dt).clone().requires_grad_(True)
    sdf0 = s.sdf.to(dt).clone().requires_grad_(True)
    out = E(th0, to0, sdf0)
    hg_fixed["hg"] = out["hand_grid"].float().numpy()     # freeze the (piecewise-constant) inside mask
    out = E(th0, to0, sdf0)
    out["total"].backward()
    eps = 1e-6
    for name, leaf in (("h", th0), ("o", to0)):
        for i in range(8):
            d = torch.zeros(8, dtype=dt); d[i] = eps
            if name == "h":
                fd = (E(th0.detach() + d, to0.detach(), sdf0.detach())["total"] - E(th0.detach() - d, to0.detach(), sdf0.detach())["total"]) / (2 * eps)
            else:
                fd = (E(th0.detach(), to0.detach() + d, sdf0.detach())["total"] - E(th0.detach(), to0.detach() - d, sdf0.detach())["total"]) / (2 * eps)
            ag = leaf.grad[i].item()
            assert abs(fd.item() - ag) <= 2e-5 * max(1.0, abs(ag), leaf.grad.abs().max().item()), (term, name, i, fd.item(), ag)
    # dE/dSDF at a few voxels with the largest gradient
    g = sdf0.grad
    idx = torch.topk(g.abs().flatten(), 4).indices
    for k in idx.tolist():
        d = torch.zeros_like(sdf0.detach()).flatten(); d[k] = 1e-5; d = d.view_as(sdf0)
        fd = (E(th0.detach(), to0.detach(), sdf0.detach() + d)["total"] - E(th0.detach(), to0.detach(), sdf0.detach() - d)["total"]) / 2e-5
        assert abs(fd.item() - g.flatten()[k].item()) <= 1e-5 * max(1.0, g.abs().max().item())


def test_knn_and_chamfer_against_ckdtree():
    from scipy.spatial import cKDTree
    a = torch.randn(300, 3, dtype=torch.float64); b = torch.randn(1000, 3, dtype=torch.float64)
    d2, idx = O.knn1_sq(a, b)
    dist, qi = cKDTree(b.numpy()).query(a.numpy())
    assert np.array_equal(idx.numpy(), qi)
    assert np.allclose(d2.numpy(), dist ** 2, atol=1e-12)


def test_mesh_edge_loss_and_keypoints():
    v, f = icosphere(1, 1.0)
    e = O.unique_edges(f)
    assert e.shape[0] == 3 * f.shape[0] // 2
    vt = torch.from_numpy(v).double()
    ref = np.mean([np.sum((v[a] - v[b]) ** 2) for a, b in e])
    assert abs(O.mesh_edge_loss(vt, torch.from_numpy(e)).item() - ref) < 1e-6
    verts = torch.randn(778, 3, dtype=torch.float64)
    J = torch.rand(16, 778, dtype=torch.float64)
    k = O.mano_vert_to_3dkps(verts, J)
    assert k.shape == (21, 3)
    assert torch.allclose(k[0], (J @ verts)[0]) and torch.allclose(k[4], verts[744]) and torch.allclose(k[20], verts[671])
    # a point on the optical axis projects to the image centre
    uv = O.fov_project_screen(torch.tensor([[0.0, 0.0, -2.0], [0.1, 0.0, -2.0], [0.0, 0.1, -2.0]], dtype=torch.float64), 60.0, 512, 512)
    assert torch.allclose(uv[0], torch.tensor([256.0, 256.0], dtype=torch.float64))
    assert uv[1, 0] > 256 and uv[2, 1] < 256      # +x (MoGe) -> right, +y (up) -> smaller row
    assert abs((uv[1, 0] - 256).item() - 256 * 0.1 / (2 * np.tan(np.deg2rad(30)))) < 1e-9


def test_adamw_step_matches_torch_optim():
    torch.manual_seed(0)
    p0 = torch.randn(64); g_seq = [torch.randn(64) for _ in range(5)]
    for wd, cls in ((0.01, torch.optim.AdamW), (0.0, torch.optim.Adam)):
        p = p0.clone().requires_grad_(True)
        opt = cls([p], lr=1e-2, eps=1e-4) if wd == 0.0 else cls([p], lr=1e-2, eps=1e-4)
        q, m, v = p0.clone(), torch.zeros(64), torch.zeros(64)
        for t, g in enumerate(g_seq, 1):
            p.grad = g.clone()
            opt.step()
            q, m, v = O.adamw_step(q, g, m, v, t, 1e-2, weight_decay=wd)
            assert torch.allclose(q, p.detach(), rtol=2e-6, atol=1e-7)
