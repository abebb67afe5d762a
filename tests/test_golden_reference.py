"""The CPU oracle and the host-side mirrors against vectors produced by EXECUTING THE REFERENCE'S
OWN CODE (tests/golden/make_golden.py -> tests/golden/ref_golden.npz; the generator runs only in the
authoring container, these tests only read the committed file).

These are the rows the reference itself can pin offline: scheduler (a2), transforms (a5/a6),
key-point regression (a11), penetration count (a9/a9'), lattice construction, optimiser wiring
(a1/a3/a4 with the real ``get_guidance_params`` + ``OptimizationConfig`` + torch AdamW), and the ICP
loop (a16-a18; trimesh entry points stood in, see the generator's docstring)."""
import os

import numpy as np
import pytest
import torch

from oracle import guidance_oracle as O
from oracle import icp_oracle as IO

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))


# --------------------------------------------------------------------------- a2 scheduler
@pytest.mark.parametrize("N,shift", [(20, 1.0), (50, 1.0), (20, 3.0)])
def test_scheduler_sigmas_and_steps_match_reference(N, shift):
    tag = f"sch_N{N}_s{int(shift)}"
    sig = O.set_timesteps_sigmas(N, shift)
    assert np.array_equal(sig.numpy(), G[tag + "_sigmas"])                       # bit-exact fp32
    assert np.array_equal((sig[:-1] * 1000).numpy(), G[tag + "_timesteps"])
    x = torch.from_numpy(G["sch_x"]); v = torch.from_numpy(G["sch_v"])
    for dt, dn in ((torch.float32, "f32"), (torch.float16, "f16")):
        for j, k in enumerate(G[tag + "_ks"]):
            prev, x1 = O.scheduler_step(x.to(dt), v.to(dt), sig[k], sig[k + 1])
            fin = O.scheduler_step_final(x.to(dt), v.to(dt), sig[k])
            assert prev.dtype == dt
            assert np.array_equal(prev.float().numpy(), G[f"{tag}_{dn}_prev"][j])
            assert np.array_equal(x1.float().numpy(), G[f"{tag}_{dn}_x1"][j])
            assert np.array_equal(fin.float().numpy(), G[f"{tag}_{dn}_final"][j])


def test_host_loop_sigmas_match_reference():
    from followmyhold_b200.guidance.loop import set_timesteps_sigmas
    for N, shift in ((20, 1.0), (50, 1.0), (20, 3.0)):
        assert np.array_equal(set_timesteps_sigmas(N, shift).numpy(), G[f"sch_N{N}_s{int(shift)}_sigmas"])


# --------------------------------------------------------------------------- a5 / a6 / a11
def _quat_from_matrix(R):
    """unit quaternion (wxyz) of a rotation matrix (test helper)."""
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = (R[2, 1] - R[1, 2]) / (4 * w); y = (R[0, 2] - R[2, 0]) / (4 * w); z = (R[1, 0] - R[0, 1]) / (4 * w)
    return np.array([w, x, y, z])


def test_similarity_about_bbox_centre_matches_reference():
    verts = torch.from_numpy(G["a6_verts"]).double()
    RT = G["a6_RT"].astype(np.float64)
    q = _quat_from_matrix(RT[:3, :3]) * 1.7                 # un-normalised on purpose: 2/(q.q) scaling
    theta = torch.tensor(np.concatenate([G["a6_scale"].astype(np.float64), RT[:3, 3], q]))
    out = O.transform_around_center_w_scale(verts, theta)
    assert np.abs(out.numpy() - G["a6_out"]).max() < 2e-6     # reference ran in fp32, oracle here in fp64


def test_hunyuan2moge_matches_reference():
    out = O.transform_hunyuan2moge(torch.from_numpy(G["a6_verts"]), torch.from_numpy(G["a5_T"]))
    assert np.abs(out.numpy() - G["a5_out"]).max() < 1e-6


def test_mano_keypoints_match_reference():
    out = O.mano_vert_to_3dkps(torch.from_numpy(G["a11_verts"]), torch.from_numpy(G["a11_J"]))
    assert out.shape == (21, 3)
    assert np.abs(out.numpy() - G["a11_out"]).max() < 1e-6


# --------------------------------------------------------------------------- a9 / lattice
def test_intersection_losses_match_reference():
    sh = torch.from_numpy(G["a9_sdf_hand"]); so = torch.from_numpy(G["a9_sdf_obj"])
    assert float(O.honerf_intersection_loss(sh, so)) == float(G["a9_count_loss"])
    safe = (torch.relu(-sh) * torch.relu(-so)).mean()       # a9' = the form NS a14 follows
    assert float(safe) == pytest.approx(float(G["a9_safe_loss"]), rel=1e-6)


def test_lattice_matches_reference_grid():
    from followmyhold_b200.guidance import sdf_ops
    xyz, gs, _ = sdf_ops.generate_dense_grid_points(G["grid7_bmin"], G["grid7_bmax"], 5, "ij", 6)
    assert gs == [7, 7, 7] and np.array_equal(xyz, G["grid7_xyz"])
    xyz, gs, _ = sdf_ops.generate_dense_grid_points(np.array([-1.10] * 3), np.array([1.10] * 3), 5, "ij", 64)
    assert gs == list(G["grid65_size"])
    assert np.array_equal(xyz[G["grid65_pick_idx"]], G["grid65_pick"])
    assert np.allclose(xyz.astype(np.float64).sum(0), G["grid65_sum"])
    # the oracle's lattice convention (index [ix,iy,iz], z fastest, linspace(-1.1,1.1,D))
    ax = G["grid65_axis"]
    D = 65
    g = torch.arange(D, dtype=torch.float64)
    back = O.world_to_grid(torch.from_numpy(ax.astype(np.float64)), D)
    assert np.abs(back.numpy() - g.numpy()).max() < 1e-5
    pick = G["grid65_pick_idx"]
    ix, iy, iz = pick // (D * D), (pick // D) % D, pick % D
    assert np.array_equal(np.stack([ax[ix], ax[iy], ax[iz]], -1), G["grid65_pick"])


# --------------------------------------------------------------------------- a1 / a3 / a4
def test_config_matches_reference():
    from followmyhold_b200.guidance.config import OptimizationConfig
    c = OptimizationConfig()
    assert [c.optimization_steps_hand, c.optimization_steps_scale, c.optimization_steps_joint, c.num_inference_steps,
            c.guidance_start_step, c.handopt_start_step, c.guidance_end_step] == list(G["cfg_steps"])
    lrs = ([c.phase1_hand_lrs[k] for k in ("scale", "trans", "rot")] + [c.phase2_hand_lrs[k] for k in ("scale", "trans", "rot")]
           + [c.obj_2half_lrs[k] for k in ("scale", "trans", "rot")] + [c.obj_lrs[k] for k in ("scale", "trans", "rot")]
           + [c.noise_obj_lr1, c.noise_obj_lr2])
    assert lrs == list(G["cfg_lrs"])
    assert [c.obj_guidance_scale, float(c.batch_size), float(c.use_intersection_loss)] == list(G["cfg_misc"])


PHASES = {"p1": dict(lr=[1e-2, 1e-2, 0.5, 0, 0, 0], mask=0b000111, lrv=0.0, wd=0.0),
          "p15": dict(lr=[0, 0, 0, 1e-2, 1e-2, 1e-2], mask=0b111000, lrv=1e-4, wd=0.01),
          "p2": dict(lr=[1e-4, 1e-4, 1e-2, 5e-2, 1e-2, 1e-2], mask=0b111111, lrv=1e-2, wd=0.01)}
GROUP_OF = [0, 1, 1, 1, 2, 2, 2, 2, 3, 4, 4, 4, 5, 5, 5, 5]


@pytest.mark.parametrize("tag", ["p1", "p15", "p2"])
def test_oracle_adamw_matches_reference_optimiser(tag):
    ph = PHASES[tag]
    theta = torch.from_numpy(G["opt_theta0"]).clone(); vel = torch.from_numpy(G["opt_vel0"]).clone()
    m = torch.zeros(16); v = torch.zeros(16); mv = torch.zeros_like(vel); vv = torch.zeros_like(vel)
    for k in range(G["opt_grads_theta"].shape[0]):
        g = torch.from_numpy(G["opt_grads_theta"][k]); gv = torch.from_numpy(G["opt_grads_vel"][k])
        for i in range(16):
            grp = GROUP_OF[i]
            if not (ph["mask"] >> grp) & 1:
                continue
            p, mi, vi = O.adamw_step(theta[i], g[i], m[i], v[i], k + 1, ph["lr"][grp], weight_decay=ph["wd"])
            theta[i], m[i], v[i] = p, mi, vi
        if ph["lrv"] > 0:
            vel, mv, vv = O.adamw_step(vel, gv, mv, vv, k + 1, ph["lrv"], weight_decay=ph["wd"])
        assert torch.allclose(theta, torch.from_numpy(G[f"opt_{tag}_theta"][k]), rtol=2e-6, atol=1e-8)
        assert torch.allclose(vel, torch.from_numpy(G[f"opt_{tag}_vel"][k]), rtol=2e-6, atol=1e-8)


def test_host_optimizer_phase_tables_match_reference_groups():
    """GuidanceOptimizer.set_phase reproduces get_guidance_params' groups (code_utils.py:33-78)."""
    from followmyhold_b200.guidance import engine as E
    from followmyhold_b200.guidance.config import OptimizationConfig

    class _Probe(E.GuidanceOptimizer):
        def __init__(self):                                   # no device, no library: tables only
            self.config = OptimizationConfig()

    for phase, tag in ((1, "p1"), (1.5, "p15"), (2, "p2")):
        o = _Probe(); o.set_phase(phase)
        ref = list(G[f"opt_{tag}_group_lrs"])
        ours = [lr for g, lr in enumerate(o.lr_theta) if (o.mask >> g) & 1] + ([o.lr_velocity] if o.opt_velocity else [])
        assert ours == ref and len(ours) == int(G[f"opt_{tag}_ngroups"])
        assert o.weight_decay == (0.0 if phase == 1 else 0.01)
    with pytest.raises(ValueError):
        _Probe().set_phase(3)


# --------------------------------------------------------------------------- a16 - a18
@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_icp_oracle_matches_reference_loop(tag):
    n_iter, outliers, fixed, mn, mx = G[f"icp_{tag}_kw"]
    src, tgt = G[f"icp_{tag}_src"], G[f"icp_{tag}_tgt"]
    n_out = int(outliers * len(src))                          # mesh_align.py:77,87
    T, cost = IO.icp_points(src, tgt, int(n_iter), n_out, bool(fixed), float(mn), float(mx))
    assert np.abs(T - G[f"icp_{tag}_T"]).max() < 1e-12
    assert abs(cost - float(G[f"icp_{tag}_cost"])) < 1e-14
    if tag == "c":                                            # scale clip active (true scale 0.5 < 0.7)
        assert np.linalg.norm(T[:3, 0]) >= 0.7 - 1e-12


def test_init_transform_and_align_pipeline_match_reference():
    from followmyhold_b200.alignment import mesh_align as MA
    from followmyhold_b200.meshio import PointCloud
    src, tgt = G["init_src"], G["init_tgt"]
    T = MA.compute_init_transform(PointCloud(src.copy()), PointCloud(tgt.copy()), False)
    assert np.abs(T - G["init_T"]).max() < 1e-12
    assert np.abs(MA.compute_init_transform(PointCloud(src.copy()), PointCloud(tgt.copy()), True) - G["init_T_fixed"]).max() < 1e-12
    assert np.abs(IO.init_transform_points(src, tgt) - G["init_T"]).max() < 1e-12
    # a18 on the CPU side: init, coarse, fine, compose -- with the oracle ICP standing in for the kernel
    Tfull, pts = IO.align_points(src, tgt, outliers=0.2, iterations_coarse=50, iterations_fine=100, min_scale=0.7,
                                 max_scale=3.0)
    assert np.abs(Tfull - G["align_T"]).max() < 1e-10
    assert np.abs(pts - G["align_pts"]).max() < 1e-10
