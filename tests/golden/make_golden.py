#!/usr/bin/env python
"""Generate tests/golden/ref_golden.npz by EXECUTING THE REFERENCE'S OWN CODE.

Run in the authoring container only (needs /root/reference, which does not exist on the GPU
box); the tests read the committed .npz and never this script's inputs:

    python tests/golden/make_golden.py

What is executed from /root/reference (never copied into this repo):

  * ``FlowMatchEulerDiscreteScheduler`` (third_party_patches/hy3dgen/shapegen/schedulers.py)
    -- the whole module is imported from its file; its three ``diffusers`` base classes
    (absent offline) are replaced by empty stand-ins that only record the constructor
    arguments as ``self.config``.  ``set_timesteps`` / ``step`` / ``step_final`` run unmodified.
  * top-level helper functions of third_party_patches/hy3dgen/shapegen/pipelines.py
    (``transform_mesh_around_center_w_scale``, ``transform_hunyuan2moge``,
    ``mano_vert_to_3dkps``, ``safe_intersection_loss``, ``honerf_intersection_loss``,
    ``generate_dense_grid_points``): the module cannot be imported (pytorch3d, kaolin,
    diffusers, kiui ... are absent), so each function's source is cut out of the file with
    ``ast`` at run time and compiled into a namespace that holds torch / numpy.  A 12-line
    ``Meshes`` stand-in gives them ``verts_padded`` / ``verts_packed`` / ``update_padded``.
  * ``get_guidance_params`` (third_party/utilz/code_utils.py) and ``OptimizationConfig``
    (src/foho/configs/guid_config.py): imported from their files, then driven with
    ``torch.optim.Adam/AdamW(eps=1e-4)`` exactly as pipelines.py:1318,1384,1478 do.
  * ``icp`` / ``compute_init_transform`` / ``align_meshes_impl``
    (src/foho/alignment/mesh_align.py): imported from its file with a ``trimesh`` stand-in
    (trimesh and pyvista are absent).  The stand-in restates the five trimesh entry points
    the file touches from their published behaviour -- ``PointCloud``, ``transform_points``,
    ``transformations.translation_matrix/scale_matrix``, ``registration.procrustes``
    (Appendix C of SURVEY.md) -- so these vectors pin the reference's OWN loop (trim rule,
    n_outliers from the sample count, scale renormalise/clip, best-by-pre-update-cost,
    fine @ coarse @ init) but NOT trimesh itself.

Everything is seeded; the file is ~150 KB.
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_golden.npz")


# --------------------------------------------------------------------------- stand-ins
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_diffusers_standin():
    import inspect

    class ConfigMixin:
        pass

    class SchedulerMixin:
        pass

    class BaseOutput:
        pass

    def register_to_config(init):
        def wrapper(self, *a, **kw):
            sig = inspect.signature(init)
            bound = sig.bind(self, *a, **kw)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            self.config = types.SimpleNamespace(**cfg)
            init(self, *a, **kw)
        return wrapper

    class _Log:
        @staticmethod
        def get_logger(name):
            import logging
            return logging.getLogger(name)

    _module("diffusers")
    _module("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    _module("diffusers.schedulers")
    _module("diffusers.schedulers.scheduling_utils", SchedulerMixin=SchedulerMixin)
    _module("diffusers.utils", BaseOutput=BaseOutput, logging=_Log)


class PointCloud:
    """trimesh.PointCloud stand-in: .vertices, apply_transform, export (npy)."""

    def __init__(self, vertices):
        self.vertices = np.asarray(vertices, dtype=np.float64)

    def apply_transform(self, T):
        self.vertices = self.vertices @ T[:3, :3].T + T[:3, 3]
        return self

    def export(self, path):
        np.save(path, self.vertices)


def install_trimesh_standin():
    def transform_points(points, matrix):
        points = np.asanyarray(points, dtype=np.float64)
        return points @ matrix[:3, :3].T + matrix[:3, 3]

    def translation_matrix(direction):
        M = np.eye(4)
        M[:3, 3] = direction[:3]
        return M

    def scale_matrix(factor, origin=None):
        M = np.diag([factor, factor, factor, 1.0])
        if origin is not None:
            M[:3, 3] = np.asarray(origin[:3]) * (1.0 - factor)
        return M

    def procrustes(a, b, weights=None, reflection=True, translation=True, scale=True, return_cost=True):
        a = np.asanyarray(a, dtype=np.float64)
        b = np.asanyarray(b, dtype=np.float64)
        acenter = a.mean(axis=0)
        bcenter = b.mean(axis=0)
        ac = a - acenter
        bc = b - bcenter
        if scale:
            ascale = np.sqrt((ac ** 2).sum() / len(a))
            bscale = np.sqrt((bc ** 2).sum() / len(b))
        else:
            ascale = bscale = 1.0
        target = bc / bscale
        u, s, vh = np.linalg.svd(np.dot(target.T, ac / ascale))
        if reflection:
            R = np.dot(u, vh)
        else:
            R = np.dot(np.dot(u, np.diag([1, 1, np.linalg.det(np.dot(u, vh))])), vh)
        t = bcenter - (bscale / ascale) * np.dot(R, acenter)
        M = np.eye(4)
        M[:3, :3] = (bscale / ascale) * R
        M[:3, 3] = t
        assert not return_cost
        return M

    def load(path, **kw):
        return PointCloud(np.load(path))

    tm = _module("trimesh", PointCloud=PointCloud, transform_points=transform_points, load=load)
    tm.transformations = _module("trimesh.transformations", translation_matrix=translation_matrix,
                                 scale_matrix=scale_matrix)
    tm.registration = _module("trimesh.registration", procrustes=procrustes)
    tm.proximity = _module("trimesh.proximity", closest_point=None)
    tm.sample = _module("trimesh.sample")
    _module("pyvista")


class Meshes:
    """pytorch3d.structures.Meshes stand-in for the three accessors the helpers use."""

    def __init__(self, verts):
        self._v = verts

    def verts_padded(self):
        return self._v.unsqueeze(0)

    def verts_packed(self):
        return self._v

    def update_padded(self, v):
        return Meshes(v.squeeze(0))


def import_from_file(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def extract_functions(path, names):
    """Compile the named top-level ``def``s of a reference file into a fresh namespace."""
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "print": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def rand_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


# --------------------------------------------------------------------------- sections
def golden_scheduler(out):
    install_diffusers_standin()
    S = import_from_file("ref_schedulers", f"{REF}/third_party_patches/hy3dgen/shapegen/schedulers.py")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 48, 8, generator=g)
    v = torch.randn(2, 48, 8, generator=g)
    out["sch_x"] = x.numpy()
    out["sch_v"] = v.numpy()
    for N, shift in ((20, 1.0), (50, 1.0), (20, 3.0)):
        sch = S.FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=shift)
        sch.set_timesteps(N, device="cpu", sigmas=np.linspace(0, 1, N))      # pipelines.py:1187-1193
        tag = f"sch_N{N}_s{int(shift)}"
        out[tag + "_sigmas"] = sch.sigmas.numpy()
        out[tag + "_timesteps"] = sch.timesteps.numpy()
        ks = sorted({1, N // 2 - 1, N // 2, N - 5, N - 1})
        out[tag + "_ks"] = np.asarray(ks)
        for dt, dn in ((torch.float32, "f32"), (torch.float16, "f16")):
            prevs, x1s, finals = [], [], []
            for k in ks:
                sch._step_index = None
                t = sch.timesteps[k]
                fin = sch.step_final(v.to(dt), t, x.to(dt))
                assert sch.step_index == k
                res = sch.step(v.to(dt), t, x.to(dt), return_dict=False)
                assert sch.step_index == k + 1
                prevs.append(res[0].float().numpy()); x1s.append(res[1].float().numpy())
                finals.append(fin.float().numpy())
            out[f"{tag}_{dn}_prev"] = np.stack(prevs)
            out[f"{tag}_{dn}_x1"] = np.stack(x1s)
            out[f"{tag}_{dn}_final"] = np.stack(finals)


def golden_helpers(out):
    ns = extract_functions(f"{REF}/third_party_patches/hy3dgen/shapegen/pipelines.py",
                           ["transform_mesh_around_center_w_scale", "transform_hunyuan2moge", "mano_vert_to_3dkps",
                            "safe_intersection_loss", "honerf_intersection_loss", "generate_dense_grid_points"])
    rng = np.random.default_rng(5)
    # a6: similarity about the bbox centre (pipelines.py:108-118), RT as built at :1482-1486
    verts = torch.from_numpy(rng.normal(size=(60, 3)).astype(np.float32) * np.float32(0.2) + np.float32(0.5))
    R = torch.from_numpy(rand_rotation(rng).astype(np.float32))
    t = torch.from_numpy(rng.normal(size=3).astype(np.float32) * np.float32(0.1))
    RT = torch.eye(4)
    RT[:3, :3] = R
    RT[:3, 3] = t
    scale = torch.tensor([1.3], dtype=torch.float32)
    o = ns["transform_mesh_around_center_w_scale"](Meshes(verts), RT, scale)
    out["a6_verts"] = verts.numpy(); out["a6_RT"] = RT.numpy(); out["a6_scale"] = scale.numpy()
    out["a6_out"] = o.verts_packed().numpy()
    # a5 (pipelines.py:242-250)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = (1.7 * rand_rotation(rng)).astype(np.float32)
    T[:3, 3] = rng.normal(size=3).astype(np.float32)
    o = ns["transform_hunyuan2moge"](Meshes(verts), torch.from_numpy(T))
    out["a5_T"] = T; out["a5_out"] = o.verts_packed().numpy()
    # a11 (pipelines.py:121-135)
    hv = torch.from_numpy(rng.normal(size=(778, 3)).astype(np.float32))
    J = rng.random(size=(16, 778)).astype(np.float32)
    J /= J.sum(1, keepdims=True)
    o = ns["mano_vert_to_3dkps"](Meshes(hv), torch.from_numpy(J), "cpu")
    out["a11_verts"] = hv.numpy(); out["a11_J"] = J; out["a11_out"] = o.numpy()
    # a9 / a9' (pipelines.py:204-239)
    sh = torch.from_numpy(rng.normal(size=4096).astype(np.float32))
    so = torch.from_numpy(rng.normal(size=4096).astype(np.float32))
    out["a9_sdf_hand"] = sh.numpy(); out["a9_sdf_obj"] = so.numpy()
    out["a9_count_loss"] = np.asarray(float(ns["honerf_intersection_loss"](sh, so)))
    out["a9_safe_loss"] = np.asarray(float(ns["safe_intersection_loss"](sh, so)), dtype=np.float32)
    # lattice (pipelines.py:341-360), called as at :1126-1137 and kaolin_sdf_ops.py:146-152
    xyz, gs, length = ns["generate_dense_grid_points"](bbox_min=np.array([-1.10, -1.10, -1.10]),
                                                      bbox_max=np.array([1.10, 1.10, 1.10]), octree_depth=5,
                                                      octree_resolution=64, indexing="ij")
    out["grid65_size"] = np.asarray(gs)
    out["grid65_axis"] = xyz.reshape(65, 65, 65, 3)[:, 0, 0, 0].copy()
    pick = np.array([0, 1, 64, 65, 4224, 4225, 137312, 274624])
    out["grid65_pick_idx"] = pick; out["grid65_pick"] = xyz[pick]
    out["grid65_sum"] = xyz.astype(np.float64).sum(0)
    bmin = np.array([-0.31, 0.05, -1.2], dtype=np.float32); bmax = np.array([0.44, 0.61, -0.7], dtype=np.float32)
    xyz, gs, _ = ns["generate_dense_grid_points"](bbox_min=bmin, bbox_max=bmax, octree_depth=5, octree_resolution=6,
                                                  indexing="ij")
    out["grid7_bmin"] = bmin; out["grid7_bmax"] = bmax; out["grid7_xyz"] = xyz


def golden_optimizer(out):
    cu = import_from_file("ref_code_utils", f"{REF}/third_party/utilz/code_utils.py")
    gc = import_from_file("ref_guid_config", f"{REF}/src/foho/configs/guid_config.py")
    cfg = gc.OptimizationConfig()()
    out["cfg_steps"] = np.asarray([cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint,
                                   cfg.num_inference_steps, cfg.guidance_start_step, cfg.handopt_start_step,
                                   cfg.guidance_end_step])
    out["cfg_lrs"] = np.asarray([cfg.phase1_hand_lrs[k] for k in ("scale", "trans", "rot")]
                                + [cfg.phase2_hand_lrs[k] for k in ("scale", "trans", "rot")]
                                + [cfg.obj_2half_lrs[k] for k in ("scale", "trans", "rot")]
                                + [cfg.obj_lrs[k] for k in ("scale", "trans", "rot")]
                                + [cfg.noise_obj_lr1, cfg.noise_obj_lr2])
    out["cfg_misc"] = np.asarray([cfg.obj_guidance_scale, float(cfg.batch_size), float(cfg.use_intersection_loss)])
    g = torch.Generator().manual_seed(3)
    L = 512
    theta0 = torch.tensor([1.0, 0.01, -0.02, 0.03, 0.9, 0.1, -0.2, 0.05,
                           1.1, -0.03, 0.02, 0.01, 1.0, -0.1, 0.2, 0.3])
    vel0 = torch.randn(1, L, generator=g) * 0.1
    n_steps = 4
    grads_theta = torch.randn(n_steps, 16, generator=g) * torch.logspace(-4, 1, 16)
    grads_vel = torch.randn(n_steps, 1, L, generator=g) * 1e-3
    out["opt_theta0"] = theta0.numpy(); out["opt_vel0"] = vel0.numpy()
    out["opt_grads_theta"] = grads_theta.numpy(); out["opt_grads_vel"] = grads_vel.numpy()
    kw = dict(phase1_hand_lrs=cfg.phase1_hand_lrs, phase2_hand_lrs=cfg.phase2_hand_lrs, noise_obj_lr1=cfg.noise_obj_lr1,
              noise_obj_lr2=cfg.noise_obj_lr2, obj_lrs=cfg.obj_lrs, obj_2half_lrs=cfg.obj_2half_lrs)
    for phase, tag in ((1, "p1"), (1.5, "p15"), (2, "p2")):
        sh, th, qh = theta0[0:1].clone(), theta0[1:4].clone(), theta0[4:8].clone()
        so, to, qo = theta0[8:9].clone(), theta0[9:12].clone(), theta0[12:16].clone()
        groups, v_opt, sh, th, qh, so, to, qo = cu.get_guidance_params(
            phase, vel0.clone(), sh, th, qh, "cpu", scale_obj=so, trans_obj=to, rotation_obj=qo, **kw)
        # pipelines.py:1318 (Adam, phase 1), :1384 / :1478 (AdamW, phases 1.5 / 2)
        opt = torch.optim.Adam(groups, eps=1e-4) if phase == 1 else torch.optim.AdamW(groups, eps=1e-4)
        leaves = [sh, th, qh, so, to, qo]
        slices = [slice(0, 1), slice(1, 4), slice(4, 8), slice(8, 9), slice(9, 12), slice(12, 16)]
        th_hist, v_hist = [], []
        for k in range(n_steps):
            opt.zero_grad()
            for leaf, sl in zip(leaves, slices):
                if leaf.requires_grad:
                    leaf.grad = grads_theta[k, sl].clone()
            if v_opt.requires_grad:
                v_opt.grad = grads_vel[k].clone()
            opt.step()
            th_hist.append(torch.cat([l.detach() for l in leaves]).numpy().copy())
            v_hist.append(v_opt.detach().numpy().copy())
        out[f"opt_{tag}_theta"] = np.stack(th_hist)
        out[f"opt_{tag}_vel"] = np.stack(v_hist)
        out[f"opt_{tag}_ngroups"] = np.asarray(len(groups))
        out[f"opt_{tag}_group_lrs"] = np.asarray([g_["lr"] for g_ in groups])


def golden_icp(out):
    install_trimesh_standin()
    ma = import_from_file("ref_mesh_align", f"{REF}/src/foho/alignment/mesh_align.py")
    import tqdm as _tq
    ma.tqdm = lambda it, **kw: it          # silence progress bars
    rng = np.random.default_rng(17)

    def make_case(ns, nt, true_scale, noise, n_out):
        tgt = rng.normal(size=(nt, 3)) * np.array([1.0, 0.6, 0.3])
        R = rand_rotation(rng) if true_scale != 1.0 else np.eye(3)
        # small rotation so plain ICP converges
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        ang = 0.25
        K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        t = rng.normal(size=3) * 0.1
        src = ((tgt[:ns] - t) @ R) / true_scale + noise * rng.normal(size=(ns, 3))
        if n_out:
            src[:n_out] += rng.normal(size=(n_out, 3)) * 1.5
        return src, tgt

    cases = [
        ("a", dict(ns=300, nt=1000, true_scale=1.2, noise=0.005, n_out=40), dict(n_iter=30, outliers=0.2, min_scale=0.7, max_scale=3.0)),
        ("b", dict(ns=257, nt=777, true_scale=1.0, noise=0.01, n_out=0), dict(n_iter=20, outliers=0.0, fixed_scale=True, min_scale=0.7, max_scale=3.0)),
        ("c", dict(ns=200, nt=600, true_scale=0.5, noise=0.005, n_out=20), dict(n_iter=25, outliers=0.15, min_scale=0.7, max_scale=3.0)),
        ("d", dict(ns=128, nt=400, true_scale=1.05, noise=0.0, n_out=0), dict(n_iter=12, outliers=0.0)),   # icp() defaults 0.5/2.0
    ]
    for tag, mk, kw in cases:
        src, tgt = make_case(**mk)
        T, cost = ma.icp(PointCloud(src.copy()), PointCloud(tgt.copy()), **kw)
        out[f"icp_{tag}_src"] = src; out[f"icp_{tag}_tgt"] = tgt
        out[f"icp_{tag}_T"] = np.asarray(T); out[f"icp_{tag}_cost"] = np.asarray(cost)
        out[f"icp_{tag}_kw"] = np.asarray([kw["n_iter"], kw.get("outliers", 0.0), float(kw.get("fixed_scale", False)),
                                           kw.get("min_scale", 0.5), kw.get("max_scale", 2.0)])
    # a17 compute_init_transform on point clouds (mesh_align.py:18-35)
    src, tgt = make_case(ns=150, nt=500, true_scale=1.6, noise=0.0, n_out=0)
    src = src * 0.8 + np.array([0.4, -0.2, 0.9])
    out["init_src"] = src; out["init_tgt"] = tgt
    out["init_T"] = ma.compute_init_transform(PointCloud(src), PointCloud(tgt), False)
    out["init_T_fixed"] = ma.compute_init_transform(PointCloud(src), PointCloud(tgt), True)
    # a18 align_meshes_impl end to end on point clouds, both callers' hyper-parameters (h2m.py:35-54)
    with tempfile.TemporaryDirectory() as td:
        sp, tp = os.path.join(td, "s.npy"), os.path.join(td, "t.npy")
        np.save(sp, src); np.save(tp, tgt)
        tpath, mpath = os.path.join(td, "T.npy"), os.path.join(td, "m.npy")
        ma.print = lambda *a, **k: None
        ma.align_meshes_impl(sp, tp, tpath, mpath, fixed_scale=False, outliers=0.2, test_rotations=False,
                             test_reflections=False, on_surface=False, iterations_coarse=50, count_source_coarse=1000,
                             count_target_coarse=5000, iterations_fine=100, count_source_fine=5000,
                             count_target_fine=10000, min_scale=0.7, max_scale=3.0, plot=False)
        out["align_T"] = np.load(tpath)
        out["align_pts"] = np.load(mpath)


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs /root/reference (authoring container only)")
    out = {}
    golden_scheduler(out)
    golden_helpers(out)
    golden_optimizer(out)
    golden_icp(out)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(out)} arrays, {os.path.getsize(OUT) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
