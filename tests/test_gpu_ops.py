"""GPU parity of the smaller C-ABI entry points: mesh2sdf + count (REF a8/a9), fused
AdamW/step_final update (REF a2+a4), scheduler step, and the graph-captured loop."""
import numpy as np
import pytest
import torch

from followmyhold_b200.synthetic import cap_boundary_loops, icosphere, make_guidance_sample, stack_samples, standin_hand_mesh

pytestmark = pytest.mark.gpu


def test_mesh2sdf_and_count_match_oracle():
    from followmyhold_b200.guidance import sdf_ops
    from oracle import guidance_oracle as O
    hv, hf = standin_hand_mesh(0.35)
    hf = cap_boundary_loops(hf)
    sv, sf = icosphere(3, 0.12)
    sv = sv * np.array([1.0, 0.8, 0.6], np.float32) + np.array([0.05, 0.02, 0.0], np.float32)
    dev = "cuda:0"
    m1 = (torch.from_numpy(hv).to(dev), torch.from_numpy(hf).to(dev))
    m2 = (torch.from_numpy(sv).to(dev), torch.from_numpy(sf).to(dev))
    res = 32
    s1, s2 = sdf_ops.get_sdf_of_meshes(m1, m2, dev, res)
    cnt = sdf_ops.honerf_intersection_loss(s1, s2)
    torch.cuda.synchronize()
    lo = np.minimum(hv.min(0), sv.min(0)); hi = np.maximum(hv.max(0), sv.max(0))
    axes = [np.linspace(lo[a], hi[a], res + 1, dtype=np.float32) for a in range(3)]
    P = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    for (v, f), got in (((hv, hf), s1), ((sv, sf), s2)):
        d, _, _ = O.point_mesh_distance(torch.from_numpy(P).double(), torch.from_numpy(v).double(), torch.from_numpy(f))
        g = got.cpu().numpy()
        assert np.abs(np.abs(g) - d.numpy()).max() < 2e-6
        # sign: the oracle rule evaluated at the same float32 lattice (rectilinear form of the parity rule)
        step = [(ax[-1] - ax[0]) / res for ax in axes]
    # sphere: analytic inside test away from the surface
    r = np.linalg.norm((P - np.array([0.05, 0.02, 0.0])) / (0.12 * np.array([1.0, 0.8, 0.6])), axis=1)
    g2 = s2.cpu().numpy()
    assert (g2[r < 0.9] < 0).all() and (g2[r > 1.05] > 0).all()
    n_both = int(((s1 < 0) & (s2 < 0)).sum())
    assert n_both > 0
    assert int(sdf_ops.intersection_count(s1, s2)[0]) == n_both                 # integer work: exact
    # the reference returns int64.sum()/1000, a float32 tensor (pipelines.py:237); torch's CUDA and CPU
    # division kernels differ in the last ulp, so the float is compared to 1 ulp
    assert float(cnt) == pytest.approx(n_both / 1000, rel=2e-7)
    assert float(O.honerf_intersection_loss(s1.cpu(), s2.cpu())) == pytest.approx(n_both / 1000, rel=2e-7)


def test_fused_update_matches_torch_adamw_and_step_final():
    from followmyhold_b200.guidance.engine import GuidanceOptimizer, scheduler_step
    from oracle import guidance_oracle as O
    B, L = 3, 4096
    dev = "cuda:0"
    torch.manual_seed(0)
    theta = torch.randn(B, 16, device=dev); v = torch.randn(B, L, device=dev); x_t = torch.randn(B, L, device=dev)
    theta_ref = theta.clone().cpu(); v_ref = v.clone().cpu()
    opt = GuidanceOptimizer(B, L, device=dev)
    opt.set_phase(2); opt.reset()
    lr_t = torch.tensor(sum([[opt.lr_theta[0]], [opt.lr_theta[1]] * 3, [opt.lr_theta[2]] * 4, [opt.lr_theta[3]],
                             [opt.lr_theta[4]] * 3, [opt.lr_theta[5]] * 4], []))
    params = [theta_ref[:, i:i + 1].clone().requires_grad_(True) for i in range(16)]
    pv = v_ref.clone().requires_grad_(True)
    topt = torch.optim.AdamW([{"params": [p], "lr": float(lr_t[i])} for i, p in enumerate(params)] +
                             [{"params": [pv], "lr": opt.lr_velocity}], eps=1e-4)
    x1 = torch.empty_like(v)
    sigma = 0.7
    for step in range(4):
        gt = torch.randn(B, 16, device=dev); gv = torch.randn(B, L, device=dev)
        opt.step(theta, gt, v, gv, x_t, x1, sigma=sigma)
        for i, p in enumerate(params):
            p.grad = gt[:, i:i + 1].cpu().clone()
        pv.grad = gv.cpu().clone()
        topt.step()
        torch.cuda.synchronize()
        ref_theta = torch.cat([p.detach() for p in params], 1)
        assert torch.allclose(theta.cpu(), ref_theta, rtol=3e-6, atol=1e-7)
        assert torch.allclose(v.cpu(), pv.detach(), rtol=3e-6, atol=1e-7)
        assert torch.allclose(x1.cpu(), O.scheduler_step_final(x_t.cpu(), v.cpu(), sigma), rtol=1e-6, atol=1e-7)
    # phase 1: Adam without weight decay, hand groups only
    opt.set_phase(1); opt.reset()
    th0 = theta.clone()
    opt.step(theta, torch.ones_like(theta))
    torch.cuda.synchronize()
    assert torch.equal(theta[:, 8:], th0[:, 8:]) and not torch.equal(theta[:, :8], th0[:, :8])
    # scheduler.step
    prev, px1 = scheduler_step(x_t, v, 0.3, 0.4)
    rp, rx = O.scheduler_step(x_t.cpu(), v.cpu(), 0.3, 0.4)
    assert torch.allclose(prev.cpu(), rp, atol=1e-6) and torch.allclose(px1.cpu(), rx, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
@pytest.mark.parametrize("phase", [2, 1.5])
def test_fused_update_bit_equal_to_oracle_and_torch_cuda_optimiser(dtype, phase):
    """k_update / k_update_f16 against (1) the op-exact oracle in ATen's CUDA grouping: every bit of the
    parameters, both moments and the fused step_final output, for float32 and for the reference's half
    velocity leaf (code_utils.py:43-78); (2) torch.optim.AdamW stepping CUDA parameters -- the optimiser
    object the reference constructs (pipelines.py:1384,1478), multi-tensor path."""
    from followmyhold_b200.guidance.engine import GROUPS, GuidanceOptimizer
    from oracle import guidance_oracle as O
    B, L, steps, sigma = 2, 8192, 6, 0.4375
    dev = "cuda:0"
    g = torch.Generator().manual_seed(11)
    theta = torch.randn(B, 16, generator=g).to(dev)
    v = (0.5 * torch.randn(B, L, generator=g)).to(dtype).to(dev)
    x_t = torch.randn(B, L, generator=g).to(dtype).to(dev)
    x1 = torch.empty_like(v)
    opt = GuidanceOptimizer(B, L, device=dev, velocity_dtype=dtype)
    opt.set_phase(phase); opt.reset()
    grp_of = [0, 1, 1, 1, 2, 2, 2, 2, 3, 4, 4, 4, 5, 5, 5, 5]
    active = [k for k in range(16) if (opt.mask >> grp_of[k]) & 1]
    # oracle state
    npdt = np.float16 if dtype == torch.float16 else np.float32
    ov = v.cpu().numpy().copy(); om = np.zeros_like(ov); ovv = np.zeros_like(ov)
    oth = theta.cpu().numpy().copy(); otm = np.zeros_like(oth); otv = np.zeros_like(oth)
    # torch's own optimiser on CUDA parameters, grouped like get_guidance_params (code_utils.py:57-78)
    tp = [theta[:, k:k + 1].clone().requires_grad_(True) for k in range(16)]
    tv = v.clone().requires_grad_(True)
    topt = torch.optim.AdamW([{"params": [tp[k]], "lr": opt.lr_theta[grp_of[k]]} for k in active] +
                             [{"params": [tv], "lr": opt.lr_velocity}], eps=1e-4)
    for k in range(steps):
        gt = torch.randn(B, 16, generator=g).to(dev)
        gv = (torch.randn(B, L, generator=g) * (0.05 if k % 2 else 2.0)).to(dtype).to(dev)
        opt.step(theta, gt, v, gv, x_t, x1, sigma=sigma)
        for j in active:
            tp[j].grad = gt[:, j:j + 1].clone()
        tv.grad = gv.clone()
        topt.step()
        ov, om, ovv = O.adamw_step_torch_ops(ov, gv.cpu().numpy(), om, ovv, k + 1, opt.lr_velocity, order="cuda")
        for j in active:
            oth[:, j], otm[:, j], otv[:, j] = O.adamw_step_torch_ops(
                oth[:, j], gt[:, j].cpu().numpy(), otm[:, j], otv[:, j], k + 1, opt.lr_theta[grp_of[j]], order="cuda")
        torch.cuda.synchronize()
        assert np.array_equal(v.cpu().numpy(), ov), f"velocity differs from the oracle at step {k + 1}"
        assert np.array_equal(opt.vel_m.cpu().numpy(), om) and np.array_equal(opt.vel_v.cpu().numpy(), ovv)
        assert np.array_equal(x1.cpu().numpy(), O.step_final_torch_ops(x_t.cpu().numpy(), ov, sigma))
        assert np.array_equal(theta.cpu().numpy(), oth)
        assert np.array_equal(opt.theta_m.cpu().numpy(), otm) and np.array_equal(opt.theta_v.cpu().numpy(), otv)
    assert v.dtype == dtype and opt.vel_m.dtype == dtype and ov.dtype == npdt
    # torch's CUDA optimiser (whatever build is installed): agreement to the last bit or two is required, the
    # count of elements that are not bit-equal is recorded (gpurun_out/adamw_vs_torch_cuda.jsonl) and reported
    # in DESIGN.md -- a torch build may group an op differently from the ATen sources the oracle restates
    ref_v = tv.detach()
    ref_t = torch.cat([tp[j].detach() if j in active else theta[:, j:j + 1] for j in range(16)], 1)
    rec = {"dtype": str(dtype), "phase": phase, "steps": steps, "velocity_mismatch": int((v != ref_v).sum()),
           "velocity_elems": v.numel(), "leaf_mismatch": int((theta != ref_t).sum()), "leaf_elems": theta.numel(),
           "velocity_max_abs_diff": float((v.float() - ref_v.float()).abs().max()), "torch": torch.__version__}
    print("AdamW vs torch CUDA optimiser:", rec)
    import json, os
    if os.path.isdir("gpurun_out"):                      # scratch directory of a GPU-box run: keep the record
        with open("gpurun_out/adamw_vs_torch_cuda.jsonl", "a") as f:
            f.write(json.dumps(rec) + "\n")
    ulp = float(torch.finfo(dtype).eps) * float(ref_v.abs().max())
    assert rec["velocity_max_abs_diff"] <= 4 * ulp
    assert torch.allclose(theta, ref_t, rtol=3e-6, atol=2e-7)


def test_fused_update_f16_rejects_misuse():
    from followmyhold_b200 import _lib
    from followmyhold_b200.guidance.engine import GuidanceOptimizer
    opt = GuidanceOptimizer(1, 64, device="cuda:0", velocity_dtype=torch.float16)
    th = torch.zeros(1, 16, device="cuda:0"); v32 = torch.zeros(1, 64, device="cuda:0")
    with pytest.raises(ValueError):
        opt.step(th, th.clone(), v32, v32.clone())                      # dtype mismatch caught on the host side
    opt6 = GuidanceOptimizer(1, 60, device="cuda:0", velocity_dtype=torch.float16)    # L not a multiple of 8
    v16 = torch.zeros(1, 60, device="cuda:0", dtype=torch.float16)
    with pytest.raises(_lib.FohoStatusError):
        opt6.step(th, th.clone(), v16, v16.clone())
    with pytest.raises(ValueError):
        GuidanceOptimizer(1, 64, device="cuda:0", velocity_dtype=torch.bfloat16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_update_nan_guard_leaves_the_sample_untouched(dtype):
    """pipelines.py:1590-1592: `if torch.isnan(total_loss): break` comes before backward()/step().  On the device:
    the sample whose total is NaN is not updated by this or any later step of the outer step, its flag holds the
    optimiser step of the first NaN, the other samples advance exactly as without the guard."""
    from followmyhold_b200 import _lib
    from followmyhold_b200.guidance.engine import GuidanceOptimizer
    B, L, dev = 3, 1024, "cuda:0"
    g = torch.Generator().manual_seed(4)
    theta0 = torch.randn(B, 16, generator=g).to(dev)
    v0 = torch.randn(B, L, generator=g).to(dtype).to(dev)
    x_t = torch.randn(B, L, generator=g).to(dtype).to(dev)
    runs = []
    for guarded in (True, False):
        theta, v, x1 = theta0.clone(), v0.clone(), torch.zeros_like(v0)
        opt = GuidanceOptimizer(B, L, device=dev, velocity_dtype=dtype)
        opt.set_phase(2); opt.reset()
        flag = torch.zeros(B, dtype=torch.int32, device=dev)
        gg = torch.Generator().manual_seed(5)
        for k in range(4):
            gt = torch.randn(B, 16, generator=gg).to(dev); gv = torch.randn(B, L, generator=gg).to(dtype).to(dev)
            terms = torch.rand(B, _lib.FOHO_NUM_TERMS, generator=gg).to(dev)
            if k == 1:
                terms[1, 0] = float("nan")              # sample 1: NaN total at the 2nd inner iteration only
            if guarded:
                opt.step(theta, gt, v, gv, x_t, x1, sigma=0.5, terms=terms, nan_flag=flag)
            else:
                opt.step(theta, gt, v, gv, x_t, x1, sigma=0.5)
            if k == 0:
                after_first = (theta.clone(), v.clone(), x1.clone(), opt.vel_m.clone(), opt.theta_v.clone())
        torch.cuda.synchronize()
        runs.append((theta, v, x1, opt, flag))
    (th_g, v_g, x1_g, opt_g, flag), (th_u, v_u, x1_u, opt_u, _) = runs
    assert flag.tolist() == [0, 2, 0]
    for b in (0, 2):                                     # untouched by the guard: bit-equal to the unguarded run
        assert torch.equal(th_g[b], th_u[b]) and torch.equal(v_g[b], v_u[b]) and torch.equal(x1_g[b], x1_u[b])
        assert torch.equal(opt_g.vel_v[b], opt_u.vel_v[b]) and torch.equal(opt_g.theta_m[b], opt_u.theta_m[b])
    # sample 1 stopped after the first iteration: parameters, moments and x1 are those of iteration 1
    # (`after_first` was taken in the unguarded run, whose first iteration is identical)
    th1, v1, x11, m1, tv1 = after_first
    assert torch.equal(th_g[1], th1[1]) and torch.equal(v_g[1], v1[1]) and torch.equal(x1_g[1], x11[1])
    assert torch.equal(opt_g.vel_m[1], m1[1]) and torch.equal(opt_g.theta_v[1], tv1[1])
    assert not torch.equal(v_u[1], v1[1])                # ... whereas without the guard it kept moving
    # the guard needs both pointers
    with pytest.raises(ValueError):
        opt_g.step(th_g, th_g.clone(), v_g, v_g.clone(), terms=torch.zeros(B, 16, device=dev))
    d = _lib.UpdateDesc()
    d.B, d.L, d.step = B, L, 1
    d.theta = d.grad_theta = d.theta_m = d.theta_v = th_g.data_ptr()
    d.terms = th_g.data_ptr()                            # nan_flag missing
    import ctypes
    assert _lib.load().foho_guidance_update(ctypes.byref(d), None) == -1          # FOHO_E_NULL


def test_graph_loop_matches_eager_and_host_api():
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.loop import GuidanceLoop
    B, D, P = 2, 64, 1024
    samples = [make_guidance_sample(D, P, 40 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, cap=True)
    cfg = OptimizationConfig(); cfg.optimization_steps_joint = 5
    loops = [GuidanceLoop(B, D, st, P, config=cfg, seed=1) for _ in range(2)]
    g = torch.Generator().manual_seed(0)
    x_t = torch.randn(B, loops[0].L, generator=g); vel = 0.1 * torch.randn(B, loops[0].L, generator=g)
    # eager: enqueue the step kernels directly
    lp = loops[0]
    lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.x_t.copy_(x_t); lp.velocity.copy_(vel); lp.theta.copy_(theta0)
    lp._enqueue_step(12, torch.cuda.current_stream())
    torch.cuda.synchronize()
    # graph + host API
    out = loops[1].denoise_step_host(12, sdf0.cpu().pin_memory(), x_t.pin_memory(), vel.pin_memory(), theta0.cpu().pin_memory())
    assert torch.allclose(out["theta"], lp.theta.cpu(), rtol=1e-4, atol=1e-6)
    assert torch.allclose(out["velocity"], lp.velocity.cpu(), rtol=1e-4, atol=1e-6)
    assert torch.allclose(out["prev_sample"], lp.prev.cpu(), rtol=1e-4, atol=1e-6)
    assert torch.isfinite(out["terms"]).all()
    # the optimiser actually moved the leaves and reduced nothing to NaN
    assert not torch.equal(out["theta"], theta0.cpu())
    sig = loops[1].sigmas
    assert torch.allclose(out["prev_sample"], x_t + (sig[13] - sig[12]) * out["velocity"], atol=1e-5)


@pytest.mark.parametrize("phase", [2, 1.5])
def test_loop_nan_guard_and_loss_history(phase):
    """An image whose total is NaN (here: a NaN 2-D key-point target) leaves its inner loop at once -- nothing of
    it moves, the step is reported -- while the other image of the batch is optimised exactly as if it were
    alone; in the object-only phase that image counts as failed (``return None``, pipelines.py:1442-1444).
    ``loss_log_every`` keeps the terms of every n-th inner iteration (the FOHO_DEBUG_DIR loss log)."""
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.loop import GuidanceLoop
    B, D, P = 2, 32, 512
    samples = [make_guidance_sample(D, P, 70 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, cap=True)
    cfg = OptimizationConfig()
    cfg.optimization_steps_joint, cfg.optimization_steps_scale, cfg.optimization_steps_hand = 5, 5, 5
    import dataclasses
    kps_bad = st.kps_2d.clone(); kps_bad[1, 3, 0] = float("nan")
    st_bad = dataclasses.replace(st, kps_2d=kps_bad)
    step = cfg.handopt_start_step + (2 if phase == 2 else 1)
    g = torch.Generator().manual_seed(0)
    x_t = torch.randn(B, 1024, generator=g); vel = 0.1 * torch.randn(B, 1024, generator=g)
    res = []
    for poison in (True, False):
        lp = GuidanceLoop(B, D, st_bad if poison else st, P, config=cfg, latent_elems=1024, seed=1, loss_log_every=2)
        lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.x_t.copy_(x_t); lp.velocity.copy_(vel); lp.theta.copy_(theta0)
        lp._enqueue_step(step, torch.cuda.current_stream(), phase)
        torch.cuda.synchronize()
        res.append(lp)
    bad, good = res
    # the other image is optimised as if it were alone (two runs agree to the order of the float scatter-adds)
    for a, b in ((bad.theta[0], good.theta[0]), (bad.velocity[0], good.velocity[0]), (bad.prev[0], good.prev[0])):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert not torch.equal(good.velocity[0], vel[0].cuda())
    assert torch.equal(bad.velocity[1], vel[1].cuda())                           # never updated
    assert torch.equal(bad.theta[1], theta0[1].cuda())
    assert not torch.equal(good.velocity[1], vel[1].cuda())
    assert bad.nan_flag.tolist() == [0, 1] and good.nan_flag.tolist() == [0, 0]
    assert bad.nan_report() == {step: {1: 0}} and good.nan_report() == {}
    assert bad.failed_images() == ([1] if phase == 1.5 else []) and good.failed_images() == []
    # loss history: iterations 0, 2, 4 of this step; finite for the clean image, total NaN for the poisoned one
    h = good.loss_history[step]
    assert h.shape[0] == 3 and torch.isfinite(h[:, :, 0]).all()
    assert torch.equal(h[2], good.terms)                                         # last logged = last evaluation (k = 4)
    assert torch.isnan(bad.loss_history[step][:, 1, 0]).all() and torch.isfinite(bad.loss_history[step][:, 0, 0]).all()
    lines = good.loss_log_lines(0)
    assert len(lines) == 3 and lines[1].startswith(f"Denoise step {step} phase {phase}, Opt step 2, total: ")


def test_pipelined_host_api_matches_single_batch_calls():
    """``denoise_steps_host`` (upload / compute / download on three streams) returns, batch by batch,
    what one ``denoise_step_host`` call per batch returns -- different inputs per batch, in order."""
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.loop import GuidanceLoop
    B, D, P = 2, 64, 512
    samples = [make_guidance_sample(D, P, 60 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, cap=True)
    cfg = OptimizationConfig(); cfg.optimization_steps_joint = 4
    lp = GuidanceLoop(B, D, st, P, config=cfg, seed=2)
    g = torch.Generator().manual_seed(3)
    batches = []
    for k in range(3):
        x_t = torch.randn(B, lp.L, generator=g).pin_memory()
        vel = (0.1 * (k + 1) * torch.randn(B, lp.L, generator=g)).pin_memory()
        th = theta0.cpu().clone(); th[:, 1] += 0.01 * k
        batches.append((sdf0.cpu().pin_memory(), x_t, vel, th.pin_memory()))
    single = [{n: t.clone() for n, t in lp.denoise_step_host(12, *b).items()} for b in batches]
    piped = lp.denoise_steps_host(12, batches)
    assert len(piped) == 3
    for k in range(3):
        for n in ("theta", "velocity", "prev_sample", "terms"):
            assert torch.allclose(piped[k][n], single[k][n], rtol=1e-4, atol=1e-6), (k, n)
    assert not torch.allclose(piped[0]["velocity"], piped[1]["velocity"])
    # volumes handed over as fp16 (the decoder's output dtype) are widened on the device: same results as
    # the fp32 call on the fp16-rounded values; a None volume keeps the resident one
    h16 = sdf0.cpu().half().pin_memory()
    rounded = h16.float().pin_memory()
    a = lp.denoise_steps_host(12, [(h16,) + batches[0][1:]])[0]
    a = {n: t.clone() for n, t in a.items()}
    b = lp.denoise_steps_host(12, [(rounded,) + batches[0][1:], (None,) + batches[0][1:]])
    for n in ("theta", "velocity", "prev_sample", "terms"):
        assert torch.allclose(a[n], b[0][n], rtol=1e-4, atol=1e-6), n
        assert torch.allclose(b[0][n], b[1][n], rtol=1e-4, atol=1e-6), n
    # a resident base volume changed in place between two calls is picked up (the decode volume is re-based only then)
    b1 = {n: t.clone() for n, t in b[1].items()}
    lp.sdf0.add_(0.05)
    c = lp.denoise_steps_host(12, [(None,) + batches[0][1:]])[0]
    c = {n: t.clone() for n, t in c.items()}
    d = lp.denoise_steps_host(12, [((rounded + 0.05).pin_memory(),) + batches[0][1:]])[0]
    for n in ("theta", "velocity", "prev_sample", "terms"):
        assert torch.allclose(c[n], d[n], rtol=1e-4, atol=1e-6), n
    assert not torch.allclose(c["terms"], b1["terms"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("m", [2, 4])
def test_micro_batches_do_not_change_results(m):
    """The batch advanced as m independent lanes inside the step graph gives what one lane gives:
    images never interact (pipelines.py processes them one at a time)."""
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.loop import GuidanceLoop, slice_statics
    B, D, P = 4, 64, 1024
    samples = [make_guidance_sample(D, P, 120 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, cap=True)
    half = slice_statics(st, 2, 2)
    assert half.hand_rest.shape[0] == 2 and torch.equal(half.cloud, st.cloud[2:4]) and half.hand_faces is st.hand_faces
    cfg = OptimizationConfig(); cfg.optimization_steps_joint = 4
    g = torch.Generator().manual_seed(7)
    res = []
    for mb in (1, m):
        lp = GuidanceLoop(B, D, st, P, config=cfg, seed=3, micro_batches=mb)
        assert len(lp.lanes) == mb and lp.launches_per_step() == mb * (4 * lp.kernels_per_eval() + 2)
        g.manual_seed(7)
        x_t = torch.randn(B, lp.L, generator=g); vel = 0.1 * torch.randn(B, lp.L, generator=g)
        out = lp.denoise_step_host(13, sdf0.cpu().pin_memory(), x_t.pin_memory(), vel.pin_memory(), theta0.cpu().pin_memory())
        res.append({n: t.clone() for n, t in out.items()})
    for n in ("theta", "velocity", "prev_sample", "terms"):
        assert torch.allclose(res[0][n], res[1][n], rtol=1e-4, atol=1e-6), n


def test_overlapped_and_serial_evaluations_agree():
    """desc.serial=1 (every kernel in series on the caller's stream) and the default fork/join layout
    give the same energy terms and gradients; the trace hook reports every kernel of the evaluation."""
    from followmyhold_b200.guidance.engine import GuidanceEngine
    B, D, P = 2, 64, 4096
    samples = [make_guidance_sample(D, P, 80 + i) for i in range(B)]
    sdf, theta, st = stack_samples(samples, cap=True)
    res = []
    for serial in (1, 0):
        eng = GuidanceEngine(B, D, 778, st.hand_faces.shape[0], P)
        eng.prepare(st)
        eng.serial = serial
        desc = eng.make_desc(sdf, theta, st)
        trace = torch.tensor([[2 ** 63 - 1, 0]] * 11, dtype=torch.int64, device="cuda")
        desc.trace = trace.data_ptr()
        for _ in range(2):
            eng.launch(desc)
        torch.cuda.synchronize()
        t = trace.cpu()
        ran = [i for i in range(11) if int(t[i, 1]) > 0]
        assert ran == [0, 1, 2, 3, 5, 6, 7, 8, 9, 10], ran          # everything but the brute-force chamfer
        assert all(int(t[i, 0]) <= int(t[i, 1]) for i in ran)
        res.append((eng.terms.cpu().clone(), eng.grad_theta.cpu().clone(), eng.grad_sdf.cpu().clone()))
    (t0, g0, s0), (t1, g1, s1) = res
    assert torch.allclose(t0, t1, rtol=1e-5, atol=1e-9)
    assert torch.allclose(g0, g1, rtol=2e-5, atol=1e-9)
    assert torch.allclose(s0, s1, rtol=1e-5, atol=1e-12)


def test_autograd_function_chains_into_torch():
    from followmyhold_b200.guidance.engine import GuidanceEngine, GuidanceFunction
    B, D, P = 1, 32, 512
    samples = [make_guidance_sample(D, P, 50)]
    sdf, theta, st = stack_samples(samples, cap=True)
    eng = GuidanceEngine(B, D, 778, st.hand_faces.shape[0], P)
    z = torch.zeros(1, device=sdf.device, requires_grad=True)
    th = theta.clone().requires_grad_(True)
    E = GuidanceFunction.apply(sdf + z.view(1, 1, 1, 1), th, eng, st)
    (2.0 * E.sum()).backward()
    torch.cuda.synchronize()
    assert torch.allclose(z.grad, 2.0 * eng.grad_sdf.sum().view(1), rtol=1e-4)
    assert torch.allclose(th.grad, 2.0 * eng.grad_theta, rtol=1e-6)
