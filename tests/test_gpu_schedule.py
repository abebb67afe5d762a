"""The whole guided denoise schedule (plain steps -> hand-only phase -> object-only phase -> joint
phase, pipelines.py:1262-1612) on the GPU against the same schedule driven on the CPU by the oracle
energy + torch.optim.Adam/AdamW wired like ``get_guidance_params`` (code_utils.py:57-78)."""
import numpy as np
import pytest
import torch

from followmyhold_b200.synthetic import make_guidance_sample, stack_samples
from oracle import guidance_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_phase_weights(phase):
    w = O.Weights()
    if phase == 1:
        w.w_hand, w.w_kp, w.w_treg_h = 1.0, 1e-2, 1e-2
        for n in ("w_pen", "w_con", "w_ivol", "w_mom", "w_treg_o", "w_int_lo", "w_int_hi", "w_dist", "w_vreg", "w_edge"):
            setattr(w, n, 0.0)
    elif phase == 1.5:
        w.w_treg_o, w.w_hand, w.w_ch = 1e-2, 0.0, 0.0
    return w


def _oracle_schedule(sample, sdf0, tap, alpha, cfg, outputs, x_t, theta0, dt=torch.float64):
    """One image.  Returns (theta [16], velocity [L], x_t [L]) after the whole schedule."""
    sig = O.set_timesteps_sigmas(cfg.num_inference_steps).to(dt)
    x_t = x_t.to(dt).clone()
    leaves = [theta0[0:1], theta0[1:4], theta0[4:8], theta0[8:9], theta0[9:12], theta0[12:16]]
    leaves = [l.to(dt).clone() for l in leaves]
    v = None
    for i in range(cfg.num_inference_steps):
        v = outputs[i].to(dt).clone()
        if i < cfg.handopt_start_step:
            phase = 0
        elif i == cfg.handopt_start_step:
            phase = 1
        elif i == cfg.handopt_start_step + 1:
            phase = 1.5
        else:
            phase = 2
        if phase:
            ps = [l.clone().requires_grad_(True) for l in leaves]
            vp = v.clone().requires_grad_(True)
            h1, h2, ob15, ob2 = cfg.phase1_hand_lrs, cfg.phase2_hand_lrs, cfg.obj_2half_lrs, cfg.obj_lrs
            if phase == 1:
                groups = [dict(params=[ps[0]], lr=h1["scale"]), dict(params=[ps[1]], lr=h1["trans"]), dict(params=[ps[2]], lr=h1["rot"])]
                opt = torch.optim.Adam(groups, eps=1e-4)
                n_it = cfg.optimization_steps_hand
            elif phase == 1.5:
                groups = [dict(params=[ps[3]], lr=ob15["scale"]), dict(params=[ps[4]], lr=ob15["trans"]), dict(params=[ps[5]], lr=ob15["rot"]),
                          dict(params=[vp], lr=cfg.noise_obj_lr1)]
                opt = torch.optim.AdamW(groups, eps=1e-4)
                n_it = cfg.optimization_steps_scale
            else:
                groups = [dict(params=[ps[0]], lr=h2["scale"]), dict(params=[ps[1]], lr=h2["trans"]), dict(params=[ps[2]], lr=h2["rot"]),
                          dict(params=[ps[3]], lr=ob2["scale"]), dict(params=[ps[4]], lr=ob2["trans"]), dict(params=[ps[5]], lr=ob2["rot"]),
                          dict(params=[vp], lr=cfg.noise_obj_lr2)]
                opt = torch.optim.AdamW(groups, eps=1e-4)
                n_it = cfg.optimization_steps_joint
            W = _oracle_phase_weights(phase)
            for _ in range(n_it):
                opt.zero_grad()
                x1 = O.scheduler_step_final(x_t, vp, sig[i])
                # the mock decoder (foho_mock_decoder_forward): sdf[tap[j]] = sdf0[tap[j]] + alpha * x1[j]
                flat0 = sdf0.to(dt).reshape(-1)
                sdf = flat0.scatter(0, tap, flat0[tap] + alpha * x1).reshape(sdf0.shape)
                out = O.guidance_energy(sdf, sample.hand_rest.to(dt), sample.hand_faces, sample.cloud.to(dt),
                                        torch.cat(ps[0:3]), torch.cat(ps[3:6]), sample.T_h2m.to(dt), sample.obj_center.to(dt), W,
                                        j_regressor=sample.j_regressor.to(dt), kps_2d=sample.kps_2d.to(dt),
                                        fov_deg=sample.fov_deg, image_hw=sample.image_hw)
                out["total"].backward()
                opt.step()
            leaves = [p.detach().clone() for p in ps]
            v = vp.detach().clone()
        x_t, _ = O.scheduler_step(x_t, v, sig[i], sig[i + 1])
    return torch.cat(leaves), v, x_t


@pytest.mark.parametrize("micro_batches", [1, 2])
def test_full_schedule_matches_oracle_schedule(micro_batches):
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.loop import GuidanceLoop
    B, D, P, L = 2, 32, 512, 1024
    samples = [make_guidance_sample(D, P, 90 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, cap=True)
    cfg = OptimizationConfig().with_steps(6)
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 5, 4, 3
    g = torch.Generator().manual_seed(11)
    x_t = torch.randn(B, L, generator=g)
    outputs = [0.1 * torch.randn(B, L, generator=g) for _ in range(cfg.num_inference_steps)]
    res = {}
    for use_graphs in (False, True):
        lp = GuidanceLoop(B, D, st, P, config=cfg, latent_elems=L, seed=4, micro_batches=micro_batches)
        lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.x_t.copy_(x_t); lp.theta.copy_(theta0)
        assert [lp.phase_of_step(i) for i in range(6)] == [0, 0, 1, 1.5, 2, 2]
        lp.run_schedule_device([o.cuda() for o in outputs], use_graphs=use_graphs)
        torch.cuda.synchronize()
        res[use_graphs] = (lp.theta.cpu().clone(), lp.velocity.cpu().clone(), lp.x_t.cpu().clone(), lp.tap.cpu().clone(), lp.alpha)
    th_e, v_e, x_e, tap, alpha = res[False]
    th_g, v_g, x_g, _, _ = res[True]
    assert torch.allclose(th_g, th_e, rtol=1e-4, atol=1e-6) and torch.allclose(x_g, x_e, rtol=1e-4, atol=1e-6)
    faces = st.hand_faces.cpu()
    for b in range(B):
        s = samples[b]
        s.hand_faces = faces
        th_o, v_o, x_o = _oracle_schedule(s, sdf0[b].cpu(), tap, alpha, cfg, [o[b] for o in outputs], x_t[b], theta0[b].cpu())
        # every phase moved its leaves
        assert not torch.allclose(th_o[:8], theta0[b, :8].cpu().double()) and not torch.allclose(th_o[8:], theta0[b, 8:].cpu().double())
        dth = (th_e[b].double() - th_o).abs().max().item()
        dv = (v_e[b].double() - v_o).abs().max().item()
        dx = (x_e[b].double() - x_o).abs().max().item()
        print(f"image {b}: max|dtheta|={dth:.3e} max|dv|={dv:.3e} max|dx_t|={dx:.3e}")
        # measured on B200: 5e-7 / 5e-8 / 2.4e-7 (fp32 kernels vs the float64 oracle schedule)
        assert torch.allclose(th_e[b].double(), th_o, rtol=1e-4, atol=1e-5), (th_e[b], th_o)
        assert dv <= 2e-6 and dx <= 5e-6


def test_torch_decoder_schedule_matches_the_graph_schedule():
    """``run_schedule_decoder`` (a differentiable PyTorch decoder in the loop, autograd from dE/dSDF back to the
    model output -- how the reference's own VAE is driven, pipelines.py:1507-1508,1600) against
    ``run_schedule_device`` when the torch decoder is the same linear tap map the stand-in kernels implement."""
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.loop import GuidanceLoop
    B, D, P, L = 2, 32, 512, 1024
    samples = [make_guidance_sample(D, P, 90 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, cap=True)
    cfg = OptimizationConfig().with_steps(6)
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 5, 4, 3
    g = torch.Generator().manual_seed(11)
    x_t = torch.randn(B, L, generator=g)
    outputs = [(0.1 * torch.randn(B, L, generator=g)).cuda() for _ in range(cfg.num_inference_steps)]
    res = []
    for torch_decoder in (False, True):
        lp = GuidanceLoop(B, D, st, P, config=cfg, latent_elems=L, seed=4, micro_batches=1)
        lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.x_t.copy_(x_t); lp.theta.copy_(theta0)
        if torch_decoder:
            flat0, tap, alpha = lp.sdf0.reshape(B, -1), lp.tap, lp.alpha

            def decode(x1):                        # foho_mock_decoder_forward in torch ops
                return flat0.scatter(1, tap.view(1, -1).expand(B, -1), flat0[:, tap] + alpha * x1).reshape(B, D, D, D)

            lp.run_schedule_decoder(outputs, decode)
        else:
            lp.run_schedule_device(outputs, use_graphs=True)
        torch.cuda.synchronize()
        res.append((lp.theta.cpu().clone(), lp.velocity.cpu().clone(), lp.x_t.cpu().clone(), lp.nan_report()))
    (th_g, v_g, x_g, n_g), (th_t, v_t, x_t2, n_t) = res
    assert n_g == {} and n_t == {}
    assert not torch.equal(th_t, theta0.cpu())
    assert torch.allclose(th_t, th_g, rtol=1e-4, atol=1e-6)
    assert torch.allclose(v_t, v_g, rtol=1e-4, atol=1e-6) and torch.allclose(x_t2, x_g, rtol=1e-4, atol=1e-6)


def test_half_latents_through_the_loop():
    """``latent_dtype=torch.float16`` (the reference's dtype, pipelines.py:1204).  The pieces are pinned on their own --
    ``foho_guidance_update_f16`` bit for bit against torch's half AdamW, ``foho_scheduler_step_f16`` against the
    reference scheduler's golden vectors -- so what is held here is the wiring: with ONE inner iteration per phase the
    half loop's ``x1`` is the float loop's rounded once and the latent gradient agrees to half precision (same inputs,
    same volume), the whole schedule runs in half end to end (eager and as graph replay), and the stand-in decoder's
    half entry points are exact.  (After many inner iterations the two dtypes legitimately part: torch's half AdamW
    keeps its second moment in half, where g^2 underflows -- which is what the reference runs.)"""
    import ctypes as C
    from followmyhold_b200.guidance.config import OptimizationConfig
    from followmyhold_b200.guidance.loop import GuidanceLoop
    B, D, P, L = 2, 32, 512, 1024
    samples = [make_guidance_sample(D, P, 90 + i) for i in range(B)]
    sdf0, theta0, st = stack_samples(samples, cap=True)
    g = torch.Generator().manual_seed(11)
    x_t = torch.randn(B, L, generator=g).half()
    outputs = [(0.1 * torch.randn(B, L, generator=g)).half() for _ in range(6)]

    def run(dt, graphs, inner):
        cfg = OptimizationConfig().with_steps(6)
        cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = inner
        lp = GuidanceLoop(B, D, st, P, config=cfg, latent_elems=L, seed=4, micro_batches=2, latent_dtype=dt)
        lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0); lp.x_t.copy_(x_t); lp.theta.copy_(theta0)
        # up to and including the object-only step (index 3): its single inner iteration is the first decode
        lp.run_schedule_device([o.cuda().to(dt) for o in outputs], use_graphs=graphs, last_step=3 if inner == (1, 1, 1) else None)
        torch.cuda.synchronize()
        assert all(t.dtype == dt for t in (lp.x_t, lp.velocity, lp.x1, lp.grad_velocity, lp.prev))
        return lp
    a32, a16 = run(torch.float32, False, (1, 1, 1)), run(torch.float16, False, (1, 1, 1))
    x1_32, x1_16 = a32.x1.cpu(), a16.x1.float().cpu()
    assert float((x1_16 - x1_32).abs().max()) <= 2.0 ** -9 * float(x1_32.abs().max())          # rounded once (two ops in half)
    g32, g16 = a32.grad_velocity.cpu(), a16.grad_velocity.float().cpu()
    assert float(g32.abs().max()) > 0
    assert float((g16 - g32).abs().max()) <= 4e-3 * float(g32.abs().max())
    full_e, full_g = run(torch.float16, False, (5, 4, 3)), run(torch.float16, True, (5, 4, 3))
    for lp in (full_e, full_g):
        assert torch.isfinite(lp.x_t.float()).all() and torch.isfinite(lp.velocity.float()).all() and torch.isfinite(lp.theta).all()
    assert not torch.equal(full_e.theta.cpu(), theta0.cpu()) and not torch.equal(full_g.theta.cpu(), theta0.cpu())
    # (no eager-vs-replay comparison of the FULL half schedule: torch's half AdamW divides by sqrt(v) + eps with v underflowing
    # to zero in half, which amplifies the last-bit differences between two runs of the energy's float atomics beyond any
    # useful tolerance; the float schedule's eager-vs-replay agreement is held in test_full_schedule_matches_oracle_schedule)
    cfg = OptimizationConfig().with_steps(6)
    # the stand-in decoder with half latents: sdf0 + alpha * float(x1) at the taps, gradient rounded once
    lp = GuidanceLoop(B, D, st, P, config=cfg, latent_elems=L, seed=4, latent_dtype=torch.float16)
    lp.sdf0.copy_(sdf0); lp.sdf.copy_(sdf0)
    x1 = torch.randn(B, L, device="cuda").half()
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    vol = D ** 3
    assert lp.lib.foho_mock_decoder_forward_f16(lp.sdf.data_ptr(), lp.sdf0.data_ptr(), x1.data_ptr(), lp.tap.data_ptr(), B, vol, L,
                                                lp.alpha, sp) == 0
    want = lp.sdf0.view(B, vol).clone()
    want[:, lp.tap] = want[:, lp.tap] + lp.alpha * x1.float()
    assert torch.allclose(lp.sdf.view(B, vol), want, rtol=1e-6, atol=1e-6)                 # fused multiply-add vs two roundings
    gs = torch.randn(B, vol, device="cuda")
    gv = torch.empty(B, L, device="cuda", dtype=torch.float16)
    assert lp.lib.foho_mock_decoder_backward_f16(gs.data_ptr(), lp.tap.data_ptr(), gv.data_ptr(), B, vol, L, 0.25, sp) == 0
    assert torch.equal(gv, (0.25 * gs[:, lp.tap]).half())                                   # one product, rounded once
