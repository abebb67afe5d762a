"""Guided-denoise loop over a batch of images: the optimisation-in-the-loop part of the reference's
``__call__`` (third_party_patches/hy3dgen/shapegen/pipelines.py:1262-1612: plain steps, the hand-only,
object-only and joint phases) driven through the C-ABI.

One *guided-denoise step* = ``optimization_steps_joint`` (50) guidance evaluations
[decode -> energy fwd+bwd -> decoder adjoint -> fused AdamW + step_final] followed by one
``scheduler.step``.  The whole step is captured once into a CUDA graph and
replayed, so the inner loop has no host round trips at all (the reference syncs several
times per iteration).  ``micro_batches`` > 1 advances groups of images as
independent lanes inside that graph; ``run_schedule_device`` runs the whole schedule;
``denoise_steps_host`` is the pinned-host-buffer API with upload / compute / download pipelined.

The VAE decoder (``latent2sdf``) is not built yet; a fixed sparse
linear decoder stands in for it (``foho_mock_decoder_*``) so latents, velocity and the
optimiser are exercised with real data flow.  A real decoder plugs in through
``GuidanceFunction`` (engine.py) instead.
"""
from __future__ import annotations

import ctypes as C
import time
import types
from typing import Dict, Optional

import torch

from .. import _lib
from .config import OptimizationConfig
from .engine import GuidanceEngine, GuidanceOptimizer, GuidanceStatics

LATENT_SHAPE = (3072, 64)   # Hunyuan3D-2 ShapeVAE latent (pipelines.py:700)


def set_timesteps_sigmas(num_inference_steps: int, shift: float = 1.0) -> torch.Tensor:
    """``FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=linspace(0,1,N))``
    (schedulers.py:171-211; call site pipelines.py:1187): N+1 float32 sigmas, last = 1."""
    s = torch.linspace(0, 1, num_inference_steps, dtype=torch.float64)
    s = shift * s / (1 + (shift - 1) * s)
    return torch.cat([s.to(torch.float32), torch.ones(1)])


def slice_statics(st: GuidanceStatics, off: int, n: int) -> GuidanceStatics:
    """Images [off, off+n) of a batch of per-image inputs (topology, regressor and camera are shared)."""
    cut = lambda t: None if t is None else t.narrow(0, off, n).contiguous()
    return GuidanceStatics(hand_rest=cut(st.hand_rest), hand_faces=st.hand_faces, cloud=cut(st.cloud), T_h2m=cut(st.T_h2m),
                           obj_center=cut(st.obj_center), j_regressor=st.j_regressor, kps_2d=cut(st.kps_2d),
                           fov_deg=st.fov_deg, image_hw=st.image_hw)


class _Lane:
    """One micro-batch of a GuidanceLoop: images [off, off+nb)."""
    def __init__(self, off, nb, engine, opt, statics, stream):
        self.off, self.nb, self.engine, self.opt, self.statics, self.stream = off, nb, engine, opt, statics, stream


class GuidanceLoop:
    """Batched guided-denoise steps for B images on one GPU."""

    def __init__(self, B: int, D: int, statics: GuidanceStatics, P: int, device="cuda:0",
                 config: Optional[OptimizationConfig] = None, weights=None, latent_elems: int = LATENT_SHAPE[0] * LATENT_SHAPE[1],
                 decoder_alpha: float = 0.05, stream_variant: int = 0, seed: int = 0, micro_batches: int = 1,
                 loss_log_every: int = 0, mock_decoder: bool = True, max_obj_verts: int = 0,
                 latent_dtype: torch.dtype = torch.float32):
        """``loss_log_every`` = n > 0 keeps the loss terms of every n-th inner iteration of every step
        (``loss_history``; the reference logs every 10th when ``FOHO_DEBUG_DIR`` is set, pipelines.py:1446-1450,
        1594-1598) -- device-to-device copies inside the step graph, no host sync.

        ``latent_dtype=torch.float16`` carries the latents, the model output being optimised, its gradient and ``x1`` in half,
        as the reference does (pipelines.py:1204; the leaf is a clone of the DiT's half output, code_utils.py:43-78): the
        update is then ``foho_guidance_update_f16`` (torch's half AdamW, rounding after every op), the scheduler
        ``foho_scheduler_step_f16`` (pinned by the reference scheduler's golden vectors) and the decoders read half ``x1``.

        ``micro_batches`` = m > 1 splits the B images into m groups that advance independently inside the
        one captured graph (own engine, optimiser state, stream and library side-stream lane): while one
        group's evaluation is in its serial tail (assemble -> decoder adjoint -> update -> decoder forward)
        the other group's dense stream keeps the HBM busy.  Results are identical: images never interact."""
        self.lib = _lib.load()
        self.device = torch.device(device)
        if latent_dtype not in (torch.float32, torch.float16):
            raise ValueError("latent_dtype must be torch.float32 or torch.float16")
        self.latent_dtype = latent_dtype
        self.B, self.D, self.P, self.L = B, D, P, latent_elems
        self.cfg = config or OptimizationConfig()
        self.statics = statics
        Vh, Fh = statics.hand_rest.shape[1], statics.hand_faces.shape[0]
        m = int(micro_batches)
        if m < 1 or m > 4 or B % m != 0:
            raise ValueError("micro_batches must be 1..4 and divide the batch")
        self.micro_batches, nb = m, B // m
        dev = self.device
        self.terms = torch.zeros(B, _lib.FOHO_NUM_TERMS, dtype=torch.float32, device=dev)
        self.grad_theta = torch.zeros(B, 16, dtype=torch.float32, device=dev)
        self.lanes = []
        for j in range(m):
            st_j = statics if m == 1 else slice_statics(statics, j * nb, nb)
            eng = GuidanceEngine(nb, D, Vh, Fh, P, device=device, weights=weights, stream_variant=stream_variant,
                                 max_obj_verts=max_obj_verts)
            eng.lane = j
            if m > 1:
                # the lanes' stream kernels share the SMs most of the time: 3 bulk loads in flight per CTA
                # (2 x 48 KB per SM when both run) measured best, 826 vs 792 image-steps/s with 4
                # (profiles/r01_microbatch_probe.json); a 5-slot ring is enough for that
                eng.stream_stages, eng.stream_prefetch = 5, 3
            eng.terms = self.terms.narrow(0, j * nb, nb)             # the lanes write straight into the
            eng.grad_theta = self.grad_theta.narrow(0, j * nb, nb)   # loop's [B, .] result buffers
            eng.prepare(st_j)
            self.lanes.append(_Lane(j * nb, nb, eng, GuidanceOptimizer(nb, self.L, device=device, config=self.cfg, velocity_dtype=latent_dtype), st_j,
                                    None if j == 0 else torch.cuda.Stream(device=dev)))
        self.engine = self.lanes[0].engine       # the whole batch when micro_batches == 1
        self.opt = self.lanes[0].opt
        vol = D * D * D
        self.mock_decoder = bool(mock_decoder)
        f32 = torch.float32
        self.alpha = float(decoder_alpha)
        if self.mock_decoder:
            # the linear stand-in decoder: one latent token (64 channels) drives one run of 64 consecutive voxels along z
            if self.L > vol:
                raise ValueError("mock decoder needs latent_elems <= D^3 (pass mock_decoder=False to drive a real decoder)")
            g = torch.Generator().manual_seed(seed)
            run = LATENT_SHAPE[1] if (self.L % LATENT_SHAPE[1] == 0 and vol % LATENT_SHAPE[1] == 0) else 1
            starts = torch.randperm(vol // run, generator=g)[: self.L // run].sort().values * run
            self.tap = (starts.view(-1, 1) + torch.arange(run).view(1, -1)).reshape(-1).to(dev)   # int64 voxel taps
            self.sdf0 = torch.empty(B, D, D, D, dtype=f32, device=dev)     # decoder output for x1 = 0 (per image)
        else:
            self.tap, self.sdf0 = None, None         # run_schedule_decoder / run_schedule_tc_decoder only
        self.sdf = torch.empty(B, D, D, D, dtype=f32, device=dev)      # current decode
        lt = latent_dtype
        self.x_t = torch.zeros(B, self.L, dtype=lt, device=dev)        # latents
        self.velocity = torch.zeros(B, self.L, dtype=lt, device=dev)   # model output being optimised
        self.x1 = torch.zeros(B, self.L, dtype=lt, device=dev)
        self.grad_velocity = torch.zeros(B, self.L, dtype=lt, device=dev)
        self.prev = torch.zeros(B, self.L, dtype=lt, device=dev)
        half = lt == torch.float16
        self._sched_fn = self.lib.foho_scheduler_step_f16 if half else self.lib.foho_scheduler_step
        self._mock_fwd = self.lib.foho_mock_decoder_forward_f16 if half else self.lib.foho_mock_decoder_forward
        self._mock_bwd = self.lib.foho_mock_decoder_backward_f16 if half else self.lib.foho_mock_decoder_backward
        self.theta = torch.zeros(B, 16, dtype=f32, device=dev)
        self.reset_leaves()
        self.sigmas = set_timesteps_sigmas(self.cfg.num_inference_steps)
        # NaN guard (pipelines.py:1442-1444,1590-1592), kept on the device: nan_flag[b] = inner iteration (1-based)
        # at which image b's total became NaN in the current outer step, nan_steps[i] = that flag after step i
        self.nan_flag = torch.zeros(B, dtype=torch.int32, device=dev)
        self.nan_steps = torch.zeros(self.cfg.num_inference_steps, B, dtype=torch.int32, device=dev)
        self.loss_log_every = int(loss_log_every)
        self.loss_history: Optional[torch.Tensor] = None
        if self.loss_log_every > 0:
            kmax = max(self.cfg.optimization_steps_hand, self.cfg.optimization_steps_scale, self.cfg.optimization_steps_joint)
            self.loss_history = torch.full((self.cfg.num_inference_steps, (kmax + self.loss_log_every - 1) // self.loss_log_every,
                                            B, _lib.FOHO_NUM_TERMS), float("nan"), dtype=f32, device=dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._graph_key = None
        self._sched_graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self.stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._pinned: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------ state
    def reset_leaves(self) -> None:
        """Identity similarity leaves (pipelines.py:1208-1215)."""
        self.theta.zero_()
        self.theta[:, 0] = 1.0; self.theta[:, 4] = 1.0
        self.theta[:, 8] = 1.0; self.theta[:, 12] = 1.0

    # ------------------------------------------------------------------ phases
    def phase_of_step(self, step_index: int) -> float:
        """Which optimisation the reference runs at denoise step ``i`` (pipelines.py:1293-1295,1361,1455):
        0 = none (plain scheduler.step), 1 = hand only, 1.5 = object only, 2 = joint."""
        c = self.cfg
        # the reference gates the hand-only and object-only steps on `i >= handopt_start_step` alone
        # (pipelines.py:1293-1361); guidance_end_step bounds the joint phase only (:1455)
        if step_index < c.handopt_start_step:
            return 0
        if step_index == c.handopt_start_step:
            return 1
        if step_index == c.handopt_start_step + 1:
            return 1.5
        if step_index > c.guidance_end_step:
            return 0
        return 2

    def phase_iterations(self, phase: float) -> int:
        c = self.cfg
        return {0: 0, 1: c.optimization_steps_hand, 1.5: c.optimization_steps_scale, 2: c.optimization_steps_joint}[phase]

    def phase_weights(self, phase: float):
        """Loss weights per phase.  REF literals: phase 1 ``total_hand_loss`` = 1e-2 kp + 1e-2 treg_h (+ image
        terms, pipelines.py:1343-1349); phase 1.5 ``total_obj_loss`` = 1e-3 verts + 1e-2 treg_o (+ image/edge
        terms, :1433-1440); phase 2 = the defaults (:1499-1504,1578-1588).  The NS data terms follow the
        leaves being optimised: the chamfer to the observed cloud in the hand phase (it stands where the
        reference's rendered hand-vs-MoGe terms stand), the volume terms in the object phases."""
        base = self.engine.weights
        if phase == 2:
            return base
        w = _lib.Weights()
        C.memmove(C.byref(w), C.byref(base), C.sizeof(_lib.Weights))
        if phase == 1:
            w.w_hand, w.w_kp, w.w_treg_h = 1.0, 1e-2, 1e-2
            for n in ("w_pen", "w_con", "w_ivol", "w_mom", "w_treg_o", "w_int_lo", "w_int_hi", "w_dist", "w_vreg", "w_edge"):
                setattr(w, n, 0.0)
        elif phase == 1.5:
            w.w_treg_o = 1e-2
            w.w_hand = 0.0
            w.w_ch = 0.0
        return w

    def kernels_per_eval(self) -> int:
        return self.engine.launches_per_eval + 3     # + decoder fwd, decoder adjoint, fused update

    # ------------------------------------------------------------------ rendered hand terms (row f2, first part)
    def enable_image_terms(self, targets, hand_faces_render: Optional[torch.Tensor] = None, tile_cap: int = 1024) -> None:
        """Switch on the rendered hand terms of the reference: phase 1 ``1 * normal + 10 * disparity + 1 * silhouette``
        (pipelines.py:1327-1349), phases 2 ``hand_loss`` part ``10 * normal + 10 * disparity`` (:1488-1504, inside the
        ``1e-3 * hand_loss`` of :1588).  ``targets``: ``render.ImageTargets`` for the B images -- MoGe normal map with
        the hand mask as valid mask, ``moge_disp * hand_mask``, the hand silhouette (:1230-1256).  ``hand_faces_render``
        [F,3] int32: the topology to render (default: the statics' faces; pass the uncapped MANO faces when those
        were closed for the sign rule).  The transformed hand is rendered by ``foho_raster_losses_fwd_bwd``; its
        vertex gradient enters the fused evaluation through ``grad_hand_ext``.  The joined hand + object terms
        (:1544-1569) need the extracted object mesh (FlexiCubes, not built) and are not part of this."""
        from .render import ImageLossRenderer, ImageTargets
        Vh = self.statics.hand_rest.shape[1]
        faces = (hand_faces_render if hand_faces_render is not None else self.statics.hand_faces).to(self.device, torch.int32)
        H, W = int(targets.gt_normals.shape[1]), int(targets.gt_normals.shape[2])
        for ln in self.lanes:
            nb = ln.nb
            r = ImageLossRenderer(nb, H, W, nb * Vh, nb * faces.shape[0], device=self.device, tile_cap=tile_cap)
            cut = lambda t: t.narrow(0, ln.off, nb)
            r.set_targets(ImageTargets(gt_normals=cut(targets.gt_normals), gt_mask=cut(targets.gt_mask), gt_disp=cut(targets.gt_disp),
                                       gt_sil=cut(targets.gt_sil), fov_deg=cut(targets.fov_deg)))
            ln.renderer = r
            ln.render_faces = torch.cat([faces + b * Vh for b in range(nb)]).contiguous()
            ln.render_vo = torch.arange(0, (nb + 1) * Vh, Vh, dtype=torch.int32, device=self.device)
            ln.render_fo = torch.arange(0, (nb + 1) * faces.shape[0], faces.shape[0], dtype=torch.int32, device=self.device)
        self.image_terms = torch.zeros(self.B, 8, dtype=torch.float32, device=self.device)

    def _hand_image_grad(self, ln: _Lane, phase: float, s: torch.cuda.Stream, weights=None, grad_out=None, prep: bool = True):
        """Enqueue prep (-> transformed hand) + renderer; returns dE_img/d(transformed hand verts) [nb,Vh,3] or None.
        ``grad_out`` [>= nb*Vh, 3]: accumulate into it instead (the joined render has already written there)."""
        r = getattr(ln, "renderer", None)
        if r is None or phase == 1.5:                        # the object-only step does not move the hand (:1361-1453)
            return None
        w_hand = float((weights or self.phase_weights(phase)).w_hand)
        r.w = (1.0 * w_hand, 10.0 * w_hand, 1.0 * w_hand) if phase == 1 else (10.0 * w_hand, 10.0 * w_hand, 0.0)
        theta = self.theta.narrow(0, ln.off, ln.nb)
        if prep:
            desc = ln.engine.make_desc(self.sdf.narrow(0, ln.off, ln.nb), theta, ln.statics)
            desc.stage_mask = 1                              # prep only: leaves -> transformed hand vertices
            ln.engine.launch(desc, s)
        Vh = ln.statics.hand_rest.shape[1]
        losses, g = r(ln.engine.hand_moge.view(-1, 3), ln.render_faces, ln.render_vo, ln.render_fo, stream=s,
                      accumulate=grad_out is not None, grad_out=grad_out)
        with torch.cuda.stream(s):
            self.image_terms.narrow(0, ln.off, ln.nb).copy_(losses)
        return g[:ln.nb * Vh].view(ln.nb, Vh, 3)

    # ------------------------------------------------------------------ extracted object mesh: REF mesh terms + joined renders
    def enable_object_terms(self, hoi_targets=None, obj_targets=None, cap_verts: int = 0, tile_cap: int = 1024) -> None:
        """Put the extracted object surface into ``run_schedule_tc_decoder`` (one lane): every inner iteration of the
        object-only and joint phases extracts the mesh from the decoded volume (``foho_dmc_extract``, where the
        reference calls FlexiCubes, pipelines.py:1393,1509), feeds it to the explicit-mesh terms a7 / a10
        (``distance_loss``, ``obj_verts_loss``, ``mesh_edge_loss``, :1529-1541,1570-1576) and -- with targets -- to the
        renderer: object alone in the object-only phase (``10 nrm + 10 disp + 100 sil``, :1413-1440), hand + object joined
        in the joint phase (``10 nrm + 10 disp + 10 sil``, :1544-1569,1580-1583); all vertex gradients return through the
        object similarity and T_h2m to the mesh, through the extraction to dE/dSDF, and on through the decoder adjoint.
        Needs ``enable_image_terms`` first when targets are given (the hand renderer supplies the topology)."""
        from .render import ImageLossRenderer
        from .surface import SurfaceExtractor
        if self.micro_batches != 1:
            raise ValueError("object terms drive one lane")
        ln = self.lanes[0]
        B, D, Vh = self.B, self.D, self.statics.hand_rest.shape[1]
        cap_v = int(cap_verts) or B * 6 * D * D
        if ln.engine.max_obj_verts < cap_v:
            raise ValueError(f"construct the loop with max_obj_verts >= {cap_v}")
        V1 = B * Vh
        rf = getattr(ln, "render_faces", None)
        F1 = 0 if rf is None else int(rf.shape[0])
        cap_f = 2 * cap_v + 64
        o = types.SimpleNamespace()
        o.V1, o.F1, o.cap_v, o.cap_f = V1, F1, cap_v, cap_f
        o.joint_verts = torch.zeros(V1 + cap_v, 3, dtype=torch.float32, device=self.device)       # [hand (MoGe) | object (MoGe)]
        o.joint_faces = torch.zeros(F1 + cap_f, 3, dtype=torch.int32, device=self.device)
        if rf is not None:
            o.joint_faces[:F1].copy_(rf)
        o.G = torch.zeros(V1 + cap_v, 3, dtype=torch.float32, device=self.device)                 # dE_img/d(joint vertices)
        o.ex = SurfaceExtractor(B, D, device=self.device, cap_verts=cap_v, cap_faces=cap_f, index_base=V1, faces_out=o.joint_faces[F1:])
        ln.engine.hand_moge = o.joint_verts[:V1].view(B, Vh, 3)          # the engine writes the transformed hand in place
        o.R_hoi = o.R_obj = None
        if hoi_targets is not None or obj_targets is not None:
            if rf is None:
                raise ValueError("enable_image_terms() first")
            H, W = ln.renderer.H, ln.renderer.W
            for name, t in (("R_hoi", hoi_targets), ("R_obj", obj_targets)):
                if t is not None:
                    r = ImageLossRenderer(B, H, W, V1 + cap_v, F1 + cap_f, device=self.device, tile_cap=tile_cap)
                    r.set_targets(t)
                    setattr(o, name, r)
        o.terms_hoi = torch.zeros(B, 8, dtype=torch.float32, device=self.device)
        o.terms_obj = torch.zeros(B, 8, dtype=torch.float32, device=self.device)
        self._obj = o

    def _object_eval(self, ln: _Lane, phase: float, late: bool, w, s: torch.cuda.Stream):
        """One evaluation with the extracted object mesh in it (called after the decode); leaves dE/dSDF complete in
        ``engine.grad_sdf``."""
        from .engine import ObjectMeshBatch
        o, eng = self._obj, ln.engine
        B, Vh = self.B, ln.statics.hand_rest.shape[1]
        ex = o.ex
        ex.extract(self.sdf, stream=s)
        mesh = ObjectMeshBatch(ex.verts, ex.vert_offsets, ex.edges, ex.edge_offsets)
        g_hand = g_obj = None
        rend = o.R_hoi if phase == 2 else o.R_obj
        if rend is not None:
            d0 = eng.make_desc(self.sdf, self.theta, ln.statics, late_step=late, obj_mesh=mesh, obj_moge=o.joint_verts[o.V1:])
            d0.w = w
            d0.stage_mask = 1 | 32                                   # leaves -> transformed hand and object vertices
            eng.launch(d0, s)
            set2 = (o.V1, o.F1, ex.vert_offsets, ex.face_offsets)
            if phase == 2:
                rend.w = (10.0, 10.0, 10.0)                          # :1580-1583
                losses, _ = rend(o.joint_verts, o.joint_faces, ln.render_vo, ln.render_fo, stream=s, set2=set2, grad_out=o.G)
                with torch.cuda.stream(s):
                    o.terms_hoi.copy_(losses)
                self._hand_image_grad(ln, 2, s, w, grad_out=o.G, prep=False)      # + 1e-3 * (10 nrm_h + 10 disp_h), :1499-1504,1588
                g_hand = o.G[:o.V1].view(B, Vh, 3)
            else:
                rend.w = (10.0, 10.0, 100.0)                         # :1433-1440
                losses, _ = rend(o.joint_verts, o.joint_faces, ln.render_vo, ln.render_fo, stream=s, set2=set2, skip_set1=True,
                                 grad_out=o.G)
                with torch.cuda.stream(s):
                    o.terms_obj.copy_(losses)
            g_obj = o.G[o.V1:]
        elif phase == 2:
            g_hand = self._hand_image_grad(ln, 2, s, w)
        desc = eng.make_desc(self.sdf, self.theta, ln.statics, late_step=late, grad_hand_ext=g_hand, obj_mesh=mesh, grad_obj_ext=g_obj)
        desc.w = w
        eng.launch(desc, s)
        ex.backward(eng.grad_obj_verts, eng.grad_sdf, stream=s)      # mesh gradient -> dE/dSDF (accumulated)

    # ------------------------------------------------------------------ one evaluation (enqueue only)
    def _enqueue_eval(self, sigma: float, late_step: bool, s: torch.cuda.Stream, phase: float = 2, lane: Optional[_Lane] = None,
                      log_slot: Optional[torch.Tensor] = None) -> None:
        """One evaluation of one lane (default: lane 0) on stream ``s``.  ``log_slot`` [nb,16]: where to keep
        this evaluation's loss terms (the values the reference prints before ``backward()``)."""
        ln = lane or self.lanes[0]
        lib, vol, off, nb = self.lib, self.D ** 3, ln.off, ln.nb
        sp = C.c_void_p(s.cuda_stream)
        sdf, sdf0, theta = self.sdf.narrow(0, off, nb), self.sdf0.narrow(0, off, nb), self.theta.narrow(0, off, nb)
        x_t, x1 = self.x_t.narrow(0, off, nb), self.x1.narrow(0, off, nb)
        vel, gvel = self.velocity.narrow(0, off, nb), self.grad_velocity.narrow(0, off, nb)
        hand_only = phase == 1
        if not hand_only:
            _lib.check("foho_mock_decoder_forward", self._mock_fwd(
                sdf.data_ptr(), sdf0.data_ptr(), x1.data_ptr(), self.tap.data_ptr(), nb, vol, self.L, self.alpha, sp))
        desc = ln.engine.make_desc(sdf, theta, ln.statics, late_step=late_step, grad_hand_ext=self._hand_image_grad(ln, phase, s))
        if phase != 2:
            self._phase_w = self.phase_weights(phase)      # keep the struct alive while the call reads it
            desc.w = self._phase_w
        if hand_only:
            desc.stage_mask = 1 | 4 | 16                   # no volume term has weight: skip the stream and the voxels
        ln.engine.launch(desc, s)
        if not hand_only:
            _lib.check("foho_mock_decoder_backward", self._mock_bwd(
                ln.engine.grad_sdf.data_ptr(), self.tap.data_ptr(), gvel.data_ptr(), nb, vol, self.L,
                self.alpha * (1.0 - sigma), sp))
        if log_slot is not None:
            with torch.cuda.stream(s):
                log_slot.copy_(ln.engine.terms)
        # NaN guard in every phase.  The reference tests the total in phases 1.5 and 2 only (:1442,1590); a NaN
        # total in its hand phase turns the hand leaves into NaN, the renders of the next phase with them and the
        # image ends at the `return None` of :1444 -- the same outcome as stopping the image here, but the
        # leaves (and with them every coordinate the kernels loop over) stay finite
        ln.opt.step(theta, ln.engine.grad_theta, vel, gvel, x_t, x1, sigma=sigma, stream=s,
                    terms=ln.engine.terms, nan_flag=self.nan_flag.narrow(0, off, nb))

    def _enqueue_step(self, step_index: int, s: torch.cuda.Stream, phase: float = 2) -> None:
        """All kernels of one guided-denoise step (pipelines.py:1293-1612), no syncs.  ``phase``: 1 hand
        only (:1295-1358), 1.5 object only (:1361-1453), 2 joint (:1455-1601), 0 plain ``scheduler.step``.
        With micro-batches every lane runs its images on its own stream, forked from and joined into ``s``."""
        cfg = self.cfg
        sigma = float(self.sigmas[step_index]); sigma_next = float(self.sigmas[step_index + 1])
        late = step_index >= cfg.num_inference_steps - 3
        if phase == 0:
            _lib.check("foho_scheduler_step", self._sched_fn(
                self.x_t.data_ptr(), self.velocity.data_ptr(), self.prev.data_ptr(), None, self.x_t.numel(), sigma,
                sigma_next, C.c_void_p(s.cuda_stream)))
            return
        for ln in self.lanes[1:]:
            ln.stream.wait_stream(s)
        for ln in self.lanes:
            st = s if ln.stream is None else ln.stream
            off, nb = ln.off, ln.nb
            x_t, vel = self.x_t.narrow(0, off, nb), self.velocity.narrow(0, off, nb)
            x1, prev = self.x1.narrow(0, off, nb), self.prev.narrow(0, off, nb)
            with torch.cuda.stream(st):
                ln.opt.set_phase(phase)
                ln.opt.reset()                             # fresh optimiser state every outer step (:1318,1384,1478)
                self.nan_flag.narrow(0, off, nb).zero_()   # the NaN `break` ends one outer step's inner loop only
                # x1 for the first decode of this step: step_final with the incoming velocity (:1507)
                _lib.check("foho_scheduler_step", self._sched_fn(
                    x_t.data_ptr(), vel.data_ptr(), None, x1.data_ptr(), x_t.numel(), sigma, sigma_next,
                    C.c_void_p(st.cuda_stream)))
                for k in range(self.phase_iterations(phase)):
                    slot = None
                    if self.loss_history is not None and k % self.loss_log_every == 0:
                        slot = self.loss_history[step_index, k // self.loss_log_every].narrow(0, off, nb)
                    self._enqueue_eval(sigma, late, st, phase, ln, slot)
                self.nan_steps[step_index].narrow(0, off, nb).copy_(self.nan_flag.narrow(0, off, nb))
                # obj_latents = scheduler.step(noise_pred_obj, t, obj_latents).prev_sample (:1612)
                _lib.check("foho_scheduler_step", self._sched_fn(
                    x_t.data_ptr(), vel.data_ptr(), prev.data_ptr(), None, x_t.numel(), sigma, sigma_next,
                    C.c_void_p(st.cuda_stream)))
        for ln in self.lanes[1:]:
            s.wait_stream(ln.stream)

    def launches_per_step(self) -> int:
        # per lane: 2 scheduler launches + the evaluations (opt.reset(): 4 memsets by torch, not counted)
        return self.micro_batches * (self.cfg.optimization_steps_joint * self.kernels_per_eval() + 2)

    # ------------------------------------------------------------------ NaN guard, loss log (host side, after the run)
    def nan_report(self) -> Dict[int, Dict[int, int]]:
        """{denoise step: {image: inner iteration (0-based, like the reference's ``k``) whose total was NaN}}.
        Synchronises; call it after the schedule."""
        ns = self.nan_steps.cpu()
        rep: Dict[int, Dict[int, int]] = {}
        for i, b in ns.nonzero().tolist():
            rep.setdefault(i, {})[b] = int(ns[i, b]) - 1
        return rep

    def failed_images(self) -> list:
        """Images whose hand-only or object-only phase hit a NaN total: the reference's ``__call__`` returns
        ``None`` for them (pipelines.py:1442-1444; a NaN hand phase gets there one step later) and the stage
        reports the image as failed (run.py:257-259).  A NaN in the joint phase only ends that step's inner
        loop (:1590-1592)."""
        bad = set()
        for i, imgs in self.nan_report().items():
            if self.phase_of_step(i) in (1, 1.5):
                bad.update(imgs)
        return sorted(bad)

    def loss_log_lines(self, image: int) -> list:
        """The reference's ``losses.txt`` lines for one image ("Opt step k, ..." every ``loss_log_every``
        iterations of every optimised step), with this path's term names."""
        if self.loss_history is None:
            return []
        h = self.loss_history[:, :, image].cpu()
        lines = []
        for i in range(h.shape[0]):
            phase = self.phase_of_step(i)
            if phase == 0:
                continue
            for j in range((self.phase_iterations(phase) + self.loss_log_every - 1) // self.loss_log_every):
                t = h[i, j]
                if bool(torch.isnan(t).all()):         # slot never written: that step has not been run
                    continue
                body = ", ".join(f"{n}: {float(t[q])}" for q, n in enumerate(_lib.TERM_NAMES) if n and q < t.numel())
                lines.append(f"Denoise step {i} phase {phase}, Opt step {j * self.loss_log_every}, {body}")
        return lines

    # ------------------------------------------------------------------ graph
    def capture(self, step_index: int) -> None:
        """Capture one guided-denoise step for ``step_index`` (sigma is baked into the graph)."""
        if self._graph is not None and self._graph_key == step_index:
            return
        with torch.cuda.device(self.device):
            s = self.stream
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for ln in self.lanes:          # warm-up outside capture (function attributes, side streams)
                    self._enqueue_eval(float(self.sigmas[step_index]), False, s, 2, ln)
                    ln.opt.reset()
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                self._enqueue_step(step_index, s)
            self._graph, self._graph_key = g, step_index
            torch.cuda.current_stream(self.device).wait_stream(s)

    def run_schedule_device(self, model_output, first_step: int = 0, last_step: Optional[int] = None,
                            use_graphs: bool = True) -> None:
        """The reference's whole guided denoise loop for the B resident images (pipelines.py:1262-1612):
        for every step i, ``velocity <- model_output(i, x_t)`` (the DiT + CFG prediction, out of this
        path's scope: a callable or a sequence of [B,L] device tensors), the phase's optimisation (hand
        only at ``handopt_start_step``, object only at the next step, joint afterwards), then
        ``x_t <- scheduler.step(velocity, x_t)``.  Leaves carry over from step to step (:1604-1610)."""
        cfg = self.cfg
        last = cfg.num_inference_steps - 1 if last_step is None else last_step
        with torch.cuda.device(self.device):
            s = self.stream
            s.wait_stream(torch.cuda.current_stream(self.device))
            cur = torch.cuda.current_stream(self.device)
            for i in range(first_step, last + 1):
                # the prediction runs on the ambient stream and reads x_t, which step i-1's graph and copy (on
                # `s`) wrote: order the two streams both ways, and keep v's memory alive for `s`
                cur.wait_stream(s)
                v = model_output(i, self.x_t) if callable(model_output) else model_output[i]
                s.wait_stream(cur)
                if v.is_cuda:
                    v.record_stream(s)
                with torch.cuda.stream(s):
                    self.velocity.copy_(v)
                    phase = self.phase_of_step(i)
                    if use_graphs and phase != 0:
                        key = (i, phase)
                        g = self._sched_graphs.get(key)
                        if g is None:
                            # one evaluation outside the capture (function attributes, lazily created side
                            # streams); it moves the leaves, so put them back before the real run
                            theta_keep = self.theta.clone()
                            for ln in self.lanes:
                                self._enqueue_eval(float(self.sigmas[i]), False, s, phase, ln)
                            s.synchronize()
                            self.theta.copy_(theta_keep); self.velocity.copy_(v)
                            g = torch.cuda.CUDAGraph()
                            with torch.cuda.graph(g, stream=s):
                                self._enqueue_step(i, s, phase)
                            self._sched_graphs[key] = g
                        g.replay()
                    else:
                        self._enqueue_step(i, s, phase)
                    self.x_t.copy_(self.prev)
            torch.cuda.current_stream(self.device).wait_stream(s)

    def run_schedule_decoder(self, model_output, decode, first_step: int = 0, last_step: Optional[int] = None) -> None:
        """``run_schedule_device`` with a differentiable PyTorch decoder in the loop instead of the linear
        stand-in -- how the reference's own networks are driven: ``decode(x1 [B, L]) -> sdf [B, D, D, D]``
        float32, negative inside (``latent2sdf``, pipelines.py:292-312), built from torch ops so that autograd
        carries dE/dSDF back to the model output (:1507-1508, 1600).  Per inner iteration: ``step_final``
        (torch, differentiable), ``decode``, the fused energy kernels through ``GuidanceFunction``,
        ``backward()`` through the decoder, the fused AdamW update with the NaN guard.  Eager (the decoder
        is not capturable in general), one lane; everything else -- phases, weights, leaf groups, optimiser
        reset per outer step, ``scheduler.step`` -- as in ``run_schedule_device``
        (tests/test_gpu_schedule.py::test_torch_decoder_schedule_matches_the_graph_schedule: equal to the graph
        path when the torch decoder is the linear stand-in).  ``self.sdf`` must hold a finite volume for the
        hand-only phase: its volume terms have zero weight there but are still evaluated."""
        from .engine import GuidanceFunction
        if self.micro_batches != 1:
            raise ValueError("run_schedule_decoder drives one lane: construct the loop with micro_batches=1")
        cfg = self.cfg
        last = cfg.num_inference_steps - 1 if last_step is None else last_step
        ln = self.lanes[0]
        eng, opt = ln.engine, ln.opt
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream(self.device)
            sp = C.c_void_p(s.cuda_stream)
            for i in range(first_step, last + 1):
                v = model_output(i, self.x_t) if callable(model_output) else model_output[i]
                self.velocity.copy_(v)
                phase = self.phase_of_step(i)
                sigma, sigma_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
                late = i >= cfg.num_inference_steps - 3
                if phase != 0:
                    opt.set_phase(phase)
                    opt.reset()
                    self.nan_flag.zero_()
                    w = self.phase_weights(phase)
                    for k in range(self.phase_iterations(phase)):
                        if phase == 1:
                            # hand only: no volume term has weight, nothing to decode (:1295-1358)
                            desc = eng.make_desc(self.sdf, self.theta, ln.statics, late_step=late)
                            desc.w = w
                            desc.stage_mask = 1 | 4 | 16
                            eng.launch(desc, s)
                            gvel = None
                        else:
                            vel = self.velocity.detach().clone().requires_grad_(True)
                            x1 = self.x_t + (1.0 - sigma) * vel                        # step_final (schedulers.py:481)
                            sdf = decode(x1)
                            E = GuidanceFunction.apply(sdf, self.theta, eng, ln.statics, w, late, 0)
                            E.sum().backward()
                            gvel = vel.grad.contiguous()
                        if self.loss_history is not None and k % self.loss_log_every == 0:
                            self.loss_history[i, k // self.loss_log_every].copy_(eng.terms)
                        opt.step(self.theta, eng.grad_theta, None if gvel is None else self.velocity, gvel,
                                 self.x_t, self.x1, sigma=sigma, stream=s, terms=eng.terms, nan_flag=self.nan_flag)
                    self.nan_steps[i].copy_(self.nan_flag)
                _lib.check("foho_scheduler_step", self._sched_fn(
                    self.x_t.data_ptr(), self.velocity.data_ptr(), self.prev.data_ptr(), None, self.x_t.numel(), sigma,
                    sigma_next, sp))
                self.x_t.copy_(self.prev)

    def run_schedule_tc_decoder(self, model_output, decoder, first_step: int = 0, last_step: Optional[int] = None,
                                grad_cap: int = 8192, keep_mom: bool = False) -> None:
        """The guided denoise loop with the reference's OWN decoder in it -- ``latent2sdf`` (pipelines.py:292-312)
        and its adjoint on the tensor cores (``decoder``: a ``followmyhold_b200.decoder.shapevae.LatentDecoder``
        whose lattice is this loop's D^3 grid) -- instead of torch autograd through a torch module
        (``run_schedule_decoder``) or the linear stand-in (``run_schedule_device``).  Per inner iteration
        (:1507-1601): ``step_final`` -> decode -> fused energy kernels -> sparse view of dE/dSDF (the energy
        touches a few thousand voxels) -> decoder adjoint -> dE/dv = (1 - sigma) dE/dx1 -> fused AdamW.
        No torch arithmetic, no host sync inside the loop.

        The builder-added dense volume term ``L_mom`` is switched off here unless ``keep_mom``: it would make
        every interior voxel carry a gradient; its reference twin (``obj_verts_loss``, :1570) lives on the
        extracted mesh.  ``grad_cap``: rows per image the adjoint handles; an overflow sets
        ``self.grad_compactor.flags`` (checked by ``check_overflow()``)."""
        from ..decoder.shapevae import GradCompactor
        if self.micro_batches != 1:
            raise ValueError("run_schedule_tc_decoder drives one lane: construct the loop with micro_batches=1")
        if decoder.B != self.B or decoder.Nq != self.D ** 3 or self.L != LATENT_SHAPE[0] * LATENT_SHAPE[1]:
            raise ValueError("decoder batch / lattice / latent shape do not match this loop")
        cfg = self.cfg
        last = cfg.num_inference_steps - 1 if last_step is None else last_step
        ln = self.lanes[0]
        eng, opt = ln.engine, ln.opt
        B, V = self.B, self.D ** 3
        if getattr(self, "grad_compactor", None) is None or self.grad_compactor.cap != grad_cap:
            self.grad_compactor = GradCompactor(B, V, grad_cap, self.device)
        comp = self.grad_compactor
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream(self.device)
            sp = C.c_void_p(s.cuda_stream)
            for i in range(first_step, last + 1):
                v = model_output(i, self.x_t) if callable(model_output) else model_output[i]
                self.velocity.copy_(v)
                phase = self.phase_of_step(i)
                sigma, sigma_next = float(self.sigmas[i]), float(self.sigmas[i + 1])
                late = i >= cfg.num_inference_steps - 3
                if phase != 0:
                    opt.set_phase(phase)
                    opt.reset()
                    self.nan_flag.zero_()
                    w = _lib.Weights()
                    C.memmove(C.byref(w), C.byref(self.phase_weights(phase)), C.sizeof(_lib.Weights))
                    if not keep_mom:
                        w.w_mom = 0.0
                    for k in range(self.phase_iterations(phase)):
                        g_img = None if (phase != 1 and getattr(self, "_obj", None) is not None) else self._hand_image_grad(ln, phase, s, w)
                        if phase == 1:
                            desc = eng.make_desc(self.sdf, self.theta, ln.statics, late_step=late, grad_hand_ext=g_img)
                            desc.w = w
                            desc.stage_mask = 1 | 4 | 16
                            eng.launch(desc, s)
                            vel = gvel = None
                        else:
                            _lib.check("foho_scheduler_step", self._sched_fn(
                                self.x_t.data_ptr(), self.velocity.data_ptr(), None, self.x1.data_ptr(), self.x_t.numel(),
                                sigma, sigma_next, sp))                                   # step_final (:1507)
                            decoder.forward(self.x1.view(B, LATENT_SHAPE[0], LATENT_SHAPE[1]), out=self.sdf.view(B, V), stream=s)
                            if getattr(self, "_obj", None) is not None:
                                if phase == 1.5:
                                    w.w_dist = 0.0                       # no hand -> object distance term in the object-only step (:1433-1440)
                                self._object_eval(ln, phase, late, w, s)
                            else:
                                desc = eng.make_desc(self.sdf, self.theta, ln.statics, late_step=late, grad_hand_ext=g_img)
                                desc.w = w
                                eng.launch(desc, s)
                            idx, val = comp(eng.grad_sdf.view(B, V), stream=s)
                            decoder.backward(idx, val, out=self.grad_velocity.view(B, LATENT_SHAPE[0], LATENT_SHAPE[1]), stream=s,
                                             out_scale=1.0 - sigma)
                            vel, gvel = self.velocity, self.grad_velocity
                        if self.loss_history is not None and k % self.loss_log_every == 0:
                            self.loss_history[i, k // self.loss_log_every].copy_(eng.terms)
                        opt.step(self.theta, eng.grad_theta, vel, gvel, self.x_t, self.x1, sigma=sigma, stream=s,
                                 terms=eng.terms, nan_flag=self.nan_flag)
                    self.nan_steps[i].copy_(self.nan_flag)
                _lib.check("foho_scheduler_step", self._sched_fn(
                    self.x_t.data_ptr(), self.velocity.data_ptr(), self.prev.data_ptr(), None, self.x_t.numel(), sigma,
                    sigma_next, sp))
                self.x_t.copy_(self.prev)

    def check_flags(self) -> None:
        """Raise if any evaluation of any lane overflowed its candidate list (see GuidanceEngine.check_flags)."""
        for ln in self.lanes:
            ln.engine.check_flags()

    def check_overflow(self) -> None:
        """Raise when the sparse gradient view of ``run_schedule_tc_decoder`` overflowed its capacity (synchronises)."""
        comp = getattr(self, "grad_compactor", None)
        if comp is not None and int(comp.flags.item()) & 1:
            raise _lib.FohoStatusError("foho_dec_compact_grad", -2, f"more than grad_cap={comp.cap} voxels of one image carry a "
                                       f"gradient (counts {comp.count.tolist()}); raise grad_cap")

    def run_step_device(self, step_index: int) -> None:
        """Replay one guided-denoise step; inputs (sdf0, x_t, velocity, theta) already in HBM."""
        self.capture(step_index)
        self._graph.replay()

    # ------------------------------------------------------------------ host-buffer API (end to end)
    def _pin(self, name: str, like: torch.Tensor) -> torch.Tensor:
        t = self._pinned.get(name)
        if t is None or t.shape != like.shape or t.dtype != like.dtype:
            t = torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            self._pinned[name] = t
        return t

    def denoise_step_host(self, step_index: int, sdf0_host: torch.Tensor, x_t_host: torch.Tensor,
                          velocity_host: torch.Tensor, theta_host: torch.Tensor, out: Optional[dict] = None) -> dict:
        """One guided-denoise step with HOST inputs/outputs (pinned CPU tensors).

        H2D: decoder base volume [B,D,D,D], latents [B,L], model output [B,L], leaves [B,16].
        D2H: optimised model output [B,L], prev_sample [B,L], leaves [B,16], loss terms [B,16].
        Returns a dict of pinned host tensors (valid after the call returns)."""
        self.capture(step_index)
        dev = self.device
        with torch.cuda.device(dev):
            cs = self.copy_stream
            with torch.cuda.stream(cs):
                self.sdf0.copy_(sdf0_host, non_blocking=True)
                self.sdf.copy_(self.sdf0, non_blocking=True)
                self.x_t.copy_(x_t_host, non_blocking=True)
                self.velocity.copy_(velocity_host, non_blocking=True)
                self.theta.copy_(theta_host, non_blocking=True)
            self.stream.wait_stream(cs)
            with torch.cuda.stream(self.stream):
                self._graph.replay()
                if out is None:
                    out = {
                        "velocity": self._pin("o_velocity", self.velocity), "prev_sample": self._pin("o_prev", self.prev),
                        "theta": self._pin("o_theta", self.theta), "terms": self._pin("o_terms", self.terms),
                    }
                out["velocity"].copy_(self.velocity, non_blocking=True)
                out["prev_sample"].copy_(self.prev, non_blocking=True)
                out["theta"].copy_(self.theta, non_blocking=True)
                out["terms"].copy_(self.terms, non_blocking=True)
            self.stream.synchronize()
        return out

    # ------------------------------------------------------------------ host-buffer API, pipelined
    def denoise_steps_host(self, step_index: int, batches) -> list:
        """Guided-denoise steps for a sequence of image batches whose inputs live in pinned HOST memory
        (``batches``: iterable of ``(sdf0, x_t, velocity, theta)`` CPU tensors, each a different batch of
        B images -- the way a rank works through its share of ``sorted(images)[rank::world]``; ``sdf0``
        may be ``None`` when the batch's decoder base volume is already on the device, or float16 -- the
        decoder's own output dtype -- in which case it is widened on the device).

        Three streams: the upload of batch k+1 into a staging set overlaps the graph replay of batch k,
        whose results leave through a second staging set while batch k+1 computes.  Every byte still
        crosses PCIe inside the call (H2D ``h2d_bytes_per_step`` and D2H ``d2h_bytes_per_step`` per
        batch).  Returns one dict of pinned host tensors per batch, valid when the call returns."""
        self.capture(step_index)
        dev = self.device
        outs = []
        t_begin = time.perf_counter()
        with torch.cuda.device(dev):
            if not hasattr(self, "_stage"):
                # two staging sets: the upload of batch k+1 starts the moment batch k's upload ends, so the
                # PCIe link -- the slowest resource of this path -- never waits for the compute stream
                self._stage = [{n: torch.empty_like(t) for n, t in (("sdf0", self.sdf0), ("x_t", self.x_t),
                                                                     ("velocity", self.velocity), ("theta", self.theta))}
                               for _ in range(2)]
                self._ostage = {n: torch.empty_like(t) for n, t in (("velocity", self.velocity), ("prev_sample", self.prev),
                                                                     ("theta", self.theta), ("terms", self.terms))}
                self.d2h_stream = torch.cuda.Stream(device=dev)
                self._ev_stage_free = [torch.cuda.Event(), torch.cuda.Event()]
                self._ev_out_free = torch.cuda.Event()
            cs, s, ds = self.copy_stream, self.stream, self.d2h_stream
            ost = self._ostage
            cur = torch.cuda.current_stream(dev)
            for x in (cs, s, ds):
                x.wait_stream(cur)
            for ev in self._ev_stage_free:
                ev.record(s)
            self._ev_out_free.record(ds)
            batches = list(batches)
            # every batch's pinned result buffers exist before the pipeline starts (cached across calls):
            # cudaHostAlloc synchronises the device and would stall the three streams mid-flight
            outs = [{n: self._pin(f"p{k}_{n}", t) for n, t in ost.items()} for k in range(len(batches))]
            for k, (sdf0_h, x_t_h, vel_h, theta_h) in enumerate(batches):
                out = outs[k]
                st = self._stage[k & 1]
                # upload into this batch's staging set as soon as batch k-2 has left it
                cs.wait_event(self._ev_stage_free[k & 1])
                with torch.cuda.stream(cs):
                    if sdf0_h is not None:          # None: the decoder state of this batch is already resident
                        if sdf0_h.dtype == torch.float16:
                            # the decoder's native output (fp16 logits, widened by `.float()` at pipelines.py:309):
                            # half the bytes on the wire, widened on the device
                            if "sdf0_h16" not in st:
                                st["sdf0_h16"] = torch.empty(self.sdf0.shape, dtype=torch.float16, device=dev)
                            st["sdf0_h16"].copy_(sdf0_h, non_blocking=True)
                        else:
                            st["sdf0"].copy_(sdf0_h, non_blocking=True)
                    st["x_t"].copy_(x_t_h, non_blocking=True)
                    st["velocity"].copy_(vel_h, non_blocking=True)
                    st["theta"].copy_(theta_h, non_blocking=True)
                    h2d_done = torch.cuda.Event()
                    h2d_done.record(cs)
                s.wait_event(h2d_done)
                with torch.cuda.stream(s):
                    if sdf0_h is not None:
                        self.sdf0.copy_(st["sdf0_h16"] if sdf0_h.dtype == torch.float16 else st["sdf0"])
                    # the stand-in decoder rewrites the tapped voxels of `sdf` from x1 at every evaluation and nothing writes
                    # the others: `sdf` only has to be re-based when the base volume changed since the last re-base (torch's
                    # version counters see every in-place torch op on either tensor; the kernels touch taps only)
                    if getattr(self, "_sdf_based", None) != (self.sdf0._version, self.sdf._version):
                        self.sdf.copy_(self.sdf0)
                        self._sdf_based = (self.sdf0._version, self.sdf._version)
                    self.x_t.copy_(st["x_t"]); self.velocity.copy_(st["velocity"]); self.theta.copy_(st["theta"])
                    self._ev_stage_free[k & 1].record(s)
                    self._graph.replay()
                    s.wait_event(self._ev_out_free)
                    ost["velocity"].copy_(self.velocity); ost["prev_sample"].copy_(self.prev)
                    ost["theta"].copy_(self.theta); ost["terms"].copy_(self.terms)
                    out_ready = torch.cuda.Event()
                    out_ready.record(s)
                ds.wait_event(out_ready)
                with torch.cuda.stream(ds):
                    for n in ost:
                        out[n].copy_(ost[n], non_blocking=True)
                    self._ev_out_free.record(ds)
            self._last_enqueue_s = time.perf_counter() - t_begin     # host time to enqueue the whole pipeline
            for x in (cs, s, ds):
                x.synchronize()
        return outs

    def h2d_bytes_per_step(self) -> int:
        return 4 * (self.B * self.D ** 3 + 2 * self.B * self.L + self.B * 16)

    def d2h_bytes_per_step(self) -> int:
        return 4 * (2 * self.B * self.L + self.B * 16 + self.B * _lib.FOHO_NUM_TERMS)
