"""Host side of the fused guidance evaluation: torch tensors in, C-ABI call, torch tensors out.

PyTorch is plumbing here (device memory, streams); all arithmetic runs in
``libfoho_b200.so``.  The seam this replaces is the body of the phase-2 inner iteration of
the reference's ``Hunyuan3DDiTFlowMatchingPipeline_main.__call__``
(third_party_patches/hy3dgen/shapegen/pipelines.py:1480-1601): leaves -> energy ->
``backward()`` -> ``joint_optimizer.step()``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence

import torch

from .. import _lib
from .config import OptimizationConfig

GRID_BOUND = 1.10  # pipelines.py:1127


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, shape: Sequence[int], dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


@dataclass
class GuidanceStatics:
    """Per-image inputs that stay fixed over a whole guided denoise run (batched)."""
    hand_rest: torch.Tensor      # [B,Vh,3] f32  aligned MANO verts in MoGe space (pipelines.py:1241)
    hand_faces: torch.Tensor     # [Fh,3]  i32
    cloud: Optional[torch.Tensor]  # [B,P,3] f32 MoGe cloud
    T_h2m: torch.Tensor          # [B,4,4] f32 (alignment/h2m.py output, loaded at pipelines.py:1240)
    obj_center: torch.Tensor     # [B,3]   f32
    j_regressor: Optional[torch.Tensor] = None   # [16,Vh]
    kps_2d: Optional[torch.Tensor] = None        # [B,21,2]
    fov_deg: float = 41.0
    image_hw: Sequence[int] = (512, 512)


@dataclass
class ObjectMeshBatch:
    """Explicit object surfaces of a batch (the FlexiCubes output of pipelines.py:1509, Hunyuan space),
    packed: sample b owns verts[vert_offsets[b]:vert_offsets[b+1]] and edges[edge_offsets[b]:edge_offsets[b+1]]
    (edge indices address the packed vertex array).  Vertex counts may differ per sample and per call."""
    verts: torch.Tensor          # [Vo,3] f32
    vert_offsets: torch.Tensor   # [B+1] i32
    edges: torch.Tensor          # [Eo,2] i32 unique undirected edges (pytorch3d ``edges_packed`` semantics)
    edge_offsets: torch.Tensor   # [B+1] i32


def unique_edges(faces) -> "torch.Tensor":
    """Unique undirected edges of a triangle list [F,3] -> [E,2] int64 (what ``mesh_edge_loss`` of
    pipelines.py:1575 averages over: pytorch3d ``Meshes.edges_packed``)."""
    import numpy as np
    f = faces.detach().cpu().numpy() if torch.is_tensor(faces) else np.asarray(faces)
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0)
    return torch.from_numpy(np.unique(np.sort(e, axis=1), axis=0).astype(np.int64))


def pack_object_meshes(meshes, device="cuda:0") -> ObjectMeshBatch:
    """[(verts [V,3], faces [F,3]), ...] (one per sample, torch or numpy) -> ObjectMeshBatch on ``device``."""
    import numpy as np
    vs, es, vo, eo = [], [], [0], [0]
    for v, f in meshes:
        v = torch.as_tensor(v, dtype=torch.float32).reshape(-1, 3)
        e = unique_edges(f) if len(f) else torch.zeros(0, 2, dtype=torch.int64)
        es.append(e + vo[-1]); vs.append(v)
        vo.append(vo[-1] + v.shape[0]); eo.append(eo[-1] + e.shape[0])
    dev = torch.device(device)
    verts = torch.cat(vs).to(dev).contiguous() if vo[-1] else torch.zeros(0, 3, device=dev)
    edges = torch.cat(es).to(torch.int32).to(dev).contiguous() if eo[-1] else torch.zeros(0, 2, dtype=torch.int32, device=dev)
    return ObjectMeshBatch(verts, torch.tensor(vo, dtype=torch.int32, device=dev), edges,
                           torch.tensor(eo, dtype=torch.int32, device=dev))


def delaunay_neighbours(hand_rest: torch.Tensor):
    """Delaunay neighbour graph of each sample's rest hand vertices [B,Vh,3] -> (offsets int32 [B,Vh+1],
    neighbours uint16-valued int16 storage [B,stride]) on the CPU, or None when Qhull cannot triangulate
    one of them (degenerate input) -- the caller then keeps the box search.  Greedy descent over this
    graph reaches the exact nearest vertex of any query (a property of Delaunay triangulations)."""
    import numpy as np
    try:
        from scipy.spatial import Delaunay
        from scipy.spatial import QhullError
    except Exception:       # pragma: no cover
        return None
    v = hand_rest.detach().cpu().numpy().astype(np.float64)
    B, Vh = v.shape[0], v.shape[1]
    if Vh < 5 or Vh > 65535:
        return None
    offs, nbrs = [], []
    for b in range(B):
        try:
            indptr, indices = Delaunay(v[b]).vertex_neighbor_vertices
        except (QhullError, ValueError):
            return None
        if len(indptr) != Vh + 1 or np.any(np.diff(indptr) == 0):     # a vertex Qhull dropped (duplicate / coplanar)
            return None
        offs.append(np.asarray(indptr, dtype=np.int32)); nbrs.append(np.asarray(indices, dtype=np.uint16))
    stride = (max(len(n) for n in nbrs) + 7) // 8 * 8          # 16-byte rows for the kernel's vector copies
    pad = np.zeros((B, stride), dtype=np.uint16)
    for b, n in enumerate(nbrs):
        pad[b, :len(n)] = n
    return torch.from_numpy(np.stack(offs)), torch.from_numpy(pad.view(np.int16))


class GuidanceEngine:
    """Owns workspace + output buffers for a fixed problem shape and launches the kernels."""

    def __init__(self, B: int, D: int, Vh: int, Fh: int, P: int, device="cuda:0",
                 weights: Optional[_lib.Weights] = None, stream_variant: int = 0, max_obj_verts: int = 0):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.FohoLibraryError("GuidanceEngine needs a CUDA device; there is no CPU fallback")
        self.B, self.D, self.Vh, self.Fh, self.P = B, D, Vh, Fh, P
        self.weights = weights if weights is not None else _lib.default_weights()
        self.stream_variant = stream_variant
        self.max_obj_verts = int(max_obj_verts)
        nbytes = self.lib.foho_guidance_workspace_bytes(B, D, Vh, Fh, P, self.max_obj_verts)
        if nbytes == 0:
            raise ValueError("unsupported guidance problem shape")
        dev = self.device
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
        off = (-self.workspace.data_ptr()) % 256
        self._ws_ptr = self.workspace.data_ptr() + off
        self._ws_bytes = nbytes
        self.grad_sdf = torch.empty(B, D, D, D, dtype=torch.float32, device=dev)
        self.grad_theta = torch.zeros(B, 16, dtype=torch.float32, device=dev)
        self.terms = torch.zeros(B, _lib.FOHO_NUM_TERMS, dtype=torch.float32, device=dev)
        self.hand_moge = torch.zeros(B, Vh, 3, dtype=torch.float32, device=dev)
        self.hand_grid = torch.zeros(B, Vh, 3, dtype=torch.float32, device=dev)
        self.grad_obj_verts = (torch.zeros(self.max_obj_verts, 3, dtype=torch.float32, device=dev)
                               if self.max_obj_verts > 0 else None)
        self.launches_per_eval = 8 if P > 0 else 7      # refreshed by make_desc; + 4 when an object mesh is passed
        self.serial = 0            # 1: every kernel in series on the caller's stream (no internal fork/join)
        self.stream_stages = 0     # TMA stream ring depth / prefetch distance (0 = library default)
        self.stream_prefetch = 0
        self.stream_ctas = 0       # TMA stream persistent CTAs per SM (0 = library default)
        self.lane = 0              # which set of library side streams the evaluations fork onto (0..3)
        self._accel = None
        self._accel_ptr = 0
        self._accel_bytes = 0
        self._accel_for = None
        self._nbr_off = None
        self._nbr = None

    def prepare(self, st: GuidanceStatics, stream: Optional[torch.cuda.Stream] = None, delaunay: bool = True) -> None:
        """Per-image setup, once per set of statics: build the chamfer search structures
        (``foho_guidance_prepare_statics``) and, on the host, the Delaunay neighbour graph of each rest
        hand (``delaunay=True``; scipy/Qhull, ~20 ms per image) for the greedy-walk cloud->hand search.
        Evaluations with the same ``st`` object then use the structured searches; other statics fall
        back to the brute-force search."""
        if self.P <= 0 or st.cloud is None or self.Vh > 1024:
            return
        self._nbr_off = self._nbr = None
        if delaunay:
            graph = delaunay_neighbours(st.hand_rest)
            if graph is not None:
                off, nbr = graph
                self._nbr_off, self._nbr = off.to(self.device), nbr.to(self.device)
        _chk(st.hand_rest, (self.B, self.Vh, 3), torch.float32, "hand_rest")
        _chk(st.cloud, (self.B, self.P, 3), torch.float32, "cloud")
        nbytes = self.lib.foho_guidance_accel_bytes(self.B, self.Vh, self.P)
        if self._accel is None:
            self._accel = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            self._accel_ptr = self._accel.data_ptr() + ((-self._accel.data_ptr()) % 256)
            self._accel_bytes = nbytes
        d = _lib.GuidanceDesc()
        d.B, d.Vh, d.P, d.Fh = self.B, self.Vh, self.P, self.Fh
        d.hand_rest, d.cloud = st.hand_rest.data_ptr(), st.cloud.data_ptr()
        d.hand_faces = _chk(st.hand_faces, (self.Fh, 3), torch.int32, "hand_faces").data_ptr()
        d.accel, d.accel_bytes = self._accel_ptr, self._accel_bytes
        with torch.cuda.device(self.device):
            s = stream if stream is not None else torch.cuda.current_stream(self.device)
            _lib.check("foho_guidance_prepare_statics",
                       self.lib.foho_guidance_prepare_statics(C.byref(d), C.c_void_p(s.cuda_stream)))
        self._accel_for = st

    def make_desc(self, sdf: torch.Tensor, theta: torch.Tensor, st: GuidanceStatics,
                  grad_sdf: Optional[torch.Tensor] = None, late_step: bool = False,
                  grad_hand_ext: Optional[torch.Tensor] = None,
                  obj_mesh: Optional[ObjectMeshBatch] = None, obj_moge: Optional[torch.Tensor] = None,
                  grad_obj_ext: Optional[torch.Tensor] = None) -> _lib.GuidanceDesc:
        B, D, Vh, Fh, P = self.B, self.D, self.Vh, self.Fh, self.P
        f32 = torch.float32
        _chk(sdf, (B, D, D, D), f32, "sdf")
        _chk(theta, (B, 16), f32, "theta")
        _chk(st.hand_rest, (B, Vh, 3), f32, "hand_rest")
        _chk(st.hand_faces, (Fh, 3), torch.int32, "hand_faces")
        _chk(st.T_h2m, (B, 4, 4), f32, "T_h2m")
        _chk(st.obj_center, (B, 3), f32, "obj_center")
        if P > 0:
            if st.cloud is None:
                raise ValueError("cloud is required when P > 0")
            _chk(st.cloud, (B, P, 3), f32, "cloud")
        use_kp = st.j_regressor is not None and st.kps_2d is not None
        if use_kp:
            _chk(st.j_regressor, (16, Vh), f32, "j_regressor")
            _chk(st.kps_2d, (B, 21, 2), f32, "kps_2d")
        g = self.grad_sdf if grad_sdf is None else _chk(grad_sdf, (B, D, D, D), f32, "grad_sdf")
        if grad_hand_ext is not None:
            _chk(grad_hand_ext, (B, Vh, 3), f32, "grad_hand_ext")
        d = _lib.GuidanceDesc()
        d.B, d.D, d.Vh, d.Fh, d.P = B, D, Vh, Fh, P
        d.n_joints = 16 if use_kp else 0
        d.image_h, d.image_w = int(st.image_hw[0]), int(st.image_hw[1])
        d.late_step = int(late_step)
        d.stream_variant = self.stream_variant
        d.serial, d.stream_stages, d.stream_prefetch, d.stream_ctas = self.serial, self.stream_stages, self.stream_prefetch, self.stream_ctas
        d.lane = self.lane
        d.fov_deg = float(st.fov_deg)
        d.bound = GRID_BOUND
        d.w = self.weights
        d.sdf, d.grad_sdf = sdf.data_ptr(), g.data_ptr()
        d.hand_rest, d.hand_faces = st.hand_rest.data_ptr(), st.hand_faces.data_ptr()
        d.cloud = _ptr(st.cloud) if P > 0 else None
        d.T_h2m, d.obj_center, d.theta = st.T_h2m.data_ptr(), st.obj_center.data_ptr(), theta.data_ptr()
        d.j_regressor = _ptr(st.j_regressor) if use_kp else None
        d.kps_2d = _ptr(st.kps_2d) if use_kp else None
        d.grad_hand_ext = _ptr(grad_hand_ext)
        d.grad_theta, d.terms = self.grad_theta.data_ptr(), self.terms.data_ptr()
        d.hand_moge, d.hand_grid = self.hand_moge.data_ptr(), self.hand_grid.data_ptr()
        d.Vo_total = 0
        d.Eo_total = 0
        if obj_mesh is not None and obj_mesh.verts.shape[0] > 0:
            Vo, Eo = int(obj_mesh.verts.shape[0]), int(obj_mesh.edges.shape[0])
            if Vo > self.max_obj_verts:
                raise ValueError(f"object mesh has {Vo} vertices; engine was sized for max_obj_verts={self.max_obj_verts}")
            _chk(obj_mesh.verts, (Vo, 3), f32, "obj_mesh.verts")
            _chk(obj_mesh.vert_offsets, (B + 1,), torch.int32, "obj_mesh.vert_offsets")
            _chk(obj_mesh.edges, (Eo, 2), torch.int32, "obj_mesh.edges")
            _chk(obj_mesh.edge_offsets, (B + 1,), torch.int32, "obj_mesh.edge_offsets")
            d.Vo_total, d.Eo_total = Vo, Eo
            d.obj_verts, d.obj_vert_offsets = obj_mesh.verts.data_ptr(), obj_mesh.vert_offsets.data_ptr()
            d.obj_edges = obj_mesh.edges.data_ptr() if Eo > 0 else None
            d.obj_edge_offsets = obj_mesh.edge_offsets.data_ptr()
            d.grad_obj_verts = self.grad_obj_verts.data_ptr()
            if obj_moge is not None:
                _chk(obj_moge, (Vo, 3), f32, "obj_moge")
                d.obj_moge = obj_moge.data_ptr()
            if grad_obj_ext is not None:
                _chk(grad_obj_ext, (Vo, 3), f32, "grad_obj_ext")
                d.grad_obj_ext = grad_obj_ext.data_ptr()
        d.workspace, d.workspace_bytes = self._ws_ptr, self._ws_bytes
        if getattr(self, "sticky_flags", None) is None or self.sticky_flags.shape[0] != B:
            self.sticky_flags = torch.zeros(B, dtype=torch.int32, device=self.device)
        d.sticky_flags = self.sticky_flags.data_ptr()
        accel = self._accel_for is st and P > 0
        if accel:
            d.accel, d.accel_bytes = self._accel_ptr, self._accel_bytes
            if self._nbr is not None:
                d.hand_nbr_off, d.hand_nbr, d.nbr_stride = self._nbr_off.data_ptr(), self._nbr.data_ptr(), int(self._nbr.shape[1])
        # prep, stream, raster, compact, voxdist, vertex_early, finalize_verts, assemble (+ key-points, + chamfer 1 or 2)
        self.launches_per_eval = 8 + (1 if use_kp and Vh > 744 else 0) + ((2 if accel else 1) if P > 0 else 0)
        return d

    def check_flags(self) -> None:
        """Hard error for what the kernels can only flag (they never synchronise): a candidate-list overflow in ANY
        evaluation since the engine was built (bit0: more voxels inside both hand and object than the workspace
        holds -- the penetration term and count were truncated).  Synchronises; call it after a run."""
        sf = getattr(self, "sticky_flags", None)
        if sf is None:
            return
        bad = [(b, int(f)) for b, f in enumerate(sf.cpu().tolist()) if int(f) & 1]
        if bad:
            raise _lib.FohoStatusError("foho_guidance_energy_fwd_bwd", _lib.FOHO_E_WORKSPACE,
                                       f"candidate list overflow for images {[b for b, _ in bad]}: the hand/object "
                                       "intersection has more voxels than the workspace capacity")

    def launch(self, desc: _lib.GuidanceDesc, stream: Optional[torch.cuda.Stream] = None) -> None:
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check("foho_guidance_energy_fwd_bwd",
                   self.lib.foho_guidance_energy_fwd_bwd(C.byref(desc), C.c_void_p(s.cuda_stream)))

    def energy_fwd_bwd(self, sdf, theta, st: GuidanceStatics, **kw):
        """One batched evaluation.  Returns (terms [B,16], grad_sdf [B,D,D,D], grad_theta [B,16]);
        the returned tensors are the engine's own buffers (overwritten by the next call)."""
        with torch.cuda.device(self.device):
            self.launch(self.make_desc(sdf, theta, st, **kw))
        return self.terms, self.grad_sdf, self.grad_theta

    def terms_dict(self) -> Dict[str, torch.Tensor]:
        t = self.terms.detach().cpu()
        return {n: t[:, i] for i, n in enumerate(_lib.TERM_NAMES)}


# --------------------------------------------------------------------------- optimiser
# bit g of theta_mask <-> leaf group g, in the order of code_utils.py:69-77
GROUPS = ("scale_hand", "trans_hand", "rot_hand", "scale_obj", "trans_obj", "rot_obj")
MASK_HAND = 0b000111
MASK_OBJ = 0b111000
MASK_ALL = 0b111111


class GuidanceOptimizer:
    """AdamW state + fused update for the 16 scalar leaves and the velocity tensor.

    Mirrors ``get_guidance_params`` (third_party/utilz/code_utils.py:3-83): state is
    created fresh for every outer denoise step, learning rates per group come from
    ``OptimizationConfig`` (src/foho/configs/guid_config.py:20-27).

    ``velocity_dtype=torch.float16`` is the reference's configuration: the velocity leaf is a clone of the
    DiT's half output (code_utils.py:43-78), so torch keeps its moments in half and rounds after every op;
    velocity, its gradient, ``x_t`` and ``x1`` are then half tensors (``foho_guidance_update_f16``).  The 16
    scalar leaves are float32 either way (pipelines.py:1208-1215).
    """

    def __init__(self, B: int, L: int, device="cuda:0", config: Optional[OptimizationConfig] = None,
                 velocity_dtype: torch.dtype = torch.float32):
        if velocity_dtype not in (torch.float32, torch.float16):
            raise ValueError("velocity_dtype must be torch.float32 or torch.float16")
        self.lib = _lib.load()
        self.B, self.L = B, L
        self.device = torch.device(device)
        self.config = config or OptimizationConfig()
        self.velocity_dtype = velocity_dtype
        dev = self.device
        self.theta_m = torch.zeros(B, 16, device=dev)
        self.theta_v = torch.zeros(B, 16, device=dev)
        self.vel_m = torch.zeros(B, L, device=dev, dtype=velocity_dtype) if L > 0 else None
        self.vel_v = torch.zeros(B, L, device=dev, dtype=velocity_dtype) if L > 0 else None
        self.step_count = 0
        self.set_phase(2)

    def set_phase(self, phase: float) -> None:
        c = self.config
        if phase == 1:
            h = c.phase1_hand_lrs
            self.lr_theta = [h["scale"], h["trans"], h["rot"], 0.0, 0.0, 0.0]
            self.mask, self.lr_velocity, self.weight_decay, self.opt_velocity = MASK_HAND, 0.0, 0.0, False
        elif phase == 1.5:
            o = c.obj_2half_lrs
            self.lr_theta = [0.0, 0.0, 0.0, o["scale"], o["trans"], o["rot"]]
            self.mask, self.lr_velocity, self.weight_decay, self.opt_velocity = MASK_OBJ, c.noise_obj_lr1, 0.01, True
        elif phase == 2:
            h, o = c.phase2_hand_lrs, c.obj_lrs
            self.lr_theta = [h["scale"], h["trans"], h["rot"], o["scale"], o["trans"], o["rot"]]
            self.mask, self.lr_velocity, self.weight_decay, self.opt_velocity = MASK_ALL, c.noise_obj_lr2, 0.01, True
        else:
            raise ValueError(f"Unknown phase {phase}. Expected 'hand_only (1)', 'obj-only (1.5)' or 'joint_hand_obj (2)'.")
        self.phase = phase

    def reset(self) -> None:
        """Fresh optimiser state (the reference re-creates AdamW every outer step, pipelines.py:1478)."""
        self.theta_m.zero_(); self.theta_v.zero_()
        if self.vel_m is not None:
            self.vel_m.zero_(); self.vel_v.zero_()
        self.step_count = 0

    def step(self, theta, grad_theta, velocity=None, grad_velocity=None, x_t=None, x1=None, sigma: float = 0.0,
             stream: Optional[torch.cuda.Stream] = None, terms: Optional[torch.Tensor] = None,
             nan_flag: Optional[torch.Tensor] = None) -> None:
        """One fused optimiser step (+ ``step_final`` into ``x1``).  ``terms`` [B,16] float32 (the evaluation
        just done) and ``nan_flag`` [B] int32 switch the reference's NaN guard on (pipelines.py:1442-1444,
        1590-1592): once a sample's total is NaN it is left untouched and ``nan_flag`` holds the optimiser
        step at which that happened -- decided on the device, no host sync."""
        if (terms is None) != (nan_flag is None):
            raise ValueError("GuidanceOptimizer.step: the NaN guard needs both `terms` and `nan_flag`")
        if nan_flag is not None and (nan_flag.dtype != torch.int32 or terms.dtype != torch.float32
                                     or not terms.is_contiguous() or terms.shape[-1] != _lib.FOHO_NUM_TERMS):
            raise ValueError("GuidanceOptimizer.step: terms must be contiguous float32 [B,16], nan_flag int32 [B]")
        self.step_count += 1
        d = _lib.UpdateDesc()
        d.B, d.L, d.step = self.B, self.L, self.step_count
        d.beta1, d.beta2, d.eps, d.weight_decay = 0.9, 0.999, 1e-4, self.weight_decay
        for i, v in enumerate(self.lr_theta):
            d.lr_theta[i] = v
        d.lr_velocity, d.sigma, d.theta_mask = self.lr_velocity, float(sigma), self.mask
        d.theta, d.grad_theta = theta.data_ptr(), grad_theta.data_ptr()
        d.theta_m, d.theta_v = self.theta_m.data_ptr(), self.theta_v.data_ptr()
        if velocity is not None and self.opt_velocity:
            for name, t in (("velocity", velocity), ("grad_velocity", grad_velocity), ("x_t", x_t), ("x1", x1)):
                if t is not None and t.dtype != self.velocity_dtype:
                    raise ValueError(f"GuidanceOptimizer.step: {name} is {t.dtype}, optimiser state is {self.velocity_dtype}")
            d.velocity, d.grad_velocity = velocity.data_ptr(), grad_velocity.data_ptr()
            d.vel_m, d.vel_v = self.vel_m.data_ptr(), self.vel_v.data_ptr()
            d.x_t, d.x1 = _ptr(x_t), _ptr(x1)
        d.terms, d.nan_flag = _ptr(terms), _ptr(nan_flag)
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        if self.velocity_dtype == torch.float16:
            _lib.check("foho_guidance_update_f16",
                       self.lib.foho_guidance_update_f16(C.byref(d), C.c_void_p(s.cuda_stream)))
        else:
            _lib.check("foho_guidance_update", self.lib.foho_guidance_update(C.byref(d), C.c_void_p(s.cuda_stream)))


def scheduler_step(x_t: torch.Tensor, velocity: torch.Tensor, sigma: float, sigma_next: float):
    """``FlowMatchEulerDiscreteScheduler.step`` arithmetic on the device
    (schedulers.py:294-309): returns (prev_sample, pred_x1)."""
    lib = _lib.load()
    if not (x_t.is_cuda and velocity.is_cuda):
        raise _lib.FohoLibraryError("scheduler_step needs CUDA tensors; there is no CPU fallback")
    if x_t.dtype != velocity.dtype or x_t.dtype not in (torch.float32, torch.float16):
        raise ValueError("scheduler_step: x_t and velocity must both be float32 or both float16")
    x_t, velocity = x_t.contiguous(), velocity.contiguous()
    prev = torch.empty_like(x_t)
    x1 = torch.empty_like(x_t)
    fn = lib.foho_scheduler_step if x_t.dtype == torch.float32 else lib.foho_scheduler_step_f16
    with torch.cuda.device(x_t.device):
        s = torch.cuda.current_stream(x_t.device)
        _lib.check("foho_scheduler_step", fn(
            x_t.data_ptr(), velocity.data_ptr(), prev.data_ptr(), x1.data_ptr(), x_t.numel(), float(sigma),
            float(sigma_next), C.c_void_p(s.cuda_stream)))
    return prev, x1


class GuidanceFunction(torch.autograd.Function):
    """Differentiable wrapper so a PyTorch decoder can sit upstream of the kernel:
    ``E = GuidanceFunction.apply(sdf, theta, engine, statics[, weights, late_step, stage_mask])`` returns the
    per-sample total energy [B]; backward scales the kernel's dE/dSDF and dE/dtheta.  ``weights`` (a
    ``_lib.Weights``) replaces the engine's loss weights for this call (the per-phase weights of the
    schedule), ``late_step`` is the reference's ``i >= num_inference_steps - 3`` switch (pipelines.py:1561)."""

    @staticmethod
    def forward(ctx, sdf, theta, engine: GuidanceEngine, statics: GuidanceStatics, weights=None,
                late_step: bool = False, stage_mask: int = 0):
        with torch.cuda.device(engine.device):
            desc = engine.make_desc(sdf.contiguous(), theta.contiguous(), statics, late_step=late_step)
            if weights is not None:
                desc.w = weights
            if stage_mask:
                desc.stage_mask = stage_mask
            engine.launch(desc)
        ctx.save_for_backward(engine.grad_sdf.clone(), engine.grad_theta.clone())
        return engine.terms[:, 0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        gs, gt = ctx.saved_tensors
        return gs * grad_out.view(-1, 1, 1, 1), gt * grad_out.view(-1, 1), None, None, None, None, None
