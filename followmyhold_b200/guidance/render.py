"""Host side of the renderer + image-space losses of the guidance loop (row f2, first part): packs the meshes of a
batch, owns the workspace and calls ``foho_raster_losses_fwd_bwd`` -- what
``render_normal_and_disparity`` + ``normal_alignment_loss`` + the disparity L1 + the silhouette BCE and their
``backward()`` are in the reference (third_party_patches/hy3dgen/shapegen/pipelines.py:272-289,178-187,1567-1569,
1580-1583; renderer set-up src/foho/guidance/run.py:84-116).  No torch arithmetic; no CPU fallback."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from .. import _lib


@dataclass
class ImageTargets:
    """Per-batch targets, rendered once from MoGe's mesh in the reference (pipelines.py:1247-1256) and the 2-D masks
    (:1230-1237).  All on the device."""
    gt_normals: torch.Tensor      # [B,H,W,3] float32
    gt_mask: torch.Tensor         # [B,H,W] uint8 / bool: valid mask of the normal loss
    gt_disp: torch.Tensor         # [B,H,W] float32
    gt_sil: torch.Tensor          # [B,H,W] float32
    fov_deg: torch.Tensor         # [B] float32 -- per image: MoGe estimates fov_x per frame


class ImageLossRenderer:
    def __init__(self, B: int, H: int, W: int, max_verts: int, max_faces: int, device="cuda:0", w_normal: float = 10.0,
                 w_disp: float = 10.0, w_sil: float = 10.0, tile_cap: int = 1024):
        self.lib = _lib.load()
        self.B, self.H, self.W = B, H, W
        self.device = torch.device(device)
        self.max_verts, self.max_faces, self.tile_cap = max_verts, max_faces, tile_cap
        self.w = (w_normal, w_disp, w_sil)
        d = _lib.RasterDesc()
        d.B, d.V_total, d.F_total, d.H, d.W, d.tile_cap = B, max_verts, max_faces, H, W, tile_cap
        nbytes = self.lib.foho_raster_workspace_bytes(C.byref(d))
        if nbytes == 0:
            raise ValueError("invalid renderer shape")
        self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        self._ws_ptr = (self.ws.data_ptr() + 255) & ~255
        self._ws_bytes = nbytes
        self.losses = torch.zeros(B, 8, dtype=torch.float32, device=self.device)
        self.grad_verts = torch.zeros(max_verts, 3, dtype=torch.float32, device=self.device)
        self.targets: Optional[ImageTargets] = None
        self._n_valid: Optional[torch.Tensor] = None

    def set_targets(self, t: ImageTargets) -> None:
        B, H, W = self.B, self.H, self.W
        f32 = torch.float32
        self.targets = ImageTargets(
            gt_normals=t.gt_normals.to(self.device, f32).reshape(B, H, W, 3).contiguous(),
            gt_mask=t.gt_mask.to(self.device).to(torch.uint8).reshape(B, H, W).contiguous(),
            gt_disp=t.gt_disp.to(self.device, f32).reshape(B, H, W).contiguous(),
            gt_sil=t.gt_sil.to(self.device, f32).reshape(B, H, W).contiguous(),
            fov_deg=t.fov_deg.to(self.device, f32).reshape(B).contiguous())
        self._n_valid = self.targets.gt_mask.view(B, -1).sum(1, dtype=torch.int32).contiguous()

    def __call__(self, verts: torch.Tensor, faces: torch.Tensor, vert_offsets: torch.Tensor, face_offsets: torch.Tensor,
                 backward: bool = True, debug: bool = False, stream: Optional[torch.cuda.Stream] = None, set2=None,
                 skip_set1: bool = False, accumulate: bool = False, grad_out: Optional[torch.Tensor] = None):
        """``verts`` [Vt,3] float32 packed world-space vertices, ``faces`` [Ft,3] int32 (packed indices),
        ``*_offsets`` [B+1] int32.  Returns (losses [B,8], grad_verts [Vt,3] or None[, debug dict]).

        ``set2 = (V1, F1, vert_offsets2, face_offsets2)``: the arrays hold a second per-image set behind the first V1
        vertices / F1 faces whose counts live on the device (the extracted object mesh); Vt / Ft are then capacities.
        ``skip_set1`` draws set 2 alone; ``accumulate`` adds into ``grad_out`` (default: the renderer's own buffer)."""
        if self.targets is None:
            raise RuntimeError("set_targets() first")
        Vt, Ft = int(verts.shape[0]), int(faces.shape[0])
        if Vt > self.max_verts or Ft > self.max_faces:
            raise ValueError(f"mesh batch has {Vt} vertices / {Ft} faces; renderer was sized for {self.max_verts} / {self.max_faces}")
        for t, dt in ((verts, torch.float32), (faces, torch.int32), (vert_offsets, torch.int32), (face_offsets, torch.int32)):
            if t.dtype != dt or not t.is_cuda or not t.is_contiguous():
                raise ValueError("verts float32, faces / offsets int32, contiguous CUDA tensors expected")
        T = self.targets
        d = _lib.RasterDesc()
        d.B, d.V_total, d.F_total, d.H, d.W, d.tile_cap = self.B, Vt, Ft, self.H, self.W, self.tile_cap
        d.w_normal, d.w_disp, d.w_sil = self.w
        d.verts, d.faces = verts.data_ptr(), faces.data_ptr()
        d.vert_offsets, d.face_offsets = vert_offsets.data_ptr(), face_offsets.data_ptr()
        d.fov_deg, d.gt_normals, d.gt_mask = T.fov_deg.data_ptr(), T.gt_normals.data_ptr(), T.gt_mask.data_ptr()
        d.n_valid, d.gt_disp, d.gt_sil = self._n_valid.data_ptr(), T.gt_disp.data_ptr(), T.gt_sil.data_ptr()
        d.losses = self.losses.data_ptr()
        gbuf = self.grad_verts if grad_out is None else grad_out
        if backward and (gbuf.dtype != torch.float32 or not gbuf.is_contiguous() or gbuf.shape[0] < Vt):
            raise ValueError("grad_out must be a contiguous float32 [>= Vt, 3] tensor")
        d.grad_verts = gbuf.data_ptr() if backward else None
        d.accumulate_grad, d.skip_set1 = int(accumulate), int(skip_set1)
        if set2 is not None:
            d.V1, d.F1 = int(set2[0]), int(set2[1])
            d.vert_offsets2, d.face_offsets2 = set2[2].data_ptr(), set2[3].data_ptr()
        dbg = None
        if debug:
            dbg = {"p2f": torch.empty(self.B, self.H, self.W, dtype=torch.int32, device=self.device),
                   "zbuf": torch.empty(self.B, self.H, self.W, dtype=torch.float32, device=self.device),
                   "nraw": torch.empty(self.B, self.H, self.W, 3, dtype=torch.float32, device=self.device)}
            d.out_p2f, d.out_zbuf, d.out_nraw = dbg["p2f"].data_ptr(), dbg["zbuf"].data_ptr(), dbg["nraw"].data_ptr()
        d.workspace, d.workspace_bytes = self._ws_ptr, self._ws_bytes
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check("foho_raster_losses_fwd_bwd", self.lib.foho_raster_losses_fwd_bwd(C.byref(d), C.c_void_p(s.cuda_stream)))
        out = (self.losses, gbuf[:Vt] if backward else None)
        return out + (dbg,) if debug else out


def pack_meshes(meshes: Sequence, device="cuda:0"):
    """[(verts [V,3], faces [F,3]), ...] -> packed (verts, faces, vert_offsets, face_offsets) on the device."""
    vs, fs, vo, fo = [], [], [0], [0]
    for v, f in meshes:
        v = torch.as_tensor(v, dtype=torch.float32)
        f = torch.as_tensor(f, dtype=torch.int64)
        fs.append(f + vo[-1])
        vs.append(v)
        vo.append(vo[-1] + v.shape[0]); fo.append(fo[-1] + f.shape[0])
    dev = torch.device(device)
    return (torch.cat(vs).to(dev).contiguous(), torch.cat(fs).to(torch.int32).to(dev).contiguous(),
            torch.tensor(vo, dtype=torch.int32, device=dev), torch.tensor(fo, dtype=torch.int32, device=dev))


def render_maps(verts, faces, fov_deg: float, H: int, W: int, device="cuda:0"):
    """``render_normal_and_disparity(renderer, mesh)`` (pipelines.py:272-289) for ONE mesh, as maps: the min/max-normalised
    normal colour [H,W,3] (background zeroed), the min/max-normalised disparity [H,W] and the coverage mask [H,W].
    Set-up code (the reference renders MoGe's mesh once per image, :1247-1256): the rasterisation is the kernel's, the
    two normalisations are a few torch reductions."""
    v, f, vo, fo = pack_meshes([(verts, faces)], device=device)
    r = ImageLossRenderer(1, H, W, v.shape[0], f.shape[0], device=device, tile_cap=max(1024, min(int(f.shape[0]), 8192)))
    z = torch.zeros(1, H, W, device=device)
    r.set_targets(ImageTargets(gt_normals=torch.zeros(1, H, W, 3), gt_mask=torch.zeros(1, H, W, dtype=torch.uint8), gt_disp=z, gt_sil=z,
                               fov_deg=torch.tensor([float(fov_deg)])))
    losses, _, dbg = r(v, f, vo, fo, backward=False, debug=True)
    if float(losses[0, 7]) != 0:
        raise _lib.FohoStatusError("foho_raster_losses_fwd_bwd", _lib.FOHO_E_WORKSPACE, "tile list overflow while rendering the targets")
    hit = dbg["p2f"][0] >= 0
    n = dbg["nraw"][0]
    rn = (n - n.min()) / (n.max() - n.min() + 1e-6) * hit[..., None]
    zb = torch.where(hit, dbg["zbuf"][0], torch.full_like(dbg["zbuf"][0], 10.0))
    d = 1.0 / (zb + 1e-6)
    rd = (d - d.min()) / (d.max() - d.min() + 1e-6)
    return rn, rd, hit


def targets_from_moge(moge_verts, moge_faces, fov_deg: float, hand_mask, obj_mask, device="cuda:0"):
    """The three target sets of the reference for one image (pipelines.py:1229-1256 and the loss call sites):
    ``moge_normal = render(moge_mesh) * hoi_mask``, ``moge_disp`` likewise; hand phase: valid = hand mask, disparity target
    ``moge_disp * hand_mask``, silhouette = hand mask (:1341-1342); object phase: the same with the object mask
    (:1421-1423); joint phase: hoi mask, ``moge_disp``, hoi silhouette (:1567-1569).  Returns a dict of tuples
    (gt_normals, gt_mask, gt_disp, gt_sil) keyed 'hand', 'obj', 'hoi' (device tensors)."""
    hm = torch.as_tensor(hand_mask).to(device) > 0
    om = torch.as_tensor(obj_mask).to(device) > 0
    H, W = hm.shape
    rn, rd, _ = render_maps(moge_verts, moge_faces, fov_deg, H, W, device=device)
    hoi = hm | om
    n = rn * hoi[..., None]
    d = rd * hoi
    return {"hand": (n, hm, d * hm, hm.float()), "obj": (n, om, d * om, om.float()), "hoi": (n, hoi, d, hoi.float())}


def stack_targets(per_image, key: str, fovs) -> ImageTargets:
    t = [p[key] for p in per_image]
    return ImageTargets(gt_normals=torch.stack([x[0] for x in t]), gt_mask=torch.stack([x[1] for x in t]),
                        gt_disp=torch.stack([x[2] for x in t]), gt_sil=torch.stack([x[3] for x in t]),
                        fov_deg=torch.tensor([float(f) for f in fovs]))
