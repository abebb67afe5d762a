"""Host side of the surface extraction (row f2, second part): owns the buffers of ``foho_dmc_extract`` /
``foho_dmc_backward`` -- where the reference calls ``FlexiCubes`` (third_party_patches/hy3dgen/shapegen/
pipelines.py:1393,1509,1642).  Dual Marching Cubes as defined in oracle/surface_oracle.py; NOT FlexiCubes."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib


class SurfaceExtractor:
    def __init__(self, B: int, D: int, device="cuda:0", cap_verts: int = 0, cap_faces: int = 0, cap_edges: int = 0,
                 bound: float = 1.10, index_base: int = 0, with_edges: bool = True, faces_out: Optional[torch.Tensor] = None):
        self.lib = _lib.load()
        self.B, self.D, self.bound, self.index_base = B, D, float(bound), int(index_base)
        dev = self.device = torch.device(device)
        # a closed surface through a D^3 lattice has O(D^2) cubes; 6 D^2 per image leaves room for folded shapes
        self.cap_verts = int(cap_verts) or B * 6 * D * D
        self.cap_faces = int(cap_faces) or 2 * self.cap_verts + 64
        self.cap_edges = int(cap_edges) or 3 * self.cap_verts + 64
        i32 = dict(dtype=torch.int32, device=dev)
        self.verts = torch.zeros(self.cap_verts, 3, dtype=torch.float32, device=dev)
        # ``faces_out``: write the triangles straight into the tail of a joint face array (hand faces in front)
        self.faces = torch.zeros(self.cap_faces, 3, **i32) if faces_out is None else faces_out
        if self.faces.shape != (self.cap_faces, 3) or self.faces.dtype != torch.int32 or not self.faces.is_contiguous():
            raise ValueError("faces_out must be a contiguous int32 [cap_faces, 3] tensor")
        self.edges = torch.zeros(self.cap_edges, 2, **i32) if with_edges else None
        self.vert_offsets = torch.zeros(B + 1, **i32)
        self.face_offsets = torch.zeros(B + 1, **i32)
        self.edge_offsets = torch.zeros(B + 1, **i32)
        self.cube_of_vert = torch.zeros(self.cap_verts, **i32)
        self.flags = torch.zeros(1, **i32)
        n = self.lib.foho_dmc_workspace_bytes(B, D)
        self.ws = torch.empty(n + 256, dtype=torch.uint8, device=dev)
        self._ws_ptr, self._ws_bytes = (self.ws.data_ptr() + 255) & ~255, n
        self._sdf: Optional[torch.Tensor] = None

    def _desc(self, sdf: torch.Tensor) -> _lib.DmcDesc:
        d = _lib.DmcDesc()
        d.B, d.D, d.bound = self.B, self.D, self.bound
        d.cap_verts, d.cap_faces, d.cap_edges, d.index_base = self.cap_verts, self.cap_faces, self.cap_edges, self.index_base
        d.sdf, d.verts, d.faces = sdf.data_ptr(), self.verts.data_ptr(), self.faces.data_ptr()
        d.edges = None if self.edges is None else self.edges.data_ptr()
        d.vert_offsets, d.face_offsets, d.edge_offsets = self.vert_offsets.data_ptr(), self.face_offsets.data_ptr(), self.edge_offsets.data_ptr()
        d.cube_of_vert, d.flags = self.cube_of_vert.data_ptr(), self.flags.data_ptr()
        d.workspace, d.workspace_bytes = self._ws_ptr, self._ws_bytes
        return d

    def extract(self, sdf: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """``sdf`` [B,D,D,D] float32 (negative inside).  Results stay in ``verts / faces / edges / *_offsets`` on the device."""
        if sdf.dtype != torch.float32 or not sdf.is_contiguous() or sdf.numel() != self.B * self.D ** 3:
            raise ValueError("expected a contiguous float32 [B,D,D,D] volume")
        self._sdf = sdf
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check("foho_dmc_extract", self.lib.foho_dmc_extract(C.byref(self._desc(sdf)), C.c_void_p(s.cuda_stream)))

    def backward(self, grad_verts: torch.Tensor, grad_sdf: torch.Tensor, stream: Optional[torch.cuda.Stream] = None) -> None:
        """``grad_sdf`` [B,D,D,D] += dE/dSDF from dE/d(verts) [cap_verts,3] (rows beyond the vertex count are ignored)."""
        if self._sdf is None:
            raise RuntimeError("extract() first")
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        _lib.check("foho_dmc_backward", self.lib.foho_dmc_backward(C.byref(self._desc(self._sdf)), grad_verts.data_ptr(), grad_sdf.data_ptr(),
                                                                   C.c_void_p(s.cuda_stream)))

    def meshes(self):
        """Host copies [(verts [V,3], faces [F,3] (without index_base), edges [E,2]), ...] per image (synchronises)."""
        vo, fo, eo = self.vert_offsets.tolist(), self.face_offsets.tolist(), self.edge_offsets.tolist()
        out = []
        for b in range(self.B):
            v = self.verts[vo[b]:vo[b + 1]].cpu()
            f = (self.faces[fo[b]:fo[b + 1]].cpu().long() - self.index_base - vo[b])
            e = None if self.edges is None else (self.edges[eo[b]:eo[b + 1]].cpu().long() - vo[b])
            out.append((v, f, e))
        return out

    def check_flags(self) -> None:
        f = int(self.flags.item())
        if f:
            raise _lib.FohoStatusError("foho_dmc_extract", _lib.FOHO_E_WORKSPACE, f"surface extraction exceeded its capacities (flags {f}: "
                                       "bit0 vertices, bit1 faces, bit2 edges)")
