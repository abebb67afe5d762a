"""Mesh -> SDF on a shared lattice and the penetration count, on the GPU.

Mirror of the reference helpers in ``third_party/utilz/kaolin_sdf_ops.py`` (same names and
argument meaning): ``generate_dense_grid_points`` (:26-45), ``mesh2sdf`` (:88-109),
``get_sdf_of_meshes`` (:131-160), and ``honerf_intersection_loss``
(third_party_patches/hy3dgen/shapegen/pipelines.py:231-239).  Meshes are (verts [V,3] f32 CUDA,
faces [F,3] i32 CUDA) pairs instead of pytorch3d ``Meshes``.  Unlike the reference, the grid is
never rebuilt on the CPU: only three (res+1)-long coordinate arrays are uploaded.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np
import torch

from .. import _lib


def lattice_axes(bbox_min, bbox_max, cells: int):
    """The three float32 coordinate arrays of a (cells+1)^3 lattice spanning a bbox."""
    n = int(cells) + 1
    return [np.linspace(bbox_min[a], bbox_max[a], n, dtype=np.float32) for a in range(3)]


def generate_dense_grid_points(bbox_min: np.ndarray, bbox_max: np.ndarray, octree_depth: int, indexing: str = "ij",
                               octree_resolution: int = None):
    """Same contract as the reference helper (pipelines.py:341-360, used at kaolin_sdf_ops.py:146-152):
    2**octree_depth cells per axis unless ``octree_resolution`` overrides it; returns
    (xyz [N,3] f32 in ``indexing`` order, [n,n,n], bbox extent).  The product path never materialises
    xyz -- it hands ``lattice_axes`` to the kernel -- this exists for callers and tests."""
    cells = int(octree_resolution) if octree_resolution is not None else int(2 ** octree_depth)
    axes = lattice_axes(bbox_min, bbox_max, cells)
    xyz = np.stack(np.meshgrid(*axes, indexing=indexing), axis=-1).reshape(-1, 3)
    return xyz, [cells + 1] * 3, np.asarray(bbox_max) - np.asarray(bbox_min)


def mesh2sdf_axes(verts: torch.Tensor, faces: torch.Tensor, xs, ys, zs) -> torch.Tensor:
    """SDF of a mesh on the rectilinear lattice xs × ys × zs ("ij" order): [nx,ny,nz] f32."""
    lib = _lib.load()
    if not verts.is_cuda:
        raise _lib.FohoLibraryError("mesh2sdf needs CUDA tensors; there is no CPU fallback")
    dev = verts.device
    v = verts.detach().to(torch.float32).contiguous()
    f = faces.to(torch.int32).contiguous()
    ax = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev) if not torch.is_tensor(a)
          else a.to(dev, torch.float32).contiguous() for a in (xs, ys, zs)]
    nx, ny, nz = [int(a.numel()) for a in ax]
    out = torch.empty(nx, ny, nz, dtype=torch.float32, device=dev)
    nbytes = lib.foho_mesh2sdf_workspace_bytes(v.shape[0], f.shape[0], nx, ny, nz)
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    ws_ptr = ws.data_ptr() + ((-ws.data_ptr()) % 256)
    with torch.cuda.device(dev):
        s = torch.cuda.current_stream(dev)
        _lib.check("foho_mesh2sdf_lattice", lib.foho_mesh2sdf_lattice(
            v.data_ptr(), v.shape[0], f.data_ptr(), f.shape[0], ax[0].data_ptr(), ax[1].data_ptr(), ax[2].data_ptr(),
            nx, ny, nz, out.data_ptr(), C.c_void_p(ws_ptr), nbytes, C.c_void_p(s.cuda_stream)))
    return out


def mesh2sdf(mesh: Tuple[torch.Tensor, torch.Tensor], grid_axes, device="cuda", resolution=64) -> torch.Tensor:
    """kaolin_sdf_ops.py:88-109: flattened sdf [(res+1)^3], negative inside."""
    return mesh2sdf_axes(mesh[0], mesh[1], *grid_axes).reshape(-1)


def get_sdf_of_meshes(mesh1, mesh2, device="cuda", resolution=64):
    """kaolin_sdf_ops.py:131-160: SDFs of two meshes on the (res+1)^3 grid spanning their union bbox."""
    v1, v2 = mesh1[0].detach(), mesh2[0].detach()
    bbox_min = torch.minimum(v1.min(dim=0)[0], v2.min(dim=0)[0]).cpu().numpy()
    bbox_max = torch.maximum(v1.max(dim=0)[0], v2.max(dim=0)[0]).cpu().numpy()
    axes = lattice_axes(bbox_min, bbox_max, resolution)
    return mesh2sdf(mesh1, axes, device, resolution), mesh2sdf(mesh2, axes, device, resolution)


def intersection_count(sdf_hand: torch.Tensor, sdf_obj: torch.Tensor) -> torch.Tensor:
    """Number of lattice points inside both meshes: int64 tensor [1] on the device."""
    lib = _lib.load()
    if not (sdf_hand.is_cuda and sdf_obj.is_cuda):
        raise _lib.FohoLibraryError("intersection_count needs CUDA tensors; there is no CPU fallback")
    a = sdf_hand.detach().to(torch.float32).contiguous().view(-1)
    b = sdf_obj.detach().to(torch.float32).contiguous().view(-1)
    cnt = torch.zeros(1, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        s = torch.cuda.current_stream(a.device)
        _lib.check("foho_intersection_count", lib.foho_intersection_count(
            a.data_ptr(), b.data_ptr(), a.numel(), cnt.data_ptr(), C.c_void_p(s.cuda_stream)))
    return cnt


def honerf_intersection_loss(sdf_hand: torch.Tensor, sdf_obj: torch.Tensor) -> torch.Tensor:
    """pipelines.py:231-239: count(sdf_obj<0 & sdf_hand<0)/1000 (no gradient); like the reference's
    ``penet_points_id.sum() / 1000`` this is an int64 device scalar divided by torch."""
    return intersection_count(sdf_hand, sdf_obj)[0] / 1000
