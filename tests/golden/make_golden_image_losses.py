#!/usr/bin/env python
"""Generate tests/golden/ref_golden_image_losses.npz by EXECUTING THE REFERENCE'S OWN CODE for the image-space
losses of the guidance loop (SURVEY.md section 8f rank 2 -- groundwork, no CUDA path yet).

Run in the authoring container only (needs /root/reference):  python tests/golden/make_golden_image_losses.py

Executed from third_party_patches/hy3dgen/shapegen/pipelines.py, cut out with ``ast`` (the module itself cannot
be imported offline): ``normal_alignment_loss`` (the second definition, :178-187, which is the one in force),
``render_normal_and_disparity`` (:272-289) with a stand-in renderer that returns prescribed shader output /
z-buffer tensors, and ``compute_loss_stable_fp32`` (:1001-1018); the disparity L1 and silhouette BCE are the
torch calls of :1567-1569.  Values AND the gradients autograd sends back to the renderer outputs are stored.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/third_party_patches/hy3dgen/shapegen/pipelines.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_image_losses.npz")


def extract(names):
    src = open(REF).read()
    ns = {"torch": torch, "F": F, "np": np}
    for node in ast.parse(src).body:                       # later definitions override earlier ones, as on import
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), REF, "exec"), ns)
    return ns


class _Frag:
    def __init__(self, zbuf):
        self.zbuf = zbuf


class _Renderer:
    """MeshRenderer stand-in: ``renderer(mesh)`` -> shader output [1,H,W,4]; ``renderer.rasterizer(mesh).zbuf``."""

    def __init__(self, norms, zbuf):
        self._n, self._z = norms, zbuf

    def __call__(self, mesh):
        return self._n

    def rasterizer(self, mesh):
        return _Frag(self._z.clone())                      # the reference writes into zbuf in place (:281)


def main():
    ns = extract({"normal_alignment_loss", "render_normal_and_disparity", "compute_loss_stable_fp32"})
    g = torch.Generator().manual_seed(7)
    H, W = 24, 32
    out = {}
    for tag, cover in (("a", 0.6), ("b", 0.25)):
        fg = torch.rand(1, H, W, generator=g) < cover                                  # pixels a face covers
        n = torch.randn(1, H, W, 3, generator=g)
        norms = torch.cat([n * fg[..., None], fg[..., None].float()], -1)              # background: zeros, alpha 0
        zbuf = torch.where(fg, 0.5 + torch.rand(1, H, W, generator=g), torch.tensor(-1.0))[..., None]   # -1 = no face
        sil = torch.rand(1, H, W, generator=g).clamp(1e-4, 1 - 1e-4)
        gt_n = torch.randn(1, H, W, 3, generator=g)
        gt_mask = torch.rand(1, H, W, generator=g) < 0.7
        gt_disp = torch.rand(1, H, W, generator=g)
        gt_sil = (torch.rand(1, H, W, generator=g) < 0.5)
        norms_l = norms.clone().requires_grad_(True)
        zbuf_l = zbuf.clone().requires_grad_(True)
        sil_l = sil.clone().requires_grad_(True)
        rn, rd = ns["render_normal_and_disparity"](_Renderer(norms_l, zbuf_l), None)
        l_n = ns["normal_alignment_loss"](rn, gt_n, valid_mask=gt_mask)                # :1567
        l_d = F.l1_loss(rd, gt_disp)                                                   # :1568
        l_s = torch.nn.functional.binary_cross_entropy(sil_l, gt_sil.float())          # :1569
        total = ns["compute_loss_stable_fp32"]({"n": 10 * l_n, "d": 10 * l_d, "s": 10 * l_s})
        total.backward()
        for k, v in (("norms", norms), ("zbuf", zbuf), ("sil", sil), ("gt_n", gt_n), ("gt_mask", gt_mask), ("gt_disp", gt_disp),
                     ("gt_sil", gt_sil), ("rn", rn.detach()), ("rd", rd.detach()), ("l_n", l_n.detach()), ("l_d", l_d.detach()),
                     ("l_s", l_s.detach()), ("total", total.detach()), ("g_norms", norms_l.grad), ("g_zbuf", zbuf_l.grad),
                     ("g_sil", sil_l.grad)):
            out[f"{tag}_{k}"] = v.numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
