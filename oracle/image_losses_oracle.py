"""CPU ORACLE for the image-space losses of the guidance loop (groundwork for a fused
rasterise-and-score kernel; no CUDA path yet) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, in numpy float64 with ANALYTIC backward passes (what a kernel has to implement; the reference gets
them from autograd), what the reference does to the renderer outputs in every inner iteration
(third_party_patches/hy3dgen/shapegen/pipelines.py):

  * ``render_normal_and_disparity`` :272-289 -- global min/max normalisation of the shaded normal map (background
    zeroed by the alpha mask) and of the disparity 1/(z + 1e-6) with background depth set to 10;
  * ``normal_alignment_loss`` :178-187 -- mean over the valid mask of 1 - cos(rendered, target);
  * ``F.l1_loss`` on the disparities :1568 and ``binary_cross_entropy`` on the soft silhouette :1569;
  * the weights 10 / 10 / 10 of :1580-1583, summed in fp32 (``compute_loss_stable_fp32`` :1001-1018).

PINNED by tests/golden/ref_golden_image_losses.npz, which tests/golden/make_golden_image_losses.py produced by
executing those reference functions (values and the gradients autograd returns to the renderer outputs).
The renderer itself (pytorch3d rasteriser + shaders) is not restated here.
"""
from __future__ import annotations

import numpy as np

EPS_RANGE = 1e-6      # :280, :284, :285
BG_DEPTH = 10.0       # :283


def _minmax_backward(x, g_y, lo, hi, r):
    """y = (x - lo) / r, r = hi - lo + eps, lo = min(x), hi = max(x) over the whole array: gradient w.r.t. x.
    torch's full-reduction min/max spread their gradient evenly over tied extrema."""
    y = (x - lo) / r
    g = g_y / r
    g_lo = float((-(1.0 - y) / r * g_y).sum())
    g_hi = float((-y / r * g_y).sum())
    is_lo, is_hi = x == lo, x == hi
    g = g + is_lo * (g_lo / is_lo.sum()) + is_hi * (g_hi / is_hi.sum())
    return g


def normals_forward(norms4):
    """:274-281.  norms4 [..., 4] = shader output (xyz normal, alpha).  Returns (normalised normals, cache)."""
    n = norms4[..., :3].astype(np.float64)
    mask = norms4[..., 3] > 0.0
    lo, hi = n.min(), n.max()
    r = hi - lo + EPS_RANGE
    rn = (n - lo) / r
    rn = rn * mask[..., None]
    return rn, (n, mask, lo, hi, r)


def normals_backward(g_rn, cache):
    n, mask, lo, hi, r = cache
    g = _minmax_backward(n, g_rn * mask[..., None], lo, hi, r)
    return np.concatenate([g, np.zeros(g.shape[:-1] + (1,))], -1)          # no gradient to alpha


def disparity_forward(zbuf):
    """:275, :283-285.  zbuf [..., 1], negative = no face."""
    z = zbuf[..., 0].astype(np.float64)
    bg = z < 0
    z = np.where(bg, BG_DEPTH, z)
    d = 1.0 / (z + EPS_RANGE)
    lo, hi = d.min(), d.max()
    r = hi - lo + EPS_RANGE
    return (d - lo) / r, (z, bg, d, lo, hi, r)


def disparity_backward(g_rd, cache):
    z, bg, d, lo, hi, r = cache
    g_d = _minmax_backward(d, g_rd, lo, hi, r)
    g_z = -g_d / (z + EPS_RANGE) ** 2
    return np.where(bg, 0.0, g_z)[..., None]                               # overwritten pixels: no gradient


def normal_alignment_loss(rn, gt_n, valid_mask):
    """:178-187 (F.normalize: x / max(|x|, 1e-12)).  Returns (loss, dloss/drn)."""
    gt_n = gt_n.astype(np.float64)
    ln = np.maximum(np.linalg.norm(rn, axis=-1, keepdims=True), 1e-12)
    lg = np.maximum(np.linalg.norm(gt_n, axis=-1, keepdims=True), 1e-12)
    u, g = rn / ln, gt_n / lg
    cos = (u * g).sum(-1)
    m = valid_mask.astype(bool)
    nv = m.sum()
    loss = float((1.0 - cos)[m].mean())
    d_u = -g * m[..., None] / nv
    # d(x/|x|) = (I - u u^T)/|x|  (where |x| > eps; background pixels have x = 0 and get no gradient: u = 0)
    d_rn = (d_u - (d_u * u).sum(-1, keepdims=True) * u) / ln
    d_rn = np.where(np.linalg.norm(rn, axis=-1, keepdims=True) > 1e-12, d_rn, d_u / 1e-12)
    return loss, d_rn


def l1_loss(x, t):
    diff = x - t.astype(np.float64)
    return float(np.abs(diff).mean()), np.sign(diff) / diff.size


def bce_loss(s, t):
    """``binary_cross_entropy`` (logs clamped at -100 like torch)."""
    s = s.astype(np.float64); t = t.astype(np.float64)
    loss = -(t * np.maximum(np.log(s), -100.0) + (1 - t) * np.maximum(np.log1p(-s), -100.0))
    return float(loss.mean()), (s - t) / (s * (1 - s)) / s.size


def image_losses(norms4, zbuf, sil, gt_n, gt_mask, gt_disp, gt_sil, w_normal=10.0, w_disp=10.0, w_sil=10.0):
    """The three image terms of the phase-2 total (:1567-1569, 1580-1583) and their gradients w.r.t. the
    renderer outputs.  Returns a dict with l_n, l_d, l_s, total, rn, rd, g_norms, g_zbuf, g_sil."""
    rn, cn = normals_forward(norms4)
    rd, cd = disparity_forward(zbuf)
    l_n, g_rn = normal_alignment_loss(rn, gt_n, gt_mask)
    l_d, g_rd = l1_loss(rd, gt_disp)
    l_s, g_s = bce_loss(sil, gt_sil)
    return {"l_n": l_n, "l_d": l_d, "l_s": l_s, "total": w_normal * l_n + w_disp * l_d + w_sil * l_s, "rn": rn, "rd": rd,
            "g_norms": normals_backward(w_normal * g_rn, cn), "g_zbuf": disparity_backward(w_disp * g_rd, cd),
            "g_sil": w_sil * g_s}
