"""Oracle of the image-space losses (groundwork for SURVEY.md section 8f rank 2) against golden vectors produced by
EXECUTING the reference's own ``render_normal_and_disparity`` / ``normal_alignment_loss`` /
``compute_loss_stable_fp32`` (tests/golden/make_golden_image_losses.py): values and the gradients autograd
returns to the renderer outputs vs the oracle's analytic backward."""
import os

import numpy as np
import pytest

from oracle import image_losses_oracle as IL

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden_image_losses.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_values_and_analytic_gradients_match_the_reference(tag):
    g = lambda k: G[f"{tag}_{k}"]
    out = IL.image_losses(g("norms"), g("zbuf"), g("sil"), g("gt_n"), g("gt_mask"), g("gt_disp"), g("gt_sil"))
    assert np.allclose(out["rn"], g("rn"), rtol=1e-5, atol=1e-6)             # fp32 reference vs fp64 oracle
    assert np.allclose(out["rd"], g("rd"), rtol=1e-5, atol=1e-6)
    for k in ("l_n", "l_d", "l_s", "total"):
        assert out[k] == pytest.approx(float(g(k)), rel=2e-6), k
    for k in ("g_norms", "g_zbuf", "g_sil"):
        ref = g(k)
        assert out[k].shape == ref.shape
        assert np.abs(out[k] - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-9, k
    # structure the kernel can rely on: background pixels and the alpha channel receive no gradient, except
    # through the global extrema of the normalisation
    bg = g("zbuf")[..., 0] < 0
    assert (out["g_zbuf"][bg] == 0).all() and (out["g_norms"][..., 3] == 0).all()


def test_background_depth_and_ties():
    z = np.full((1, 2, 3, 1), -1.0); z[0, 0, 0, 0] = 2.0; z[0, 1, 2, 0] = 4.0
    rd, cache = IL.disparity_forward(z)
    assert rd.max() == pytest.approx(1.0, abs=1e-5) and rd.min() == 0.0      # background depth 10 is the far plane
    assert np.isclose(rd[0, 0, 1], 0.0)
    g = IL.disparity_backward(np.ones_like(rd), cache)
    assert g.shape == z.shape and (g[z < 0] == 0).all()
