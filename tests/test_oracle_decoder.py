"""Oracle of the latent -> SDF decode (no CUDA path yet).  The architecture is restated
from memory of the un-vendored hy3dgen package (parity unpinned, see oracle/decoder_oracle.py); these tests pin
what the reference's call site fixes (pipelines.py:292-312) and the two structural facts the kernel design uses."""
import numpy as np
import pytest
import torch

from oracle import decoder_oracle as DO


def small_vae(dtype=torch.float64, seed=0):
    torch.manual_seed(seed)
    vae = DO.ShapeVAE(num_latents=24, embed_dim=8, width=32, heads=4, num_decoder_layers=2).to(dtype)
    with torch.no_grad():                       # default inits leave LayerNorm affine at identity: randomise them
        for n, p in vae.named_parameters():
            if "ln" in n or "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    return vae


def lattice(D, dtype=torch.float64):
    ax = torch.linspace(-1.10, 1.10, D, dtype=dtype)
    return torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(-1, 3)        # pipelines.py:341-360


def test_call_site_semantics_chunking_scale_sign_dtype():
    vae = small_vae()
    D = 5
    xyz = lattice(D)
    pred = torch.randn(1, 24, 8, dtype=torch.float64)
    with torch.no_grad():
        full = DO.latent2sdf(pred, xyz, (D, D, D), vae, num_chunks=10 ** 9, query_dtype=None)
    assert full.shape == (1, D, D, D) and full.dtype == torch.float32                        # .float() at :309
    for chunk in (1, 7, 8000):                                                               # :299-306
        with torch.no_grad():
            assert torch.allclose(DO.latent2sdf(pred, xyz, (D, D, D), vae, num_chunks=chunk, query_dtype=None), full, atol=1e-6)
    # 1 / scale_factor is applied before the transformer (:294), the result is negated (:312)
    with torch.no_grad():
        lat = vae(pred / vae.scale_factor)
        direct = vae.geo_decoder(xyz[None], lat).view(1, D, D, D)
        # the reference rounds the query coordinates to fp16 (:302): a different, but close, set of sample points
        h = DO.latent2sdf(pred, xyz, (D, D, D), vae, query_dtype=torch.float16)
    assert torch.allclose(full.double(), -direct, atol=1e-6)
    assert 0 < float((h - full).abs().max()) < 0.05 * float(full.abs().max()) + 1e-3


def test_query_side_is_latent_independent_and_queries_tile_freely():
    vae = small_vae()
    dec = vae.geo_decoder
    xyz = lattice(4)
    x0, qn = dec.precompute_queries(xyz[None])                                               # once per lattice
    for seed in (1, 2):
        lat = vae(torch.randn(1, 24, 8, dtype=torch.float64, generator=torch.Generator().manual_seed(seed)))
        ref = dec(xyz[None], lat)
        assert torch.allclose(dec.decode_precomputed(x0, qn, lat), ref, atol=1e-12)
        tiles = torch.cat([dec.decode_precomputed(x0[:, a:a + 9], qn[:, a:a + 9], lat) for a in range(0, 64, 9)], 1)
        assert torch.allclose(tiles, ref, atol=1e-12)


def test_gradient_to_the_velocity_leaf_matches_finite_differences():
    """The loop differentiates an energy of the volume w.r.t. the model output through step_final and this
    decode (pipelines.py:1507-1508,1600): dE/dpred by autograd vs central differences in float64."""
    vae = small_vae()
    D = 3
    xyz = lattice(D)
    w = torch.randn(1, D, D, D, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    pred = torch.randn(1, 24, 8, dtype=torch.float64, generator=torch.Generator().manual_seed(4)).requires_grad_(True)

    def energy(p):
        vol = -torch.cat([vae.geo_decoder(xyz[None], vae(p / vae.scale_factor))], 1).view(1, D, D, D)
        return (w * torch.relu(-vol)).sum()

    e = energy(pred)
    e.backward()
    g = pred.grad.clone()
    rng = np.random.default_rng(0)
    for _ in range(6):
        i, j = int(rng.integers(24)), int(rng.integers(8))
        d = torch.zeros_like(pred); d[0, i, j] = 1e-6
        fd = (energy(pred.detach() + d) - energy(pred.detach() - d)) / 2e-6
        assert abs(float(fd.detach()) - float(g[0, i, j])) <= 1e-6 * max(1.0, abs(float(fd.detach())))


def test_state_dict_names_and_flop_accounting():
    vae = DO.ShapeVAE(num_latents=8, embed_dim=4, width=16, heads=2, num_decoder_layers=1)
    names = set(vae.state_dict())
    for n in ("post_kl.weight", "transformer.resblocks.0.attn.c_qkv.weight", "transformer.resblocks.0.attn.attention.q_norm.weight",
              "transformer.resblocks.0.mlp.c_fc.bias", "geo_decoder.query_proj.weight", "geo_decoder.cross_attn_decoder.attn.c_kv.weight",
              "geo_decoder.cross_attn_decoder.attn.attention.k_norm.bias", "geo_decoder.cross_attn_decoder.ln_3.weight",
              "geo_decoder.ln_post.weight", "geo_decoder.output_proj.weight"):
        assert n in names
    assert "transformer.resblocks.0.attn.c_qkv.bias" not in names                           # qkv_bias False
    assert vae.fourier_embedder.out_dim == 51 and vae.geo_decoder.query_proj.in_features == 51
    f = DO.decode_flops(65 ** 3)                                                             # the loop's lattice (:1126-1137)
    assert 8.5e12 < f["cross_attention"] + f["per_query_mlp"] < 8.8e12                      # ~8.6 TFLOP per decode forward
    assert 1.8e12 < f["transformer"] < 1.95e12
    assert f["query_side_once"] < 0.07 * f["per_decode"]
