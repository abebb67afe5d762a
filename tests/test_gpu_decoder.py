"""Row f1 on the GPU: the tensor-core latent -> SDF decoder (``latent2sdf``, pipelines.py:292-312) and its adjoint
against ``oracle/decoder_oracle.py`` in float32 with random weights (the architecture itself is restated from
memory -- PARITY UNPINNED against hy3dgen, see the oracle's header; what IS pinned here is that the CUDA path
computes the oracle's function and its gradient).  Tolerances: fp16 operands / fp32 accumulation against an fp32
oracle -> 3e-3 of the output range (measured 5e-4 .. 1.2e-3)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 3e-3


def _vae(layers, seed=0):
    from oracle import decoder_oracle as DO
    torch.manual_seed(seed)
    vae = DO.ShapeVAE(num_decoder_layers=layers).float()
    with torch.no_grad():
        for n, p in vae.named_parameters():          # non-trivial LayerNorm affine parameters and biases
            if n.endswith("norm.weight") or (".ln_" in n and n.endswith("weight")):
                p.add_(0.2 * torch.randn_like(p))
            elif n.endswith("bias"):
                p.add_(0.05 * torch.randn_like(p))
        vae.geo_decoder.output_proj.weight.mul_(4.0)
    return vae


def _lattice(D):
    axis = torch.linspace(-1.10, 1.10, D)
    return torch.stack(torch.meshgrid(axis, axis, axis, indexing="ij"), -1).reshape(-1, 3)


@pytest.fixture(scope="module")
def small():
    from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder
    from oracle import decoder_oracle as DO
    dev, D, B = "cuda:0", 13, 2
    vae = _vae(2)
    xyz = _lattice(D)
    W = DecoderWeights(vae.state_dict(), dev)
    dec = LatentDecoder(W, B, query_chunk=1024, active_chunk=256)
    dec.set_queries(xyz)
    lat = torch.randn(B, 3072, 64, device=dev)
    vae_g = vae.to(dev)
    lat_o = lat.clone().requires_grad_(True)
    ref = torch.stack([DO.latent2sdf(lat_o[b:b + 1], xyz.to(dev), (D, D, D), vae_g).reshape(-1) for b in range(B)])
    return dict(dec=dec, W=W, vae=vae_g, xyz=xyz, lat=lat, lat_o=lat_o, ref=ref, D=D, B=B, DO=DO)


def test_forward_matches_the_oracle(small):
    s = small
    sdf = s["dec"].forward(s["lat"])
    torch.cuda.synchronize()
    assert sdf.dtype == torch.float32 and sdf.shape == (s["B"], s["D"] ** 3)          # the `.float()` of pipelines.py:309
    ref = s["ref"].detach()
    assert (sdf - ref).abs().max().item() <= TOL * ref.abs().max().item()
    # the sign flip of :311-312 and the 1/scale_factor of :297 are part of what is compared (the oracle applies both)
    with torch.no_grad():
        data_ref = s["vae"](s["lat"] / s["vae"].scale_factor)
    err = (s["dec"].data.float().view(s["B"], 3072, 1024) - data_ref).abs().max().item()
    assert err <= TOL * data_ref.abs().max().item()


def test_result_does_not_depend_on_the_query_chunking(small):
    """The reference decodes in 8000-query chunks (pipelines.py:300-306); results must not depend on the chunk."""
    from followmyhold_b200.decoder.shapevae import LatentDecoder
    s = small
    a = s["dec"].forward(s["lat"]).clone()
    d2 = LatentDecoder(s["W"], s["B"], query_chunk=640, active_chunk=256)
    d2.set_queries(s["xyz"])
    b = d2.forward(s["lat"])
    torch.cuda.synchronize()
    assert torch.equal(a, b)


def test_adjoint_matches_autograd_through_the_oracle(small):
    s = small
    B, D = s["B"], s["D"]
    s["dec"].forward(s["lat"])
    g = torch.Generator().manual_seed(1)
    M = 520                                              # spans three active chunks of 256, the last one ragged
    idx = torch.randint(0, D ** 3, (B, M), generator=g).to(torch.int32).cuda()
    gs = (torch.randn(B, M, generator=g) * 1e-2).cuda()
    gs[:, -8:] = 0                                       # padded entries: index anything, gradient 0
    E = sum((s["ref"][b][idx[b].long()] * gs[b]).sum() for b in range(B))
    (gref,) = torch.autograd.grad(E, s["lat_o"], retain_graph=True)
    got = s["dec"].backward(idx, gs)
    torch.cuda.synchronize()
    assert got.shape == (B, 3072, 64) and got.dtype == torch.float32
    assert (got - gref).abs().max().item() <= TOL * gref.abs().max().item()
    # repeated lattice indices accumulate like autograd does
    idx2 = idx.clone(); idx2[:, 1] = idx2[:, 0]
    E2 = sum((s["ref"][b][idx2[b].long()] * gs[b]).sum() for b in range(B))
    (gref2,) = torch.autograd.grad(E2, s["lat_o"], retain_graph=True)
    got2 = s["dec"].backward(idx2, gs)
    assert (got2 - gref2).abs().max().item() <= TOL * gref2.abs().max().item()


def test_adjoint_is_linear_in_the_incoming_gradient(small):
    """Size-independent property: the adjoint is a linear map of dE/dSDF (loss scaling must not break it)."""
    s = small
    B, D = s["B"], s["D"]
    s["dec"].forward(s["lat"])
    g = torch.Generator().manual_seed(3)
    idx = torch.randint(0, D ** 3, (B, 256), generator=g).to(torch.int32).cuda()
    g1 = (torch.randn(B, 256, generator=g) * 1e-2).cuda()
    g2 = (torch.randn(B, 256, generator=g) * 1e-2).cuda()
    a = s["dec"].backward(idx, g1).clone()
    b = s["dec"].backward(idx, g2).clone()
    c = s["dec"].backward(idx, g1 + 3 * g2)
    assert (c - (a + 3 * b)).abs().max().item() <= 3 * TOL * c.abs().max().item()


def test_reference_lattice_65_sampled_rows():
    """The reference's own lattice (65^3 = 274 625 queries, pipelines.py:1126-1137) with the full 16-layer
    transformer, one image: finite everywhere, and 4 096 sampled rows equal the oracle's decode of those rows."""
    from followmyhold_b200.decoder.shapevae import DecoderWeights, LatentDecoder
    from oracle import decoder_oracle as DO
    dev, D = "cuda:0", 65
    vae = _vae(16, seed=5)
    xyz = _lattice(D)
    dec = LatentDecoder(DecoderWeights(vae.state_dict(), dev), 1)
    dec.set_queries(xyz)
    lat = torch.randn(1, 3072, 64, device=dev)
    sdf = dec.forward(lat)
    torch.cuda.synchronize()
    assert torch.isfinite(sdf).all()
    vae_g = vae.to(dev)
    rows = torch.randperm(D ** 3, generator=torch.Generator().manual_seed(0))[:4096]
    with torch.no_grad():
        pred = vae_g(lat / vae.scale_factor)
        ref = -vae_g.geo_decoder(xyz[rows].to(dev).half().float()[None], pred).reshape(-1)
    got = sdf[0, rows.to(dev)]
    assert (got - ref).abs().max().item() <= 2 * TOL * ref.abs().max().item()
