"""CPU ORACLE for the latent -> SDF decode (NOT built as a CUDA path yet) --
TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/`` may import it.

What the reference runs (third_party_patches/hy3dgen/shapegen/pipelines.py:292-312, ``latent2sdf``)::

    pred = 1 / vae.scale_factor * pred
    pred = vae(pred)                                   # ShapeVAE.forward: post_kl + 16-layer transformer
    for 8000-query chunks of the (res+1)^3 lattice:    # fp16 queries
        logits = vae.geo_decoder(queries[None], pred)  # Fourier embedding -> cross attention -> MLP -> 1 logit
    grid_logits = -cat(logits).view(1, D, D, D).float()

PARITY UNPINNED, doubly: the networks live in the un-vendored ``hy3dgen`` package (Hunyuan3D-2 @
e664e7471642c09921d23baaeba8ebe79bd6c48b, README.md:41) and their weights are not available offline.  The
module structure below restates ``hy3dgen/shapegen/models`` of that commit FROM MEMORY (class and parameter
names kept so a state_dict would load: ``post_kl``, ``transformer.resblocks.N.{ln_1,attn.c_qkv,attn.c_proj,
attn.attention.{q_norm,k_norm},ln_2,mlp.c_fc,mlp.c_proj}``, ``geo_decoder.{query_proj,cross_attn_decoder.
{ln_1,ln_2,ln_3,attn.c_q,attn.c_kv,attn.c_proj,attn.attention.{q_norm,k_norm},mlp},ln_post,output_proj}``);
hyper-parameters are those of the released ``hunyuan3d-dit-v2-0`` config (num_latents 3072, embed_dim 64,
width 1024, heads 16, 16 decoder layers, num_freqs 8, include_pi False, qkv_bias False, qk_norm True,
scale_factor 0.9990943042622529).  Every one of these must be re-checked against the package before a
CUDA path is declared parity-green; until then this file pins only what the reference's own call site fixes:
the 1/scale_factor, the chunking (results must not depend on it), the fp32 cast and the sign flip.

Two facts the kernel design for this row rests on (checked by tests/test_oracle_decoder.py):
  * the query side of the cross attention -- Fourier embedding, ``query_proj``, ``ln_1``, ``c_q``, ``q_norm`` --
    depends on the lattice and the weights only, not on the latents: it is computed once per lattice and
    re-used by every one of the ~750 decodes of an image and by every image;
  * everything after the attention is per-query (LayerNorm, MLP, 1-channel head), so queries can be tiled
    freely and the [D^3, width] activations never need to exist in HBM.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

SCALE_FACTOR = 0.9990943042622529
LN_EPS = 1e-6


class FourierEmbedder(nn.Module):
    """x -> [x, sin(x f_k), cos(x f_k)], f_k = 2^k (k < num_freqs), times pi when include_pi."""

    def __init__(self, num_freqs: int = 8, input_dim: int = 3, include_pi: bool = False):
        super().__init__()
        f = 2.0 ** torch.arange(num_freqs, dtype=torch.float32)
        if include_pi:
            f = f * math.pi
        self.register_buffer("frequencies", f, persistent=False)
        self.out_dim = input_dim * (2 * num_freqs + 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        e = (x[..., None] * self.frequencies.to(x.dtype)).reshape(*x.shape[:-1], -1)
        return torch.cat([x, e.sin(), e.cos()], dim=-1)


class MLP(nn.Module):
    def __init__(self, width: int):
        super().__init__()
        self.c_fc = nn.Linear(width, 4 * width)
        self.c_proj = nn.Linear(4 * width, width)

    def forward(self, x):
        return self.c_proj(F.gelu(self.c_fc(x)))


class QKVMultiheadAttention(nn.Module):
    def __init__(self, heads: int, width: int, qk_norm: bool):
        super().__init__()
        self.heads = heads
        hd = width // heads
        self.q_norm = nn.LayerNorm(hd, eps=LN_EPS) if qk_norm else nn.Identity()
        self.k_norm = nn.LayerNorm(hd, eps=LN_EPS) if qk_norm else nn.Identity()

    def forward(self, qkv):
        bs, n, w3 = qkv.shape
        qkv = qkv.view(bs, n, self.heads, -1)
        q, k, v = torch.split(qkv, w3 // self.heads // 3, dim=-1)
        q, k = self.q_norm(q), self.k_norm(k)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2).reshape(bs, n, -1)


class MultiheadAttention(nn.Module):
    def __init__(self, width: int, heads: int, qkv_bias: bool, qk_norm: bool):
        super().__init__()
        self.c_qkv = nn.Linear(width, 3 * width, bias=qkv_bias)
        self.c_proj = nn.Linear(width, width)
        self.attention = QKVMultiheadAttention(heads, width, qk_norm)

    def forward(self, x):
        return self.c_proj(self.attention(self.c_qkv(x)))


class ResidualAttentionBlock(nn.Module):
    def __init__(self, width: int, heads: int, qkv_bias: bool, qk_norm: bool):
        super().__init__()
        self.attn = MultiheadAttention(width, heads, qkv_bias, qk_norm)
        self.ln_1 = nn.LayerNorm(width, eps=LN_EPS)
        self.mlp = MLP(width)
        self.ln_2 = nn.LayerNorm(width, eps=LN_EPS)

    def forward(self, x):
        x = x + self.attn(self.ln_1(x))
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, qkv_bias: bool, qk_norm: bool):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads, qkv_bias, qk_norm) for _ in range(layers)])

    def forward(self, x):
        for blk in self.resblocks:
            x = blk(x)
        return x


class QKVMultiheadCrossAttention(nn.Module):
    def __init__(self, heads: int, width: int, qk_norm: bool):
        super().__init__()
        self.heads = heads
        hd = width // heads
        self.q_norm = nn.LayerNorm(hd, eps=LN_EPS) if qk_norm else nn.Identity()
        self.k_norm = nn.LayerNorm(hd, eps=LN_EPS) if qk_norm else nn.Identity()

    def project_q(self, q):
        bs, n, _ = q.shape
        return self.q_norm(q.view(bs, n, self.heads, -1))

    def attend(self, qn, kv):
        bs, n = qn.shape[:2]
        kv = kv.view(bs, kv.shape[1], self.heads, -1)
        k, v = torch.split(kv, kv.shape[-1] // 2, dim=-1)
        k = self.k_norm(k)
        o = F.scaled_dot_product_attention(qn.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2).reshape(bs, n, -1)

    def forward(self, q, kv):
        return self.attend(self.project_q(q), kv)


class MultiheadCrossAttention(nn.Module):
    def __init__(self, width: int, heads: int, qkv_bias: bool, qk_norm: bool):
        super().__init__()
        self.c_q = nn.Linear(width, width, bias=qkv_bias)
        self.c_kv = nn.Linear(width, 2 * width, bias=qkv_bias)
        self.c_proj = nn.Linear(width, width)
        self.attention = QKVMultiheadCrossAttention(heads, width, qk_norm)

    def forward(self, x, data):
        return self.c_proj(self.attention(self.c_q(x), self.c_kv(data)))


class ResidualCrossAttentionBlock(nn.Module):
    def __init__(self, width: int, heads: int, qkv_bias: bool, qk_norm: bool):
        super().__init__()
        self.attn = MultiheadCrossAttention(width, heads, qkv_bias, qk_norm)
        self.ln_1 = nn.LayerNorm(width, eps=LN_EPS)
        self.ln_2 = nn.LayerNorm(width, eps=LN_EPS)
        self.mlp = MLP(width)
        self.ln_3 = nn.LayerNorm(width, eps=LN_EPS)

    def forward(self, x, data):
        x = x + self.attn(self.ln_1(x), self.ln_2(data))
        return x + self.mlp(self.ln_3(x))


class CrossAttentionDecoder(nn.Module):
    def __init__(self, fourier_embedder: FourierEmbedder, width: int, heads: int, qkv_bias: bool, qk_norm: bool,
                 out_channels: int = 1):
        super().__init__()
        self.fourier_embedder = fourier_embedder
        self.query_proj = nn.Linear(fourier_embedder.out_dim, width)
        self.cross_attn_decoder = ResidualCrossAttentionBlock(width, heads, qkv_bias, qk_norm)
        self.ln_post = nn.LayerNorm(width)
        self.output_proj = nn.Linear(width, out_channels)

    def forward(self, queries, latents):
        x = self.query_proj(self.fourier_embedder(queries).to(latents.dtype))
        x = self.cross_attn_decoder(x, latents)
        return self.output_proj(self.ln_post(x))

    # ---- the same computation split where the kernel design splits it
    def precompute_queries(self, queries, dtype=None):
        """Everything on the query side that does not depend on the latents: the residual stream entry
        ``x0 = query_proj(embed(q))`` and the normalised per-head queries ``q_norm(c_q(ln_1(x0)))``."""
        e = self.fourier_embedder(queries)
        x0 = self.query_proj(e if dtype is None else e.to(dtype))
        blk = self.cross_attn_decoder
        qn = blk.attn.attention.project_q(blk.attn.c_q(blk.ln_1(x0)))
        return x0, qn

    def decode_precomputed(self, x0, qn, latents):
        blk = self.cross_attn_decoder
        kv = blk.attn.c_kv(blk.ln_2(latents))                      # once per decode: [B, 3072, 2 width]
        x = x0 + blk.attn.c_proj(blk.attn.attention.attend(qn, kv))
        x = x + blk.mlp(blk.ln_3(x))
        return self.output_proj(self.ln_post(x))


class ShapeVAE(nn.Module):
    """Decode half of Hunyuan3D-2's ShapeVAE (the encoder is not on the path)."""

    def __init__(self, num_latents: int = 3072, embed_dim: int = 64, width: int = 1024, heads: int = 16,
                 num_decoder_layers: int = 16, num_freqs: int = 8, include_pi: bool = False, qkv_bias: bool = False,
                 qk_norm: bool = True, scale_factor: float = SCALE_FACTOR):
        super().__init__()
        self.fourier_embedder = FourierEmbedder(num_freqs=num_freqs, include_pi=include_pi)
        self.post_kl = nn.Linear(embed_dim, width)
        self.transformer = Transformer(width, num_decoder_layers, heads, qkv_bias, qk_norm)
        self.geo_decoder = CrossAttentionDecoder(self.fourier_embedder, width, heads, qkv_bias, qk_norm)
        self.scale_factor = scale_factor
        self.latent_shape = (num_latents, embed_dim)

    def forward(self, latents):
        return self.transformer(self.post_kl(latents))


def latent2sdf(pred: torch.Tensor, xyz_samples: torch.Tensor, grid_size, vae: ShapeVAE, num_chunks: int = 8000,
               query_dtype: Optional[torch.dtype] = torch.float16) -> torch.Tensor:
    """REF pipelines.py:292-312.  ``pred`` [1, num_latents, embed_dim]; ``xyz_samples`` [D^3, 3] (the lattice of
    ``generate_dense_grid_points``); returns the NEGATED logits [1, D, D, D] float32 (negative inside).
    ``query_dtype``: the reference rounds the query coordinates to fp16 before embedding them (:302); pass None
    to keep them as given (fp64 checks)."""
    pred = 1 / vae.scale_factor * pred
    pred = vae(pred)
    batch_logits = []
    for start in range(0, xyz_samples.shape[0], num_chunks):
        queries = xyz_samples[start:start + num_chunks, :]
        if query_dtype is not None:
            queries = queries.to(query_dtype)
        logits = vae.geo_decoder(queries.unsqueeze(0).to(pred.dtype), pred)
        batch_logits.append(logits)
    grid_logits = torch.cat(batch_logits, dim=1)
    grid_logits = grid_logits.view((1, grid_size[0], grid_size[1], grid_size[2])).float()
    return -grid_logits


def decode_flops(n_queries: int, num_latents: int = 3072, width: int = 1024, layers: int = 16) -> dict:
    """Forward multiply-add FLOPs (2 per MAC) of one ``latent2sdf`` call, for the roofline of this row."""
    tr_layer = 2 * num_latents * (4 * width * width + 8 * width * width) + 4 * num_latents * num_latents * width
    transformer = layers * tr_layer + 2 * num_latents * 64 * width
    kv = 2 * num_latents * width * 2 * width
    q_side = 2 * n_queries * (51 * width + width * width)            # latent-independent: once per lattice
    attn = 4 * n_queries * num_latents * width
    post = 2 * n_queries * (width * width + 8 * width * width + width)
    return {"transformer": transformer, "kv": kv, "query_side_once": q_side, "cross_attention": attn,
            "per_query_mlp": post, "per_decode": transformer + kv + attn + post}
