"""The reference's ``latent2sdf`` (third_party_patches/hy3dgen/shapegen/pipelines.py:292-312) and its adjoint
on the B200's tensor cores: host-side orchestration of the C-ABI building blocks (``foho_tc_gemm``,
``foho_tc_attention``, ``foho_dec_*``).  Nothing here computes with torch; torch owns the buffers and streams.

    pred = 1 / vae.scale_factor * pred
    pred = vae(pred)                                   # post_kl + 16-layer transformer over 3072 tokens
    logits = vae.geo_decoder(queries, pred)            # Fourier embedding -> cross attention -> MLP -> 1 logit
    sdf = -logits.view(1, D, D, D).float()

Module / parameter names follow ``hy3dgen/shapegen/models`` (Hunyuan3D-2 @ e664e747, un-vendored; restated from
memory in ``oracle/decoder_oracle.py`` -- ARCHITECTURE UNPINNED until checked against the package) so a released
state_dict loads by name: ``post_kl``, ``transformer.resblocks.N.*``, ``geo_decoder.*``.

What is latent-independent is computed once per lattice (``set_queries``): the residual-stream entry
``x0 = query_proj(embed(q))`` and the normalised per-head queries ``q_norm(c_q(ln_1(x0)))``.

Adjoint: weights are frozen, so only input gradients exist.  ``backward`` takes dE/dSDF on a sparse set of
lattice points (the voxels the energy touches), recomputes the query-side activations of those rows only, and
carries the gradient through the cross attention into K/V, through the token transformer (activations kept from
the forward), to dE/d(latents).  fp16 gradient activations carry a static loss scale.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from .. import _lib
from . import tc

WIDTH, HEADS, HD, TOKENS, EMBED = 1024, 16, 64, 3072, 64
LN_EPS = 1e-6            # hy3dgen blocks: LayerNorm(eps=1e-6); ln_post keeps torch's default 1e-5
LN_POST_EPS = 1e-5
SCALE_FACTOR = 0.9990943042622529
LOG2E = 1.4426950408889634


def _sp(stream):
    return C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)


class _Ops:
    """ctypes fronts of the row-wise kernels (raw pointers in, status checked)."""

    def __init__(self):
        self.lib = _lib.load()

    def layernorm(self, x, w, b, out, width=WIDTH, eps=LN_EPS, stream=None):
        """x, out: [rows, width] (stride(0) free) or [tokens, heads, 64] strided views."""
        (ix, lox, lix, rows), (iy, loy, liy, _) = self._view(x, width), self._view(out, width)
        _lib.check("foho_dec_layernorm", self.lib.foho_dec_layernorm(
            x.data_ptr(), ix, lox, lix, None if w is None else w.data_ptr(), None if b is None else b.data_ptr(), eps,
            out.data_ptr(), iy, loy, liy, rows, width, _sp(stream)))
        return out

    def layernorm_bwd(self, x, w, dy, dx, add=None, width=WIDTH, eps=LN_EPS, stream=None):
        (ix, lox, lix, rows), (ig, log_, lig, _), (id_, lod, lid, _) = self._view(x, width), self._view(dy, width), self._view(dx, width)
        if add is not None and (add.stride() != dx.stride() or add.shape != dx.shape):
            raise ValueError("`add` must have the layout of dx")
        _lib.check("foho_dec_layernorm_bwd", self.lib.foho_dec_layernorm_bwd(
            x.data_ptr(), ix, lox, lix, None if w is None else w.data_ptr(), eps, dy.data_ptr(), ig, log_, lig,
            None if add is None else add.data_ptr(), dx.data_ptr(), id_, lod, lid, rows, width, _sp(stream)))
        return dx

    @staticmethod
    def _view(t, width):
        if t.stride(-1) != 1 or t.shape[-1] != width:
            raise ValueError("last dimension must be contiguous and equal to the LayerNorm width")
        if t.dim() == 2:
            return 1, t.stride(0), 0, t.shape[0]
        if t.dim() == 3:
            return t.shape[1], t.stride(0), t.stride(1), t.shape[0] * t.shape[1]
        raise ValueError("expected a 2-D or 3-D view")

    def softmax(self, S, P, stream=None):
        T = S.shape[-1]
        _lib.check("foho_dec_softmax", self.lib.foho_dec_softmax(S.data_ptr(), P.data_ptr(), S.numel() // T, T, _sp(stream)))
        return P

    def softmax_bwd(self, P, dP, dS, scale, stream=None):
        T = P.shape[-1]
        _lib.check("foho_dec_softmax_bwd", self.lib.foho_dec_softmax_bwd(P.data_ptr(), dP.data_ptr(), dS.data_ptr(), P.numel() // T, T,
                                                                         scale, _sp(stream)))
        return dS

    def fourier(self, xyz, out, num_freqs, include_pi, stream=None):
        _lib.check("foho_dec_fourier_embed", self.lib.foho_dec_fourier_embed(xyz.data_ptr(), out.data_ptr(), xyz.shape[0], out.shape[1],
                                                                             num_freqs, int(include_pi), _sp(stream)))
        return out

    def head(self, x, ln_w, ln_b, w_out, b_out, out, idx=None, stream=None):
        _lib.check("foho_dec_head", self.lib.foho_dec_head(x.data_ptr(), x.stride(0), ln_w.data_ptr(), ln_b.data_ptr(), LN_POST_EPS,
                                                           w_out.data_ptr(), b_out, None if idx is None else idx.data_ptr(),
                                                           out.data_ptr(), x.shape[0], _sp(stream)))

    def head_bwd(self, x, ln_w, w_out, dS, g_scale, dx, idx=None, stream=None):
        _lib.check("foho_dec_head_bwd", self.lib.foho_dec_head_bwd(x.data_ptr(), x.stride(0), ln_w.data_ptr(), LN_POST_EPS, w_out.data_ptr(),
                                                                   None if idx is None else idx.data_ptr(), dS.data_ptr(), g_scale,
                                                                   dx.data_ptr(), dx.stride(0), x.shape[0], _sp(stream)))
        return dx

    def gather(self, src, idx, out, stream=None):
        W = src.shape[1]
        _lib.check("foho_dec_gather_rows", self.lib.foho_dec_gather_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), out.data_ptr(),
                                                                         out.stride(0), idx.numel(), W, _sp(stream)))
        return out

    def rowdot(self, a, b, out, heads=HEADS, stream=None):
        """out[h, r] = <a[r, h, :], b[r, h, :]> (float32 [heads, rows] view, e.g. a column slice of a longer table):
        rowsum(dO o O) of the attention backward."""
        if out.dim() != 2 or out.shape != (heads, a.shape[0]) or out.stride(1) != 1:
            raise ValueError("out must be a float32 [heads, rows] view with contiguous rows")
        _lib.check("foho_dec_rowdot", self.lib.foho_dec_rowdot(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(),
                                                               out.stride(0), a.shape[0], heads, _sp(stream)))
        return out

    def gather_f32(self, src, idx, out, stream=None):
        """out[h, i] = src[h, idx[i]] for float32 src [heads, n_src]."""
        _lib.check("foho_dec_gather_f32", self.lib.foho_dec_gather_f32(src.data_ptr(), src.stride(0), idx.data_ptr(), out.data_ptr(),
                                                                       idx.numel(), src.shape[0], _sp(stream)))
        return out

    def cast(self, src, dst, scale=1.0, accumulate=False, stream=None):
        """2-D row-major views; f32 -> f16, f16 -> f32 (optionally accumulating) or f16 -> f16 (scaled copy)."""
        if src.dim() != 2 or dst.shape != src.shape or src.stride(1) != 1 or dst.stride(1) != 1:
            raise ValueError("cast expects matching 2-D views with contiguous columns")
        if src.dtype == torch.float32 and dst.dtype == torch.float16:
            mode = 0
        elif src.dtype == torch.float16 and dst.dtype == torch.float32:
            mode = 2 if accumulate else 1
        elif src.dtype == torch.float16 and dst.dtype == torch.float16 and not accumulate:
            mode = 4
        else:
            raise ValueError("unsupported cast")
        _lib.check("foho_dec_cast", self.lib.foho_dec_cast(src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), src.shape[0],
                                                           src.shape[1], scale, mode, _sp(stream)))
        return dst


class GradCompactor:
    """dense dE/dSDF [B, V] -> (idx int32 [B, cap], val float32 [B, cap], count int32 [B]) on the device
    (``foho_dec_compact_grad``: deterministic order, zero padded, overflow flag)."""

    def __init__(self, B: int, V: int, cap: int, device):
        self.lib = _lib.load()
        self.B, self.V, self.cap = B, V, cap
        self.idx = torch.zeros(B, cap, dtype=torch.int32, device=device)
        self.val = torch.zeros(B, cap, dtype=torch.float32, device=device)
        self.count = torch.zeros(B, dtype=torch.int32, device=device)
        self.flags = torch.zeros(1, dtype=torch.int32, device=device)
        nbytes = self.lib.foho_dec_compact_workspace_bytes(B, V)
        self.ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)

    def __call__(self, g: torch.Tensor, stream=None):
        if g.dtype != torch.float32 or not g.is_contiguous() or g.numel() != self.B * self.V:
            raise ValueError("expected a contiguous float32 [B, V] gradient")
        _lib.check("foho_dec_compact_grad", self.lib.foho_dec_compact_grad(
            g.data_ptr(), self.B, self.V, self.cap, self.idx.data_ptr(), self.val.data_ptr(), self.count.data_ptr(),
            self.flags.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _sp(stream)))
        return self.idx, self.val


class DecoderWeights:
    """Device copies of the decode half of ``ShapeVAE``: fp16 matrices for the tensor cores, fp32 biases and
    LayerNorm parameters.  ``state_dict`` keys are the reference package's (see the module docstring)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, num_layers: Optional[int] = None, num_freqs: int = 8,
                 include_pi: bool = False, scale_factor: float = SCALE_FACTOR):
        sd = state_dict
        self.device = torch.device(device)
        self.num_freqs, self.include_pi, self.scale_factor = num_freqs, include_pi, float(scale_factor)
        if num_layers is None:
            num_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
        self.num_layers = num_layers
        h = lambda k: sd[k].detach().to(self.device, torch.float16).contiguous()
        f = lambda k: sd[k].detach().to(self.device, torch.float32).contiguous()
        opt = lambda k: f(k) if k in sd else None
        self.post_kl_w, self.post_kl_b = h("post_kl.weight"), f("post_kl.bias")
        self.layers: List[dict] = []
        for i in range(num_layers):
            p = f"transformer.resblocks.{i}."
            self.layers.append(dict(
                ln1_w=f(p + "ln_1.weight"), ln1_b=f(p + "ln_1.bias"), qkv_w=h(p + "attn.c_qkv.weight"), qkv_b=opt(p + "attn.c_qkv.bias"),
                proj_w=h(p + "attn.c_proj.weight"), proj_b=f(p + "attn.c_proj.bias"),
                qn_w=opt(p + "attn.attention.q_norm.weight"), qn_b=opt(p + "attn.attention.q_norm.bias"),
                kn_w=opt(p + "attn.attention.k_norm.weight"), kn_b=opt(p + "attn.attention.k_norm.bias"),
                ln2_w=f(p + "ln_2.weight"), ln2_b=f(p + "ln_2.bias"), fc_w=h(p + "mlp.c_fc.weight"), fc_b=f(p + "mlp.c_fc.bias"),
                fc2_w=h(p + "mlp.c_proj.weight"), fc2_b=f(p + "mlp.c_proj.bias")))
            if self.layers[-1]["qn_w"] is None:
                raise ValueError("qk_norm=False checkpoints are not supported by the per-head LayerNorm kernels")
        g = "geo_decoder."
        c = g + "cross_attn_decoder."
        qp = sd[g + "query_proj.weight"].detach().to(self.device, torch.float32)
        self.embed_dim = qp.shape[1]                                   # 3 * (2 * num_freqs + 1) = 51
        self.embed_ld = 64
        qpw = torch.zeros(qp.shape[0], self.embed_ld, dtype=torch.float16, device=self.device)   # K padded to 64 for the tensor core
        qpw[:, :self.embed_dim] = qp.to(torch.float16)
        self.query_proj_w, self.query_proj_b = qpw, f(g + "query_proj.bias")
        self.x = dict(
            ln1_w=f(c + "ln_1.weight"), ln1_b=f(c + "ln_1.bias"), ln2_w=f(c + "ln_2.weight"), ln2_b=f(c + "ln_2.bias"),
            ln3_w=f(c + "ln_3.weight"), ln3_b=f(c + "ln_3.bias"), q_w=h(c + "attn.c_q.weight"), q_b=opt(c + "attn.c_q.bias"),
            kv_w=h(c + "attn.c_kv.weight"), kv_b=opt(c + "attn.c_kv.bias"), proj_w=h(c + "attn.c_proj.weight"),
            proj_b=f(c + "attn.c_proj.bias"), qn_w=f(c + "attn.attention.q_norm.weight"), qn_b=f(c + "attn.attention.q_norm.bias"),
            kn_w=f(c + "attn.attention.k_norm.weight"), kn_b=f(c + "attn.attention.k_norm.bias"),
            fc_w=h(c + "mlp.c_fc.weight"), fc_b=f(c + "mlp.c_fc.bias"), fc2_w=h(c + "mlp.c_proj.weight"), fc2_b=f(c + "mlp.c_proj.bias"))
        self.ln_post_w, self.ln_post_b = f(g + "ln_post.weight"), f(g + "ln_post.bias")
        self.out_w = f(g + "output_proj.weight").reshape(-1).contiguous()
        self.out_b = float(sd[g + "output_proj.bias"].reshape(-1)[0])


class LatentDecoder:
    """``latent2sdf`` for a batch of B images on one lattice of Nq query points, forward and adjoint."""

    def __init__(self, weights: DecoderWeights, B: int, device=None, query_chunk: int = 0, active_chunk: int = 2048,
                 loss_scale: float = 4096.0):
        self.w = weights
        self.B = int(B)
        self.dev = torch.device(device or weights.device)
        self.ops = _Ops()
        self.loss_scale = float(loss_scale)
        self.active_chunk = int(active_chunk)
        self.query_chunk = int(query_chunk) or max(1024, (262144 // self.B) // 128 * 128)     # B=8: 103 ms per forward against 114 at 32768 // B
        R = self.B * TOKENS
        f16 = dict(dtype=torch.float16, device=self.dev)
        L = weights.num_layers
        # token-side activations kept for the adjoint, one set per layer
        self.act = [dict(x_in=torch.empty(R, WIDTH, **f16), qkv=torch.empty(R, 3 * WIDTH, **f16), qn=torch.empty(R, HEADS, HD, **f16),
                         kn=torch.empty(R, HEADS, HD, **f16), x_mid=torch.empty(R, WIDTH, **f16), u_pre=torch.empty(R, 4 * WIDTH, **f16),
                         o=torch.empty(self.B, TOKENS, WIDTH, **f16),
                         lse=torch.empty(self.B, HEADS, TOKENS, dtype=torch.float32, device=self.dev))
                    for _ in range(L)]
        self.lat16 = torch.empty(R, EMBED, **f16)
        self.h = torch.empty(R, WIDTH, **f16)            # LayerNorm output scratch
        self.u = torch.empty(R, 4 * WIDTH, **f16)
        self.data = torch.empty(R, WIDTH, **f16)         # transformer output = the geo decoder's `latents`
        self.kv = torch.empty(R, 2 * WIDTH, **f16)
        self.kvn = torch.empty(R, HEADS, HD, **f16)      # k_norm(k)
        self.x0: Optional[torch.Tensor] = None
        self.qn: Optional[torch.Tensor] = None
        self.Nq = 0
        self._fwd_done = False

    # ------------------------------------------------------------------ lattice (once)
    def set_queries(self, xyz: torch.Tensor) -> None:
        """``xyz`` [Nq, 3] float32 lattice points (``generate_dense_grid_points``, pipelines.py:341-360)."""
        w, ops = self.w, self.ops
        xyz = xyz.to(self.dev, torch.float32).contiguous()
        self.Nq = Nq = xyz.shape[0]
        f16 = dict(dtype=torch.float16, device=self.dev)
        self.x0 = torch.empty(Nq, WIDTH, **f16)
        self.qn = torch.empty(Nq, HEADS, HD, **f16)
        step = 65536
        emb = torch.empty(min(step, Nq), w.embed_ld, **f16)
        t1 = torch.empty(min(step, Nq), WIDTH, **f16)
        t2 = torch.empty(min(step, Nq), WIDTH, **f16)
        for s in range(0, Nq, step):
            n = min(step, Nq - s)
            ops.fourier(xyz[s:s + n], emb[:n], w.num_freqs, w.include_pi)
            tc.gemm(emb[:n], w.query_proj_w, out=self.x0[s:s + n], bias=w.query_proj_b)
            ops.layernorm(self.x0[s:s + n], w.x["ln1_w"], w.x["ln1_b"], t1[:n])
            tc.gemm(t1[:n], w.x["q_w"], out=t2[:n], bias=w.x["q_b"])
            ops.layernorm(t2[:n].view(n, HEADS, HD), w.x["qn_w"], w.x["qn_b"], self.qn[s:s + n], width=HD)
        qc = self.query_chunk = min(self.query_chunk, (Nq + 127) // 128 * 128)      # never larger than the lattice
        self.q_attn = torch.empty(self.B, qc, WIDTH, **f16)
        self.q_x = torch.empty(self.B, qc, WIDTH, **f16)
        self.q_h = torch.empty(self.B, qc, WIDTH, **f16)
        self.q_u = torch.empty(self.B, qc, 4 * WIDTH, **f16)
        self.q_y = torch.empty(self.B, qc, WIDTH, **f16)
        self.q_lse = torch.empty(self.B, HEADS, Nq, dtype=torch.float32, device=self.dev)      # log2-sum-exp of every query row

    # ------------------------------------------------------------------ forward
    def forward(self, latents: torch.Tensor, out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        """``latents`` [B, 3072, 64] float32 or float16 (x1 of ``step_final``; half in the reference, pipelines.py:1204);
        returns sdf [B, Nq] float32, negative inside."""
        w, ops, B = self.w, self.ops, self.B
        if self.x0 is None:
            raise RuntimeError("set_queries() first")
        R = B * TOKENS
        lat = latents.reshape(R, EMBED)
        ops.cast(lat, self.lat16, scale=1.0 / w.scale_factor, stream=stream)            # pred = 1/scale_factor * pred (:297)
        tc.gemm(self.lat16, w.post_kl_w, out=self.act[0]["x_in"] if self.act else self.data, bias=w.post_kl_b, stream=stream)
        for i, (lw, a) in enumerate(zip(w.layers, self.act)):
            ops.layernorm(a["x_in"], lw["ln1_w"], lw["ln1_b"], self.h, stream=stream)
            tc.gemm(self.h, lw["qkv_w"], out=a["qkv"], bias=lw["qkv_b"], stream=stream)
            qkv = a["qkv"].view(R, HEADS, 3 * HD)
            ops.layernorm(qkv[:, :, :HD], lw["qn_w"], lw["qn_b"], a["qn"], width=HD, stream=stream)
            ops.layernorm(qkv[:, :, HD:2 * HD], lw["kn_w"], lw["kn_b"], a["kn"], width=HD, stream=stream)
            tc.attention(a["qn"], a["kn"], qkv[:, :, 2 * HD:], B, out=a["o"], lse2=a["lse"], stream=stream)
            tc.gemm(a["o"].view(R, WIDTH), lw["proj_w"], out=a["x_mid"], bias=lw["proj_b"], res=a["x_in"], stream=stream)
            ops.layernorm(a["x_mid"], lw["ln2_w"], lw["ln2_b"], self.h, stream=stream)
            tc.gemm(self.h, lw["fc_w"], out=self.u, bias=lw["fc_b"], act=tc.ACT_GELU, aux_out=a["u_pre"], stream=stream)
            nxt = self.act[i + 1]["x_in"] if i + 1 < len(self.act) else self.data
            tc.gemm(self.u, lw["fc2_w"], out=nxt, bias=lw["fc2_b"], res=a["x_mid"], stream=stream)
        # geo decoder, token side: k, v of the cross attention
        xw = w.x
        ops.layernorm(self.data, xw["ln2_w"], xw["ln2_b"], self.h, stream=stream)
        tc.gemm(self.h, xw["kv_w"], out=self.kv, bias=xw["kv_b"], stream=stream)
        kv = self.kv.view(R, HEADS, 2 * HD)
        ops.layernorm(kv[:, :, :HD], xw["kn_w"], xw["kn_b"], self.kvn, width=HD, stream=stream)
        # query side, chunk by chunk (per-query work: nothing of width 1024 outlives its chunk)
        if out is None:
            out = torch.empty(B, self.Nq, dtype=torch.float32, device=self.dev)
        qc = self.query_chunk
        for s in range(0, self.Nq, qc):
            n = min(qc, self.Nq - s)
            att = self.q_attn[:, :n]
            tc.attention(self.qn[s:s + n], self.kvn, kv[:, :, HD:], B, out=att, q_shared=True, lse2=self.q_lse[:, :, s:s + n], stream=stream)
            xq = tc.gemm(att, xw["proj_w"], out=self.q_x[:, :n], bias=xw["proj_b"], res=self.x0[s:s + n].unsqueeze(0).expand(B, n, WIDTH),
                         stream=stream)
            if n == self.q_x.shape[1]:          # full chunk: the images' rows are one contiguous block, one launch
                ops.layernorm(xq.reshape(B * n, WIDTH), xw["ln3_w"], xw["ln3_b"], self.q_h.view(B * n, WIDTH), stream=stream)
            else:
                for b in range(B):
                    ops.layernorm(xq[b], xw["ln3_w"], xw["ln3_b"], self.q_h[b, :n], stream=stream)
            tc.gemm(self.q_h[:, :n], xw["fc_w"], out=self.q_u[:, :n], bias=xw["fc_b"], act=tc.ACT_GELU, stream=stream)
            tc.gemm(self.q_u[:, :n], xw["fc2_w"], out=self.q_y[:, :n], bias=xw["fc2_b"], res=xq, stream=stream)
            for b in range(B):
                ops.head(self.q_y[b, :n], w.ln_post_w, w.ln_post_b, w.out_w, w.out_b, out[b, s:s + n], stream=stream)
        self._fwd_done = True
        return out

    # ------------------------------------------------------------------ export lattice (forward only, query side on the fly)
    def decode_lattice(self, latents: torch.Tensor, D: int, bound: float = 1.10, chunk: int = 0, stream=None) -> torch.Tensor:
        """``latent2sdf`` on ANOTHER lattice than the resident one -- the reference's final export re-grids to 385^3
        (pipelines.py:1624-1641; 57 M queries: their latent-independent query side would take 233 GB, so it is
        computed chunk by chunk here instead of once).  Runs the token side, then streams the lattice; returns
        float32 [B, D, D, D], negative inside.  Forward only."""
        w, ops, B = self.w, self.ops, self.B
        xw = w.x
        R = B * TOKENS
        if self.x0 is None:
            raise RuntimeError("set_queries() first (any lattice): the token side shares its buffers")
        saved = (self.Nq, self.x0, self.qn, self.q_lse)
        try:
            # token side exactly as in forward(): reuse it by decoding zero queries of the resident lattice
            self.Nq = 0
            self.forward(latents, out=torch.empty(B, 0, dtype=torch.float32, device=self.dev), stream=stream)
        finally:
            self.Nq, self.x0, self.qn, self.q_lse = saved
        kv = self.kv.view(R, HEADS, 2 * HD)
        qc = int(chunk) or self.query_chunk
        qc = min(qc, self.query_chunk)
        f16 = dict(dtype=torch.float16, device=self.dev)
        axis = torch.linspace(-bound, bound, D, device=self.dev)
        out = torch.empty(B, D * D * D, dtype=torch.float32, device=self.dev)
        emb = torch.empty(qc, w.embed_ld, **f16)
        x0 = torch.empty(qc, WIDTH, **f16)
        t1 = torch.empty(qc, WIDTH, **f16)
        t2 = torch.empty(qc, WIDTH, **f16)
        qn = torch.empty(qc, HEADS, HD, **f16)
        N = D * D * D
        for s in range(0, N, qc):
            n = min(qc, N - s)
            idx = torch.arange(s, s + n, device=self.dev)                       # lattice points of this chunk, `ij` order, z fastest
            xyz = torch.stack([axis[idx // (D * D)], axis[(idx // D) % D], axis[idx % D]], -1).contiguous()
            ops.fourier(xyz, emb[:n], w.num_freqs, w.include_pi, stream=stream)
            tc.gemm(emb[:n], w.query_proj_w, out=x0[:n], bias=w.query_proj_b, stream=stream)
            ops.layernorm(x0[:n], xw["ln1_w"], xw["ln1_b"], t1[:n], stream=stream)
            tc.gemm(t1[:n], xw["q_w"], out=t2[:n], bias=xw["q_b"], stream=stream)
            ops.layernorm(t2[:n].view(n, HEADS, HD), xw["qn_w"], xw["qn_b"], qn[:n], width=HD, stream=stream)
            att = self.q_attn[:, :n]
            tc.attention(qn[:n], self.kvn, kv[:, :, HD:], B, out=att, q_shared=True, stream=stream)
            xq = tc.gemm(att, xw["proj_w"], out=self.q_x[:, :n], bias=xw["proj_b"], res=x0[:n].unsqueeze(0).expand(B, n, WIDTH), stream=stream)
            if n == self.q_x.shape[1]:          # full chunk: the images' rows are one contiguous block, one launch
                ops.layernorm(xq.reshape(B * n, WIDTH), xw["ln3_w"], xw["ln3_b"], self.q_h.view(B * n, WIDTH), stream=stream)
            else:
                for b in range(B):
                    ops.layernorm(xq[b], xw["ln3_w"], xw["ln3_b"], self.q_h[b, :n], stream=stream)
            tc.gemm(self.q_h[:, :n], xw["fc_w"], out=self.q_u[:, :n], bias=xw["fc_b"], act=tc.ACT_GELU, stream=stream)
            tc.gemm(self.q_u[:, :n], xw["fc2_w"], out=self.q_y[:, :n], bias=xw["fc2_b"], res=xq, stream=stream)
            for b in range(B):
                ops.head(self.q_y[b, :n], w.ln_post_w, w.ln_post_b, w.out_w, w.out_b, out[b, s:s + n], stream=stream)
        return out.view(B, D, D, D)

    # ------------------------------------------------------------------ adjoint
    def backward(self, idx: torch.Tensor, g_sdf: torch.Tensor, out: Optional[torch.Tensor] = None, stream=None,
                 out_scale: float = 1.0) -> torch.Tensor:
        """dE/d(latents) [B, 3072, 64] float32 from dE/dSDF on a sparse set of lattice points.

        ``idx`` [B, M] int32 lattice indices of the rows that carry a gradient, ``g_sdf`` [B, M] float32 their
        dE/dSDF (pad with index 0 / gradient 0).  Uses the token-side activations of the last ``forward``.
        ``out_scale`` multiplies the result (the ``1 - sigma`` of ``x1 = x_t + (1 - sigma) v``, schedulers.py:481,
        turns dE/dx1 into dE/dv)."""
        if not self._fwd_done:
            raise RuntimeError("forward() first: the adjoint reuses its token-side activations")
        w, ops, B = self.w, self.ops, self.B
        xw = w.x
        R = B * TOKENS
        M = idx.shape[1]
        if M % 8 or self.active_chunk % 8 or idx.dtype != torch.int32 or g_sdf.dtype != torch.float32:
            raise ValueError("idx must be int32 [B, M] and g_sdf float32 [B, M] with M a multiple of 8")
        f16 = dict(dtype=torch.float16, device=self.dev)
        f32 = dict(dtype=torch.float32, device=self.dev)
        ls = self.loss_scale
        if not hasattr(self, "_bw"):
            mc = self.active_chunk
            self._bw = dict(
                x0=torch.empty(mc, WIDTH, **f16), a=torch.empty(mc, WIDTH, **f16),
                x=torch.empty(mc, WIDTH, **f16), h=torch.empty(mc, WIDTH, **f16), u=torch.empty(mc, 4 * WIDTH, **f16),
                u_pre=torch.empty(mc, 4 * WIDTH, **f16), y=torch.empty(mc, WIDTH, **f16), dy=torch.empty(mc, WIDTH, **f16),
                du=torch.empty(mc, 4 * WIDTH, **f16), dh=torch.empty(mc, WIDTH, **f16), dx=torch.empty(mc, WIDTH, **f16),
                dkn16=torch.empty(R, HEADS, HD, **f16), dkv=torch.empty(R, 2 * WIDTH, **f16),
                g=torch.empty(R, WIDTH, **f16), g2=torch.empty(R, WIDTH, **f16), g3=torch.empty(R, WIDTH, **f16),
                gu=torch.empty(R, 4 * WIDTH, **f16), dqkv=torch.empty(R, 3 * WIDTH, **f16),
                tdelta=torch.empty(B, HEADS, TOKENS, **f32), dqn=torch.empty(R, HEADS, HD, **f16),
                dknl=torch.empty(R, HEADS, HD, **f16))
        bw = self._bw
        mc = self.active_chunk
        if getattr(self, "_bwq_M", None) != M:
            # everything the fused attention adjoint needs of the M active lattice points of every image, kept over the chunks
            self._bwq = dict(qa=torch.empty(B * M, HEADS, HD, **f16), da=torch.empty(B * M, WIDTH, **f16),
                             lse=torch.empty(B, HEADS, M, **f32), delta=torch.empty(B, HEADS, M, **f32))
            self._bwq_M = M
        bq = self._bwq
        kvv = self.kv.view(B, TOKENS, HEADS, 2 * HD)
        kvn = self.kvn.view(B, TOKENS, HEADS, HD)
        # ---- query side: recompute the rows that carry a gradient (flash attention forward, then the MLP block), walk
        #      back to dA; the attention adjoint of all of them runs once at the end (dK, dV only: the lattice queries
        #      do not depend on the latents)
        for b in range(B):
            for s in range(0, M, mc):
                n = min(mc, M - s)
                ix = idx[b, s:s + n]
                r0 = b * M + s
                qn_a = ops.gather(self.qn.view(self.Nq, WIDTH), ix, bq["qa"].view(B * M, WIDTH)[r0:r0 + n], stream=stream).view(n, HEADS, HD)
                x0_a = ops.gather(self.x0, ix, bw["x0"][:n], stream=stream)
                a = bw["a"][:n]
                tc.attention(qn_a, kvn[b], kvv[b, :, :, HD:], 1, out=a.view(1, n, WIDTH), lse2=bq["lse"][b:b + 1, :, s:s + n], stream=stream)
                x = tc.gemm(a, xw["proj_w"], out=bw["x"][:n], bias=xw["proj_b"], res=x0_a, stream=stream)
                ops.layernorm(x, xw["ln3_w"], xw["ln3_b"], bw["h"][:n], stream=stream)
                tc.gemm(bw["h"][:n], xw["fc_w"], out=bw["u"][:n], bias=xw["fc_b"], act=tc.ACT_GELU, aux_out=bw["u_pre"][:n], stream=stream)
                y = tc.gemm(bw["u"][:n], xw["fc2_w"], out=bw["y"][:n], bias=xw["fc2_b"], res=x, stream=stream)
                # backward of the head, the MLP block, c_proj
                dy = ops.head_bwd(y, w.ln_post_w, w.out_w, g_sdf[b, s:s + n], ls, bw["dy"][:n], stream=stream)
                tc.gemm(dy, xw["fc2_w"], out=bw["du"][:n], b_mn=True, act=tc.ACT_DGELU, aux_in=bw["u_pre"][:n], stream=stream)
                tc.gemm(bw["du"][:n], xw["fc_w"], out=bw["dh"][:n], b_mn=True, stream=stream)
                dx = ops.layernorm_bwd(x, xw["ln3_w"], bw["dh"][:n], bw["dx"][:n], add=dy, stream=stream)
                da = tc.gemm(dx, xw["proj_w"], out=bq["da"][r0:r0 + n], b_mn=True, stream=stream)
                ops.rowdot(da, a, bq["delta"][b][:, s:s + n], stream=stream)
        # ---- attention adjoint (fused, k_attn_bwd): dV into the v half of dkv, dKn through the per-head LayerNorm adjoint
        #      into the k half
        dkv = bw["dkv"].view(R, HEADS, 2 * HD)
        kv_r = self.kv.view(R, HEADS, 2 * HD)
        tc.attention_bwd(bq["qa"], self.kvn.view(R, HEADS, HD), kv_r[:, :, HD:], bq["da"].view(B * M, HEADS, HD), bq["lse"], bq["delta"],
                         None, bw["dkn16"], dkv[:, :, HD:], B, stream=stream)
        ops.layernorm_bwd(kv_r[:, :, :HD], xw["kn_w"], bw["dkn16"], dkv[:, :, :HD], width=HD, stream=stream)
        tc.gemm(bw["dkv"], xw["kv_w"], out=bw["g2"], b_mn=True, stream=stream)
        g = ops.layernorm_bwd(self.data, xw["ln2_w"], bw["g2"], bw["g"], stream=stream)     # d(data)
        # ---- token transformer, last layer first
        for i in range(len(w.layers) - 1, -1, -1):
            lw, a = w.layers[i], self.act[i]
            qkv = a["qkv"].view(B, TOKENS, HEADS, 3 * HD)
            dqkv = bw["dqkv"].view(B, TOKENS, HEADS, 3 * HD)
            tc.gemm(g, lw["fc2_w"], out=bw["gu"], b_mn=True, act=tc.ACT_DGELU, aux_in=a["u_pre"], stream=stream)
            tc.gemm(bw["gu"], lw["fc_w"], out=bw["g2"], b_mn=True, stream=stream)
            g_mid = ops.layernorm_bwd(a["x_mid"], lw["ln2_w"], bw["g2"], bw["g3"], add=g, stream=stream)
            da = tc.gemm(g_mid, lw["proj_w"], out=bw["g2"], b_mn=True, stream=stream).view(B, TOKENS, HEADS, HD)
            # attention adjoint, fused (k_attn_bwd): P and dS live in shared memory tiles only; dV goes straight into the
            # v third of dqkv, dQn / dKn through the per-head LayerNorm adjoints into the q and k thirds
            for b in range(B):
                ops.rowdot(da[b].reshape(TOKENS, WIDTH), a["o"][b], bw["tdelta"][b], stream=stream)
            qkv_r, dqkv_r = a["qkv"].view(R, HEADS, 3 * HD), bw["dqkv"].view(R, HEADS, 3 * HD)
            tc.attention_bwd(a["qn"].view(R, HEADS, HD), a["kn"].view(R, HEADS, HD), qkv_r[:, :, 2 * HD:], da.view(R, HEADS, HD),
                             a["lse"], bw["tdelta"], bw["dqn"], bw["dknl"], dqkv_r[:, :, 2 * HD:], B, stream=stream)
            ops.layernorm_bwd(qkv_r[:, :, :HD], lw["qn_w"], bw["dqn"], dqkv_r[:, :, :HD], width=HD, stream=stream)
            ops.layernorm_bwd(qkv_r[:, :, HD:2 * HD], lw["kn_w"], bw["dknl"], dqkv_r[:, :, HD:2 * HD], width=HD, stream=stream)
            tc.gemm(bw["dqkv"], lw["qkv_w"], out=bw["g2"], b_mn=True, stream=stream)
            g = ops.layernorm_bwd(a["x_in"], lw["ln1_w"], bw["g2"], bw["g"], add=g_mid, stream=stream)
        # ---- post_kl and the 1/scale_factor of the call site
        if out is None:
            out = torch.empty(B, TOKENS, EMBED, **f32)
        # float32 (or float16) out; the epilogue undoes the loss scale and applies the call site's 1/scale_factor (:297)
        tc.gemm(g, w.post_kl_w, out=out.view(R, EMBED), b_mn=True, alpha=out_scale / (w.scale_factor * ls), stream=stream)
        return out


def random_state_dict(num_layers: int = 16, seed: int = 0, device="cpu", num_freqs: int = 8) -> Dict[str, torch.Tensor]:
    """Random-init weights with the names and shapes of the decode half of Hunyuan3D-2's ShapeVAE (3072 x 64 latents,
    width 1024, 16 heads, qk_norm, no qkv bias) -- for benchmarks and synthetic runs: there is no network for the
    released checkpoint.  nn.Linear-style uniform init, LayerNorm weights around 1."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = torch.device(device)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out_f, in_f, bias=True):
        b = 1.0 / in_f ** 0.5
        sd[name + ".weight"] = ((torch.rand(out_f, in_f, generator=g) * 2 - 1) * b).to(dev)
        if bias:
            sd[name + ".bias"] = ((torch.rand(out_f, generator=g) * 2 - 1) * b).to(dev)

    def ln(name, n):
        sd[name + ".weight"] = (1.0 + 0.1 * torch.randn(n, generator=g)).to(dev)
        sd[name + ".bias"] = (0.02 * torch.randn(n, generator=g)).to(dev)

    lin("post_kl", WIDTH, EMBED)
    for i in range(num_layers):
        p = f"transformer.resblocks.{i}."
        ln(p + "ln_1", WIDTH); ln(p + "ln_2", WIDTH)
        lin(p + "attn.c_qkv", 3 * WIDTH, WIDTH, bias=False); lin(p + "attn.c_proj", WIDTH, WIDTH)
        ln(p + "attn.attention.q_norm", HD); ln(p + "attn.attention.k_norm", HD)
        lin(p + "mlp.c_fc", 4 * WIDTH, WIDTH); lin(p + "mlp.c_proj", WIDTH, 4 * WIDTH)
    gd = "geo_decoder."
    c = gd + "cross_attn_decoder."
    lin(gd + "query_proj", WIDTH, 3 * (2 * num_freqs + 1))
    for n in ("ln_1", "ln_2", "ln_3"):
        ln(c + n, WIDTH)
    lin(c + "attn.c_q", WIDTH, WIDTH, bias=False); lin(c + "attn.c_kv", 2 * WIDTH, WIDTH, bias=False); lin(c + "attn.c_proj", WIDTH, WIDTH)
    ln(c + "attn.attention.q_norm", HD); ln(c + "attn.attention.k_norm", HD)
    lin(c + "mlp.c_fc", 4 * WIDTH, WIDTH); lin(c + "mlp.c_proj", WIDTH, 4 * WIDTH)
    ln(gd + "ln_post", WIDTH)
    lin(gd + "output_proj", 1, WIDTH)
    return sd


def lattice_points(D: int, bound: float = 1.10) -> torch.Tensor:
    """``generate_dense_grid_points`` (pipelines.py:341-360): linspace(-bound, bound, D)^3, ``ij`` order, z fastest."""
    axis = torch.linspace(-bound, bound, D)
    return torch.stack(torch.meshgrid(axis, axis, axis, indexing="ij"), -1).reshape(-1, 3)


def decode_flops(n_queries: int, num_layers: int = 16) -> Dict[str, float]:
    """Multiply-add FLOPs (2 per MAC) of one ``latent2sdf`` of one image (the accounting of DESIGN.md section 8)."""
    T, W = TOKENS, WIDTH
    layer = 2 * T * 12 * W * W + 4 * T * T * W
    transformer = num_layers * layer + 2 * T * EMBED * W
    kv = 2 * T * W * 2 * W
    attn = 4 * n_queries * T * W
    post = 2 * n_queries * (W * W + 8 * W * W + W)
    return {"transformer": transformer, "kv": kv, "cross_attention": attn, "per_query_mlp": post, "forward": transformer + kv + attn + post}


def adjoint_flops(n_active: int, num_layers: int = 16) -> float:
    """FLOPs of ``LatentDecoder.backward`` for one image with ``n_active`` rows carrying a gradient: input gradients
    only (weights are frozen): every linear layer once more, attention backward = recomputed scores + four products,
    plus the recompute of the active query rows."""
    T, W, M = TOKENS, WIDTH, n_active
    token = num_layers * (2 * T * 12 * W * W + 10 * T * T * W) + 2 * T * W * 2 * W + 2 * T * EMBED * W
    query = 4 * M * T * W + 2 * M * 9 * W * W            # recompute: attention + c_proj + MLP
    query += 2 * M * 9 * W * W + 6 * M * T * W           # backward: MLP + c_proj, then dP, dV, dKn
    return float(token + query)
