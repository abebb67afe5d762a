"""Thin torch-tensor front of the tensor-core building blocks of the C-ABI (``foho_tc_gemm`` ...).

Nothing here computes: it only fills descriptors with raw device pointers.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib

ACT_NONE, ACT_GELU, ACT_DGELU, ACT_EXP2_ROW, ACT_DSOFTMAX = 0, 1, 2, 3, 4


def _stream_ptr(stream: Optional[torch.cuda.Stream]) -> C.c_void_p:
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _rows(t: torch.Tensor):
    """(batch, rows, cols, ld, batch_stride) of a 2-D or 3-D tensor whose last dim is contiguous."""
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.stride(2) != 1:
        raise ValueError("operand must be [rows, cols] or [batch, rows, cols] with a contiguous last dimension")
    return t.shape[0], t.shape[1], t.shape[2], t.stride(1), (t.stride(0) if t.shape[0] > 1 else 0)


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *, a_mn: bool = False, b_mn: bool = False,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, res: Optional[torch.Tensor] = None,
         aux_in: Optional[torch.Tensor] = None, aux_out: Optional[torch.Tensor] = None, alpha: float = 1.0,
         out_dtype: torch.dtype = torch.float16, block_n: int = 0, max_ctas: int = 0,
         row_vec: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """``out[b] = res[b] + act(alpha * A[b] @ B[b]^T + bias)`` on the tensor cores (fp16 operands, fp32 accumulate).

    ``a``: [.., M, K] (or [.., K, M] when ``a_mn``); ``b``: [.., N, K] -- the ``nn.Linear`` weight layout -- (or
    [.., K, N] when ``b_mn``).  Batched operands may be strided views (e.g. one attention head of a
    [tokens, heads*64] tensor) as long as the last dimension is contiguous."""
    lib = _lib.load()
    if a.dtype != torch.float16 or b.dtype != torch.float16 or not a.is_cuda:
        raise ValueError("tensor-core GEMM operands must be CUDA float16 tensors")
    ba, ra, ca, lda, bsa = _rows(a)
    bb, rb, cb, ldb, bsb = _rows(b)
    M, K = (ca, ra) if a_mn else (ra, ca)
    N, Kb = (cb, rb) if b_mn else (rb, cb)
    if K != Kb:
        raise ValueError(f"inner dimensions differ: {K} vs {Kb}")
    batch = max(ba, bb)
    if (ba not in (1, batch)) or (bb not in (1, batch)):
        raise ValueError("batch sizes differ")
    if out is None:
        out = torch.empty((batch, M, N) if (a.dim() == 3 or b.dim() == 3) else (M, N), dtype=out_dtype, device=a.device)
    bo, ro, co, ldc, bsc = _rows(out)
    if (ro, co) != (M, N) or bo != batch:
        raise ValueError("output shape mismatch")
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.c_f32 = int(out.dtype == torch.float32)
    d.act, d.block_n, d.max_ctas, d.alpha = act, block_n, max_ctas, alpha
    d.A, d.lda, d.bsa = a.data_ptr(), lda, bsa
    d.B, d.ldb, d.bsb = b.data_ptr(), ldb, bsb
    d.C, d.ldc, d.bsc = out.data_ptr(), ldc, bsc
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N:
            raise ValueError("bias must be float32 [N]")
        d.bias = bias.data_ptr()
    if res is not None:
        br, rr, cr, ldr, bsr = _rows(res)
        if (rr, cr) != (M, N):
            raise ValueError("residual shape mismatch")
        d.res, d.ldr, d.bsr, d.res_f32 = res.data_ptr(), ldr, bsr, int(res.dtype == torch.float32)
    for t in (aux_in, aux_out):
        if t is not None:
            bx, rx, cx, ldx, bsx = _rows(t)
            if (rx, cx) != (M, N) or t.dtype != torch.float16:
                raise ValueError("aux tensors must be float16 [.., M, N]")
            d.ldaux, d.bsaux = ldx, bsx
    if row_vec is not None:
        if row_vec.dtype != torch.float32 or row_vec.stride(-1) != 1 or row_vec.shape[-1] != M:
            raise ValueError("row_vec must be float32 [.., M]")
        d.row_vec, d.bs_rowvec = row_vec.data_ptr(), (row_vec.stride(0) if row_vec.dim() == 2 and row_vec.shape[0] > 1 else 0)
    if aux_in is not None:
        d.aux_in = aux_in.data_ptr()
    if aux_out is not None:
        d.aux_out = aux_out.data_ptr()
    _lib.check("foho_tc_gemm", lib.foho_tc_gemm(C.byref(d), _stream_ptr(stream)))
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_img: int, out: Optional[torch.Tensor] = None, *,
              q_shared: bool = False, scale: float = 0.125, max_ctas: int = 0, variant: int = 0,
              lse2: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """``out[i, q, h, :] = softmax_k(scale * Q[q, h] . K[i, k, h]) V[i, k, h]`` on the tensor cores.

    ``q``: [n_q (shared) or n_img*n_q, heads, 64]; ``k``, ``v``: [n_img*n_k, heads, 64] -- strided views of the
    fused projections are fine (last dimension contiguous).  Returns fp16 [n_img, n_q, heads*64]."""
    lib = _lib.load()
    for t in (q, k, v):
        if t.dtype != torch.float16 or not t.is_cuda or t.dim() != 3 or t.shape[2] != 64 or t.stride(2) != 1:
            raise ValueError("q, k, v must be CUDA float16 [rows, heads, 64] views with a contiguous last dimension")
    heads = q.shape[1]
    n_q = q.shape[0] if q_shared else q.shape[0] // n_img
    n_k = k.shape[0] // n_img
    if out is None:
        out = torch.empty(n_img, n_q, heads * 64, dtype=torch.float16, device=q.device)
    d = _lib.AttnDesc()
    d.n_img, d.heads, d.n_q, d.n_k = n_img, heads, n_q, n_k
    d.q_shared, d.max_ctas, d.scale, d.variant = int(q_shared), max_ctas, scale, variant
    d.q, d.ldq, d.hsq = q.data_ptr(), q.stride(0), q.stride(1)
    d.k, d.ldk, d.hsk = k.data_ptr(), k.stride(0), k.stride(1)
    d.v, d.ldv, d.hsv = v.data_ptr(), v.stride(0), v.stride(1)
    d.out, d.ldo, d.out_img_stride = out.data_ptr(), out.stride(1), out.stride(0)
    if lse2 is not None:
        # [n_img, heads, n_q] view of a (possibly longer) [n_img, heads, N] table: row stride = N
        if lse2.dtype != torch.float32 or lse2.shape != (n_img, heads, n_q) or lse2.stride(2) != 1 or \
                (n_img > 1 and lse2.stride(0) != heads * lse2.stride(1)):
            raise ValueError("lse2 must be a float32 [n_img, heads, n_q] view with uniform head stride")
        d.lse2, d.lse2_stride = lse2.data_ptr(), lse2.stride(1)
    _lib.check("foho_tc_attention", lib.foho_tc_attention(C.byref(d), _stream_ptr(stream)))
    return out


def attention_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, d_out: torch.Tensor, lse2: torch.Tensor, delta: torch.Tensor,
                  dq: Optional[torch.Tensor], dk: torch.Tensor, dv: torch.Tensor, n_img: int, *, scale: float = 0.125, max_ctas: int = 0,
                  stream: Optional[torch.cuda.Stream] = None) -> None:
    """The adjoint of ``attention`` without a score matrix in memory (``foho_tc_attention_bwd``).

    ``q``, ``d_out``, ``dq``: [n_img*n_q, heads, 64]; ``k``, ``v``, ``dk``, ``dv``: [n_img*n_k, heads, 64] -- float16 views with a
    contiguous last dimension; ``lse2``, ``delta``: float32 [n_img, heads, n_q] (``lse2`` as ``attention`` wrote it,
    ``delta[i, h, q] = d_out[q, h] . out[q, h]``).  ``dq=None`` skips the query gradient (cross attention of the lattice)."""
    lib = _lib.load()
    for t in (q, k, v, d_out, dq, dk, dv):
        if t is None:
            continue
        if t.dtype != torch.float16 or not t.is_cuda or t.dim() != 3 or t.shape[2] != 64 or t.stride(2) != 1:
            raise ValueError("q, k, v, d_out, dq, dk, dv must be CUDA float16 [rows, heads, 64] views with a contiguous last dimension")
    heads = q.shape[1]
    n_q, n_k = q.shape[0] // n_img, k.shape[0] // n_img
    for t in (lse2, delta):
        if t.dtype != torch.float32 or t.shape != (n_img, heads, n_q) or t.stride(2) != 1 or \
                (n_img > 1 and t.stride(0) != heads * t.stride(1)):
            raise ValueError("lse2 / delta must be float32 [n_img, heads, n_q] views with uniform head stride")
    d = _lib.AttnBwdDesc()
    d.n_img, d.heads, d.n_q, d.n_k, d.max_ctas, d.scale = n_img, heads, n_q, n_k, max_ctas, scale
    d.q, d.ldq, d.hsq = q.data_ptr(), q.stride(0), q.stride(1)
    d.k, d.ldk, d.hsk = k.data_ptr(), k.stride(0), k.stride(1)
    d.v, d.ldv, d.hsv = v.data_ptr(), v.stride(0), v.stride(1)
    d.d_out, d.lddo, d.hsdo = d_out.data_ptr(), d_out.stride(0), d_out.stride(1)
    d.lse2, d.lse2_stride = lse2.data_ptr(), lse2.stride(1)
    d.delta, d.delta_stride = delta.data_ptr(), delta.stride(1)
    if dq is not None:
        d.dq, d.lddq, d.hsdq = dq.data_ptr(), dq.stride(0), dq.stride(1)
    d.dk, d.lddk, d.hsdk = dk.data_ptr(), dk.stride(0), dk.stride(1)
    d.dv, d.lddv, d.hsdv = dv.data_ptr(), dv.stride(0), dv.stride(1)
    _lib.check("foho_tc_attention_bwd", lib.foho_tc_attention_bwd(C.byref(d), _stream_ptr(stream)))
