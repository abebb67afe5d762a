// Row-wise kernels around the tensor-core GEMM / attention of the latent -> SDF decoder (row f1,
// third_party_patches/hy3dgen/shapegen/pipelines.py:292-312): LayerNorm forward / backward (width 1024 and the
// per-head q_norm / k_norm of width 64), row softmax forward / backward for the adjoint's materialised attention,
// Fourier query embedding, the ln_post + output_proj head and its backward, row gather, casts.  All HBM-bound,
// one warp per row, 16-byte accesses, fp32 arithmetic on fp16 activations.
#include "foho_common.cuh"
#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// row r of a [outer, inner, W] view: base + (r / inner) * ld_outer + (r % inner) * ld_inner
__device__ __forceinline__ long long row_off(long long r, int inner, long long ld_outer, long long ld_inner) {
  return (r / inner) * ld_outer + (r % inner) * ld_inner;
}

// ---------------------------------------------------------------- LayerNorm forward, W = 32 * VPL * 8 / ... generic
// One warp per row; W in {64, 1024}.  W=1024: each lane owns 4 chunks of 8 halves (chunk = lane + 32 c); W=64: lanes
// 0..7 own one chunk each (the other lanes idle; rows of 64 are tiny and this kernel is bandwidth-trivial).
template <int W>
__global__ void k_ln_fwd(const __half *__restrict__ x, int inner_x, long long ldo_x, long long ldi_x, const float *__restrict__ w,
                         const float *__restrict__ b, float eps, __half *__restrict__ y, int inner_y, long long ldo_y, long long ldi_y,
                         long long rows) {
  constexpr int CH = W / 8;                       // 16-byte chunks per row
  constexpr int CPL = (CH + 31) / 32;             // chunks per lane
  const int lane = threadIdx.x & 31;
  // W = 1024: the warps are persistent (grid-stride over the rows) and keep this lane's 32 weights and biases in registers
  // -- read per row, the 64 scalar parameter loads per lane were most of the kernel's memory instructions
  float wr[W == 1024 ? CPL : 1][8], br[W == 1024 ? CPL : 1][8];
  if (W == 1024) {
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        wr[c][t] = w ? __ldg(w + (lane + 32 * c) * 8 + t) : 1.f;
        br[c][t] = b ? __ldg(b + (lane + 32 * c) * 8 + t) : 0.f;
      }
  }
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
  const __half *xp = x + row_off(r, inner_x, ldo_x, ldi_x);
  float v[CPL][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ch = lane + 32 * c;
    if (ch < CH) {
      uint4 u = *reinterpret_cast<const uint4 *>(xp + ch * 8);
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[c][2 * t] = f.x; v[c][2 * t + 1] = f.y; s += f.x + f.y; }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) v[c][t] = 0.f;
    }
  }
  const float mean = warp_sum(s) * (1.f / W);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
    if (lane + 32 * c < CH)
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float d = v[c][t] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / W) + eps);
  __half *yp = y + row_off(r, inner_y, ldo_y, ldi_y);
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ch = lane + 32 * c;
    if (ch < CH) {
      __align__(16) __half2 o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float a = (v[c][2 * t] - mean) * rstd, bb = (v[c][2 * t + 1] - mean) * rstd;
        if (W == 1024) {
          a = a * wr[c][2 * t] + br[c][2 * t]; bb = bb * wr[c][2 * t + 1] + br[c][2 * t + 1];
        } else {
          if (w) { a *= __ldg(w + ch * 8 + 2 * t); bb *= __ldg(w + ch * 8 + 2 * t + 1); }
          if (b) { a += __ldg(b + ch * 8 + 2 * t); bb += __ldg(b + ch * 8 + 2 * t + 1); }
        }
        o[t] = __floats2half2_rn(a, bb);
      }
      *reinterpret_cast<uint4 *>(yp + ch * 8) = *reinterpret_cast<uint4 *>(o);
    }
  }
  }
}

// LayerNorm backward w.r.t. the input (weights are frozen): dx = rstd * (g - mean(g) - xhat * mean(g xhat)), g = dy * w;
// optional `add` (same layout as dx) is the gradient already flowing on the residual stream.
template <int W>
__global__ void k_ln_bwd(const __half *__restrict__ x, int inner_x, long long ldo_x, long long ldi_x, const float *__restrict__ w, float eps,
                         const __half *__restrict__ dy, int inner_dy, long long ldo_dy, long long ldi_dy, const __half *__restrict__ add,
                         __half *__restrict__ dx, int inner_dx, long long ldo_dx, long long ldi_dx, long long rows) {
  constexpr int CH = W / 8;
  constexpr int CPL = (CH + 31) / 32;
  const int lane = threadIdx.x & 31;
  float wr[CPL][8];                               // persistent warps: this lane's weights stay in registers
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int t = 0; t < 8; ++t) wr[c][t] = (w && lane + 32 * c < CH) ? __ldg(w + (lane + 32 * c) * 8 + t) : 1.f;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
  const __half *xp = x + row_off(r, inner_x, ldo_x, ldi_x);
  const __half *gp = dy + row_off(r, inner_dy, ldo_dy, ldi_dy);
  float v[CPL][8], g[CPL][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ch = lane + 32 * c;
    if (ch < CH) {
      uint4 u = *reinterpret_cast<const uint4 *>(xp + ch * 8);
      uint4 ug = *reinterpret_cast<const uint4 *>(gp + ch * 8);
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
      const __half2 *hg = reinterpret_cast<const __half2 *>(&ug);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __half22float2(h[t]); v[c][2 * t] = f.x; v[c][2 * t + 1] = f.y; s += f.x + f.y;
        float2 fg = __half22float2(hg[t]);
        g[c][2 * t] = fg.x * wr[c][2 * t];
        g[c][2 * t + 1] = fg.y * wr[c][2 * t + 1];
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) { v[c][t] = 0.f; g[c][t] = 0.f; }
    }
  }
  const float mean = warp_sum(s) * (1.f / W);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
    if (lane + 32 * c < CH)
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float d = v[c][t] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / W) + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
    if (lane + 32 * c < CH)
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float xh = (v[c][t] - mean) * rstd; sg += g[c][t]; sgx += g[c][t] * xh; }
  sg = warp_sum(sg) * (1.f / W);
  sgx = warp_sum(sgx) * (1.f / W);
  const long long off_dx = row_off(r, inner_dx, ldo_dx, ldi_dx);
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int ch = lane + 32 * c;
    if (ch < CH) {
      float o[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float xh = (v[c][t] - mean) * rstd; o[t] = rstd * (g[c][t] - sg - xh * sgx); }
      if (add) {
        uint4 ua = *reinterpret_cast<const uint4 *>(add + off_dx + ch * 8);
        const __half2 *ha = reinterpret_cast<const __half2 *>(&ua);
#pragma unroll
        for (int t = 0; t < 4; ++t) { float2 f = __half22float2(ha[t]); o[2 * t] += f.x; o[2 * t + 1] += f.y; }
      }
      __align__(16) __half2 oh[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) oh[t] = __floats2half2_rn(o[2 * t], o[2 * t + 1]);
      *reinterpret_cast<uint4 *>(dx + off_dx + ch * 8) = *reinterpret_cast<uint4 *>(oh);
    }
  }
  }
}

// Width 64 (the per-head q_norm / k_norm): eight lanes per row, four rows per warp, persistent warps with this lane's eight
// weights (and biases) in registers.  (The generic kernel above leaves 24 of 32 lanes idle at this width and re-reads the
// parameters per row: 4x the warps and 16 scalar loads per lane and row.)
__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__global__ void k_ln64_fwd(const __half *__restrict__ x, int inner_x, long long ldo_x, long long ldi_x, const float *__restrict__ w,
                           const float *__restrict__ b, float eps, __half *__restrict__ y, int inner_y, long long ldo_y, long long ldi_y,
                           long long rows) {
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  float wr[8], br[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { wr[t] = w ? __ldg(w + sub * 8 + t) : 1.f; br[t] = b ? __ldg(b + sub * 8 + t) : 0.f; }
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4; base < rows; base += warps * 4) {
    const long long r = base + grp;
    const bool live = r < rows;
    float v[8];
    float s = 0.f;
    if (live) {
      uint4 u = *reinterpret_cast<const uint4 *>(x + row_off(r, inner_x, ldo_x, ldi_x) + sub * 8);
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[2 * t] = f.x; v[2 * t + 1] = f.y; s += f.x + f.y; }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) v[t] = 0.f;
    }
    const float mean = group8_sum(s) * (1.f / 64);
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) { const float d = v[t] - mean; q += d * d; }
    const float rstd = rsqrtf(group8_sum(q) * (1.f / 64) + eps);
    if (live) {
      __align__(16) __half2 o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        o[t] = __floats2half2_rn((v[2 * t] - mean) * rstd * wr[2 * t] + br[2 * t], (v[2 * t + 1] - mean) * rstd * wr[2 * t + 1] + br[2 * t + 1]);
      *reinterpret_cast<uint4 *>(y + row_off(r, inner_y, ldo_y, ldi_y) + sub * 8) = *reinterpret_cast<uint4 *>(o);
    }
  }
}
__global__ void k_ln64_bwd(const __half *__restrict__ x, int inner_x, long long ldo_x, long long ldi_x, const float *__restrict__ w, float eps,
                           const __half *__restrict__ dy, int inner_dy, long long ldo_dy, long long ldi_dy, const __half *__restrict__ add,
                           __half *__restrict__ dx, int inner_dx, long long ldo_dx, long long ldi_dx, long long rows) {
  const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
  float wr[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) wr[t] = w ? __ldg(w + sub * 8 + t) : 1.f;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long base = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4; base < rows; base += warps * 4) {
    const long long r = base + grp;
    const bool live = r < rows;
    float v[8], g[8];
    float s = 0.f;
    if (live) {
      uint4 u = *reinterpret_cast<const uint4 *>(x + row_off(r, inner_x, ldo_x, ldi_x) + sub * 8);
      uint4 ug = *reinterpret_cast<const uint4 *>(dy + row_off(r, inner_dy, ldo_dy, ldi_dy) + sub * 8);
      const __half2 *h = reinterpret_cast<const __half2 *>(&u), *hg = reinterpret_cast<const __half2 *>(&ug);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __half22float2(h[t]); v[2 * t] = f.x; v[2 * t + 1] = f.y; s += f.x + f.y;
        float2 fg = __half22float2(hg[t]); g[2 * t] = fg.x * wr[2 * t]; g[2 * t + 1] = fg.y * wr[2 * t + 1];
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) { v[t] = 0.f; g[t] = 0.f; }
    }
    const float mean = group8_sum(s) * (1.f / 64);
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) { const float d = v[t] - mean; q += d * d; }
    const float rstd = rsqrtf(group8_sum(q) * (1.f / 64) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) { const float xh = (v[t] - mean) * rstd; sg += g[t]; sgx += g[t] * xh; }
    sg = group8_sum(sg) * (1.f / 64);
    sgx = group8_sum(sgx) * (1.f / 64);
    if (live) {
      const long long off_dx = row_off(r, inner_dx, ldo_dx, ldi_dx) + sub * 8;
      float o[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float xh = (v[t] - mean) * rstd; o[t] = rstd * (g[t] - sg - xh * sgx); }
      if (add) {
        uint4 ua = *reinterpret_cast<const uint4 *>(add + off_dx);
        const __half2 *ha = reinterpret_cast<const __half2 *>(&ua);
#pragma unroll
        for (int t = 0; t < 4; ++t) { float2 f = __half22float2(ha[t]); o[2 * t] += f.x; o[2 * t + 1] += f.y; }
      }
      __align__(16) __half2 oh[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) oh[t] = __floats2half2_rn(o[2 * t], o[2 * t + 1]);
      *reinterpret_cast<uint4 *>(dx + off_dx) = *reinterpret_cast<uint4 *>(oh);
    }
  }
}

// ---------------------------------------------------------------- row softmax (materialised attention of the adjoint)
// P = softmax(S) over T columns; S fp32 (already scaled), P fp16.  One warp per row.
__global__ void k_softmax_fwd(const float *__restrict__ S, __half *__restrict__ P, long long rows, int T) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const float *sp = S + r * T;
  float m = -INFINITY;
  for (int i = lane * 4; i < T; i += 128) {
    float4 v = *reinterpret_cast<const float4 *>(sp + i);
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  m = warp_max(m);
  float sum = 0.f;
  for (int i = lane * 4; i < T; i += 128) {
    float4 v = *reinterpret_cast<const float4 *>(sp + i);
    sum += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
  }
  const float inv = 1.f / warp_sum(sum);
  __half *pp = P + r * T;
  for (int i = lane * 4; i < T; i += 128) {
    float4 v = *reinterpret_cast<const float4 *>(sp + i);
    __align__(8) __half2 o[2] = {__floats2half2_rn(__expf(v.x - m) * inv, __expf(v.y - m) * inv),
                                 __floats2half2_rn(__expf(v.z - m) * inv, __expf(v.w - m) * inv)};
    *reinterpret_cast<uint2 *>(pp + i) = *reinterpret_cast<uint2 *>(o);
  }
}
// dS = scale * P o (dP - rowsum(P o dP)); P fp16, dP fp32, dS fp16.
__global__ void k_softmax_bwd(const __half *__restrict__ P, const float *__restrict__ dP, __half *__restrict__ dS, long long rows, int T,
                              float scale) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const __half *pp = P + r * T;
  const float *dp = dP + r * T;
  float dot = 0.f;
  for (int i = lane * 4; i < T; i += 128) {
    uint2 u = *reinterpret_cast<const uint2 *>(pp + i);
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
    float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
    float4 d = *reinterpret_cast<const float4 *>(dp + i);
    dot += a.x * d.x + a.y * d.y + b.x * d.z + b.y * d.w;
  }
  dot = warp_sum(dot);
  __half *op = dS + r * T;
  for (int i = lane * 4; i < T; i += 128) {
    uint2 u = *reinterpret_cast<const uint2 *>(pp + i);
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
    float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
    float4 d = *reinterpret_cast<const float4 *>(dp + i);
    __align__(8) __half2 o[2] = {__floats2half2_rn(scale * a.x * (d.x - dot), scale * a.y * (d.y - dot)),
                                 __floats2half2_rn(scale * b.x * (d.z - dot), scale * b.y * (d.w - dot))};
    *reinterpret_cast<uint2 *>(op + i) = *reinterpret_cast<uint2 *>(o);
  }
}

// ---------------------------------------------------------------- Fourier embedding of the query points
// hy3dgen FourierEmbedder(num_freqs, include_pi=False): [x, sin(x f_k), cos(x f_k)], f_k = 2^k, feature order
// [xyz | sin(x f0..f7, y f0..f7, z f0..f7) | cos(...)], coordinates rounded to fp16 first (pipelines.py:302).
// Output fp16 [n, ld] with columns >= 3 + 6 num_freqs zeroed (K padded to a multiple of 64 for the tensor core).
__global__ void k_fourier(const float *__restrict__ xyz, __half *__restrict__ out, long long n, int ld, int num_freqs, float pi_mul) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * ld) return;
  const long long r = i / ld;
  const int c = (int)(i - r * ld);
  const int nf3 = 3 * num_freqs;
  float v = 0.f;
  if (c < 3) {
    v = __half2float(__float2half_rn(xyz[r * 3 + c]));
  } else if (c < 3 + 2 * nf3) {
    const int e = (c - 3) % nf3, is_cos = (c - 3) / nf3;
    const float x = __half2float(__float2half_rn(xyz[r * 3 + e / num_freqs]));
    // the reference forms x * f in fp16 before sin / cos
    const float a = __half2float(__float2half_rn(x * exp2f((float)(e % num_freqs)) * pi_mul));
    v = is_cos ? cosf(a) : sinf(a);
  }
  out[i] = __float2half_rn(v);
}

// ---------------------------------------------------------------- head: logits = output_proj(ln_post(x)); sdf = -logits
// One warp per row of width 1024; out[idx ? idx[r] : r] = -(w_out . LN(x) + b_out)  (float32: pipelines.py:309-312).
__global__ void k_head_fwd(const __half *__restrict__ x, long long ldx, const float *__restrict__ lw, const float *__restrict__ lb, float eps,
                           const float *__restrict__ wo, float bo, const int *__restrict__ idx, float *__restrict__ out, long long rows) {
  constexpr int W = 1024, CPL = 4;
  const int lane = threadIdx.x & 31;
  // w_out . (xhat * lw + lb) = sum_c xhat_c (lw_c wo_c) + sum_c lb_c wo_c: persistent warps keep this lane's 32 products
  // in registers and the constant once (read per row, the 96 scalar parameter loads per lane dominated the kernel)
  float a[CPL][8];
  float c0 = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int col = (lane + 32 * c) * 8 + t;
      const float o = __ldg(wo + col);
      a[c][t] = __ldg(lw + col) * o;
      c0 += __ldg(lb + col) * o;
    }
  c0 = warp_sum(c0) + bo;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const __half *xp = x + r * ldx;
    float v[CPL][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      uint4 u = *reinterpret_cast<const uint4 *>(xp + (lane + 32 * c) * 8);
      const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[c][2 * t] = f.x; v[c][2 * t + 1] = f.y; s += f.x + f.y; }
    }
    const float mean = warp_sum(s) * (1.f / W);
    float q = 0.f, acc = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c)
#pragma unroll
      for (int t = 0; t < 8; ++t) { const float d = v[c][t] - mean; q += d * d; acc += d * a[c][t]; }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / W) + eps);
    acc = warp_sum(acc);
    if (lane == 0) out[idx ? idx[r] : r] = -(acc * rstd + c0);
  }
}
// backward of the head for the rows that carry a gradient: dlogit = -g_scale * dS[idx ? idx[r] : r];
// dx = LN_bwd(dlogit * w_out)   (fp16, scaled by the caller's loss scale through g_scale)
__global__ void k_head_bwd(const __half *__restrict__ x, long long ldx, const float *__restrict__ lw, float eps, const float *__restrict__ wo,
                           const int *__restrict__ idx, const float *__restrict__ dS, float g_scale, __half *__restrict__ dx, long long lddx,
                           long long rows) {
  constexpr int W = 1024, CPL = 4;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const __half *xp = x + r * ldx;
  const float dl = -g_scale * dS[idx ? idx[r] : r];
  float v[CPL][8], g[CPL][8];
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    uint4 u = *reinterpret_cast<const uint4 *>(xp + (lane + 32 * c) * 8);
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) { float2 f = __half22float2(h[t]); v[c][2 * t] = f.x; v[c][2 * t + 1] = f.y; s += f.x + f.y; }
#pragma unroll
    for (int t = 0; t < 8; ++t) { const int col = (lane + 32 * c) * 8 + t; g[c][t] = dl * __ldg(wo + col) * __ldg(lw + col); }
  }
  const float mean = warp_sum(s) * (1.f / W);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int t = 0; t < 8; ++t) { const float d = v[c][t] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / W) + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int c = 0; c < CPL; ++c)
#pragma unroll
    for (int t = 0; t < 8; ++t) { const float xh = (v[c][t] - mean) * rstd; sg += g[c][t]; sgx += g[c][t] * xh; }
  sg = warp_sum(sg) * (1.f / W);
  sgx = warp_sum(sgx) * (1.f / W);
  __half *op = dx + r * lddx;
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    __align__(16) __half2 oh[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float xh0 = (v[c][2 * t] - mean) * rstd, xh1 = (v[c][2 * t + 1] - mean) * rstd;
      oh[t] = __floats2half2_rn(rstd * (g[c][2 * t] - sg - xh0 * sgx), rstd * (g[c][2 * t + 1] - sg - xh1 * sgx));
    }
    *reinterpret_cast<uint4 *>(op + (lane + 32 * c) * 8) = *reinterpret_cast<uint4 *>(oh);
  }
}

// ---------------------------------------------------------------- gather / casts
// out[r, :] = in[idx[r], :], rows of W halves (W multiple of 8), 16 bytes per thread
__global__ void k_gather_rows(const __half *__restrict__ in, long long ld_in, const int *__restrict__ idx, __half *__restrict__ out,
                              long long ld_out, long long rows, int W) {
  const int cpr = W / 8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cpr) return;
  const long long r = i / cpr;
  const int c = (int)(i - r * cpr);
  *reinterpret_cast<uint4 *>(out + r * ld_out + c * 8) = __ldg(reinterpret_cast<const uint4 *>(in + (long long)idx[r] * ld_in + c * 8));
}
// 2-D casts between row-major views (cols contiguous): mode 0 f32 -> f16, 1 f16 -> f32, 2 f16 -> f32 accumulate, 4 f16 -> f16 (x scale)
__global__ void k_cast2d(const void *__restrict__ in, long long ld_in, void *__restrict__ out, long long ld_out, long long rows, int cols,
                         float scale, int mode) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long long r = i / cols;
  const int c = (int)(i - r * cols);
  if (mode == 0) {
    reinterpret_cast<__half *>(out)[r * ld_out + c] = __float2half_rn(reinterpret_cast<const float *>(in)[r * ld_in + c] * scale);
  } else if (mode == 4) {     // a half tensor times a python float in torch: product in float, rounded to half
    reinterpret_cast<__half *>(out)[r * ld_out + c] = __float2half_rn(__half2float(reinterpret_cast<const __half *>(in)[r * ld_in + c]) * scale);
  } else {
    const float v = __half2float(reinterpret_cast<const __half *>(in)[r * ld_in + c]) * scale;
    float *o = reinterpret_cast<float *>(out) + r * ld_out + c;
    *o = (mode == 2) ? *o + v : v;
  }
}
// y = a + b (fp16), 8 elements per thread
__global__ void k_add_f16(const __half *a, const __half *b, __half *y, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 7 < n) {
    uint4 ua = *reinterpret_cast<const uint4 *>(a + i), ub = *reinterpret_cast<const uint4 *>(b + i);
    const __half2 *ha = reinterpret_cast<const __half2 *>(&ua), *hb = reinterpret_cast<const __half2 *>(&ub);
    __align__(16) __half2 o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) o[t] = __hadd2(ha[t], hb[t]);
    *reinterpret_cast<uint4 *>(y + i) = *reinterpret_cast<uint4 *>(o);
  } else {
    for (long long j = i; j < n; ++j) y[j] = __hadd(a[j], b[j]);
  }
}


// ---------------------------------------------------------------- sparse view of a dense gradient volume
// The energy touches a few thousand voxels of the lattice (trilinear corners of the hand vertices, the voxels inside
// both hand and object); the adjoint of the decoder only needs those rows.  Deterministic three-pass compaction
// (count per 4096-voxel block, exclusive scan, scatter) of the non-zero entries of g [B, V] into idx / val [B, cap].
constexpr int NZ_BLOCK = 4096;
__device__ __forceinline__ int nz_local(const float *g, long long V, long long base, int t, float (&v)[16]) {
  // element i of thread t = base + i * 256 + t: coalesced scalar loads (V and the per-image base need no alignment:
  // the reference's 65^3 lattice is odd)
  int c = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const long long e = base + (long long)i * 256 + t;
    v[i] = e < V ? __ldg(g + e) : 0.f;
    c += v[i] != 0.f;
  }
  return c;
}
__global__ void k_nz_count(const float *__restrict__ g, long long V, int nblk, int *__restrict__ counts) {
  const int b = blockIdx.y, blk = blockIdx.x, t = threadIdx.x;
  float v[16];
  int c = nz_local(g + (long long)b * V, V, (long long)blk * NZ_BLOCK, t, v);
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int ws[8];
  if ((t & 31) == 0) ws[t >> 5] = c;
  __syncthreads();
  if (t == 0) { int s = 0; for (int i = 0; i < 8; ++i) s += ws[i]; counts[(long long)b * nblk + blk] = s; }
}
__global__ void k_nz_scan(int *__restrict__ counts, int nblk, int *__restrict__ total, int cap, int *__restrict__ flags) {
  // one CTA per image: exclusive scan of its block counts in place
  const int b = blockIdx.x, t = threadIdx.x;
  int *c = counts + (long long)b * nblk;
  __shared__ int ws[32];
  __shared__ int carry;
  if (t == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + t;
    const int x = i < nblk ? c[i] : 0;
    int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += y; }
    if ((t & 31) == 31) ws[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
      int w = ws[t], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, wi, o); if (t >= o) wi += y; }
      ws[t] = wi - w;
    }
    __syncthreads();
    const int excl = carry + ws[t >> 5] + incl - x;
    if (i < nblk) c[i] = excl;
    __syncthreads();
    if (t == 1023) carry = excl + x;
    __syncthreads();
  }
  if (t == 0) { total[b] = carry; if (carry > cap && flags) atomicOr(flags, 1); }
}
__global__ void k_nz_scatter(const float *__restrict__ g, long long V, int nblk, const int *__restrict__ offsets, int cap, int *__restrict__ idx,
                             float *__restrict__ val) {
  const int b = blockIdx.y, blk = blockIdx.x, t = threadIdx.x;
  float v[16];
  const int c = nz_local(g + (long long)b * V, V, (long long)blk * NZ_BLOCK, t, v);
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, incl, o); if ((t & 31) >= o) incl += y; }
  __shared__ int ws[8];
  if ((t & 31) == 31) ws[t >> 5] = incl;
  __syncthreads();
  int pre = 0;
  for (int i = 0; i < (t >> 5); ++i) pre += ws[i];
  int pos = offsets[(long long)b * nblk + blk] + pre + incl - c;
  if (c == 0) return;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (v[i] != 0.f) {
      if (pos < cap) {
        idx[(long long)b * cap + pos] = (int)((long long)blk * NZ_BLOCK + (long long)i * 256 + t);
        val[(long long)b * cap + pos] = v[i];
      }
      ++pos;
    }
}

// delta[h][r] = <a[r][h][:], b[r][h][:]>, 64 channels: 8 lanes per (row, head), one 16-byte load each
__global__ void k_rowdot(const __half *__restrict__ a, long long lda, const __half *__restrict__ b, long long ldb, float *__restrict__ out,
                         long long out_head_stride, long long rows, int heads) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pair = i >> 3;
  const int sub = (int)(i & 7);
  float acc = 0.f;
  const bool live = pair < rows * heads;
  long long r = 0; int h = 0;
  if (live) {
    r = pair / heads; h = (int)(pair - r * heads);
    uint4 ua = __ldg(reinterpret_cast<const uint4 *>(a + r * lda + h * 64 + sub * 8));
    uint4 ub = __ldg(reinterpret_cast<const uint4 *>(b + r * ldb + h * 64 + sub * 8));
    const __half2 *ha = reinterpret_cast<const __half2 *>(&ua), *hb = reinterpret_cast<const __half2 *>(&ub);
#pragma unroll
    for (int t = 0; t < 4; ++t) { float2 x = __half22float2(ha[t]), y = __half22float2(hb[t]); acc += x.x * y.x + x.y * y.y; }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2); acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (live && sub == 0) out[(long long)h * out_head_stride + r] = acc;
}
__global__ void k_gather_f32(const float *__restrict__ src, long long hs, const int *__restrict__ idx, float *__restrict__ out, long long n,
                             int heads) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * heads) return;
  const int h = (int)(i / n);
  const long long k = i - (long long)h * n;
  out[i] = __ldg(src + (long long)h * hs + idx[k]);
}

inline int blocks_for(long long work, int per_block) { return (int)((work + per_block - 1) / per_block); }

}  // namespace

extern "C" int foho_dec_layernorm(const void *x, int32_t inner_x, int64_t ldo_x, int64_t ldi_x, const float *w, const float *b, float eps,
                                  void *y, int32_t inner_y, int64_t ldo_y, int64_t ldi_y, int64_t rows, int32_t width, void *cuda_stream) {
  if (!x || !y) return FOHO_E_NULL;
  if (rows <= 0 || inner_x <= 0 || inner_y <= 0 || (width != 64 && width != 1024)) return FOHO_E_SHAPE;
  if (ldo_x % 8 || ldi_x % 8 || ldo_y % 8 || ldi_y % 8) return FOHO_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const int g = blocks_for(rows, 8);
  if (width == 64) {
    const int g4 = blocks_for((rows + 3) / 4, 8);
    k_ln64_fwd<<<g4 < 148 * 8 ? g4 : 148 * 8, 256, 0, st>>>((const __half *)x, inner_x, ldo_x, ldi_x, w, b, eps, (__half *)y, inner_y, ldo_y,
                                                            ldi_y, rows);
  }
  else
    k_ln_fwd<1024><<<g < 148 * 2 ? g : 148 * 2, 256, 0, st>>>((const __half *)x, inner_x, ldo_x, ldi_x, w, b, eps, (__half *)y, inner_y, ldo_y,
                                                              ldi_y, rows);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_layernorm_bwd(const void *x, int32_t inner_x, int64_t ldo_x, int64_t ldi_x, const float *w, float eps, const void *dy,
                                      int32_t inner_dy, int64_t ldo_dy, int64_t ldi_dy, const void *add, void *dx, int32_t inner_dx,
                                      int64_t ldo_dx, int64_t ldi_dx, int64_t rows, int32_t width, void *cuda_stream) {
  if (!x || !dy || !dx) return FOHO_E_NULL;
  if (rows <= 0 || inner_x <= 0 || inner_dy <= 0 || inner_dx <= 0 || (width != 64 && width != 1024)) return FOHO_E_SHAPE;
  if (ldo_x % 8 || ldi_x % 8 || ldo_dy % 8 || ldi_dy % 8 || ldo_dx % 8 || ldi_dx % 8) return FOHO_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const int g = blocks_for(rows, 8);
  if (width == 64) {
    const int g4 = blocks_for((rows + 3) / 4, 8);
    k_ln64_bwd<<<g4 < 148 * 8 ? g4 : 148 * 8, 256, 0, st>>>((const __half *)x, inner_x, ldo_x, ldi_x, w, eps, (const __half *)dy, inner_dy,
                                                            ldo_dy, ldi_dy, (const __half *)add, (__half *)dx, inner_dx, ldo_dx, ldi_dx, rows);
  }
  else
    k_ln_bwd<1024><<<g < 148 * 2 ? g : 148 * 2, 256, 0, st>>>((const __half *)x, inner_x, ldo_x, ldi_x, w, eps, (const __half *)dy, inner_dy, ldo_dy, ldi_dy,
                                       (const __half *)add, (__half *)dx, inner_dx, ldo_dx, ldi_dx, rows);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_softmax(const float *S, void *P, int64_t rows, int32_t T, void *cuda_stream) {
  if (!S || !P) return FOHO_E_NULL;
  if (rows <= 0 || T <= 0 || T % 4) return FOHO_E_SHAPE;
  k_softmax_fwd<<<blocks_for(rows, 8), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(S, (__half *)P, rows, T);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_softmax_bwd(const void *P, const float *dP, void *dS, int64_t rows, int32_t T, float scale, void *cuda_stream) {
  if (!P || !dP || !dS) return FOHO_E_NULL;
  if (rows <= 0 || T <= 0 || T % 4) return FOHO_E_SHAPE;
  k_softmax_bwd<<<blocks_for(rows, 8), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>((const __half *)P, dP, (__half *)dS, rows, T,
                                                                                                scale);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_fourier_embed(const float *xyz, void *out, int64_t n, int32_t ld, int32_t num_freqs, int32_t include_pi,
                                      void *cuda_stream) {
  if (!xyz || !out) return FOHO_E_NULL;
  if (n <= 0 || num_freqs <= 0 || ld < 3 + 6 * num_freqs) return FOHO_E_SHAPE;
  k_fourier<<<blocks_for(n * ld, 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(xyz, (__half *)out, n, ld, num_freqs,
                                                                                               include_pi ? 3.14159265358979323846f : 1.f);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_head(const void *x, int64_t ldx, const float *ln_w, const float *ln_b, float eps, const float *w_out, float b_out,
                             const int32_t *idx, float *out, int64_t rows, void *cuda_stream) {
  if (!x || !ln_w || !ln_b || !w_out || !out) return FOHO_E_NULL;
  if (rows <= 0 || ldx % 8) return FOHO_E_SHAPE;
  const int hg = blocks_for(rows, 8);
  k_head_fwd<<<hg < 148 * 6 ? hg : 148 * 6, 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>((const __half *)x, ldx, ln_w, ln_b, eps, w_out,
                                                                                                    b_out, idx, out, rows);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_head_bwd(const void *x, int64_t ldx, const float *ln_w, float eps, const float *w_out, const int32_t *idx,
                                 const float *dS, float g_scale, void *dx, int64_t lddx, int64_t rows, void *cuda_stream) {
  if (!x || !ln_w || !w_out || !dS || !dx) return FOHO_E_NULL;
  if (rows <= 0 || ldx % 8 || lddx % 8) return FOHO_E_SHAPE;
  k_head_bwd<<<blocks_for(rows, 8), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>((const __half *)x, ldx, ln_w, eps, w_out, idx, dS,
                                                                                            g_scale, (__half *)dx, lddx, rows);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_gather_rows(const void *in, int64_t ld_in, const int32_t *idx, void *out, int64_t ld_out, int64_t rows, int32_t width,
                                    void *cuda_stream) {
  if (!in || !idx || !out) return FOHO_E_NULL;
  if (rows <= 0 || width <= 0 || width % 8 || ld_in % 8 || ld_out % 8) return FOHO_E_SHAPE;
  k_gather_rows<<<blocks_for(rows * (width / 8), 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      (const __half *)in, ld_in, idx, (__half *)out, ld_out, rows, width);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_cast(const void *in, int64_t ld_in, void *out, int64_t ld_out, int64_t rows, int32_t cols, float scale, int32_t mode,
                             void *cuda_stream) {
  // mode 0: f32 -> f16 (scaled); 1: f16 -> f32 (scaled); 2: f16 -> f32, accumulate; 3: out += in (fp16, contiguous rows*cols)
  if (!in || !out) return FOHO_E_NULL;
  if (rows <= 0 || cols <= 0) return FOHO_E_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const long long n = rows * cols;
  if ((mode >= 0 && mode <= 2) || mode == 4) k_cast2d<<<blocks_for(n, 256), 256, 0, st>>>(in, ld_in, out, ld_out, rows, cols, scale, mode);
  else if (mode == 3) {
    if (ld_in != cols || ld_out != cols) return FOHO_E_ARG;
    k_add_f16<<<blocks_for((n + 7) / 8, 256), 256, 0, st>>>((const __half *)in, (const __half *)out, (__half *)out, n);
  } else return FOHO_E_ARG;
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" size_t foho_dec_compact_workspace_bytes(int32_t B, int64_t V) {
  if (B <= 0 || V <= 0) return 0;
  return (size_t)B * (size_t)((V + NZ_BLOCK - 1) / NZ_BLOCK) * sizeof(int);
}

extern "C" int foho_dec_compact_grad(const float *g, int32_t B, int64_t V, int32_t cap, int32_t *idx, float *val, int32_t *count,
                                     int32_t *flags, void *workspace, size_t workspace_bytes, void *cuda_stream) {
  if (!g || !idx || !val || !count || !workspace) return FOHO_E_NULL;
  if (B <= 0 || V <= 0 || cap <= 0 || V > 0x7fffffffLL) return FOHO_E_SHAPE;
  if (workspace_bytes < foho_dec_compact_workspace_bytes(B, V)) return FOHO_E_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const int nblk = (int)((V + NZ_BLOCK - 1) / NZ_BLOCK);
  int *counts = reinterpret_cast<int *>(workspace);
  FOHO_CUDA_TRY(cudaMemsetAsync(idx, 0, (size_t)B * cap * sizeof(int), st));
  FOHO_CUDA_TRY(cudaMemsetAsync(val, 0, (size_t)B * cap * sizeof(float), st));
  k_nz_count<<<dim3(nblk, B), 256, 0, st>>>(g, V, nblk, counts);
  k_nz_scan<<<B, 1024, 0, st>>>(counts, nblk, count, cap, flags);
  k_nz_scatter<<<dim3(nblk, B), 256, 0, st>>>(g, V, nblk, counts, cap, idx, val);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_rowdot(const void *a, int64_t lda, const void *b, int64_t ldb, float *out, int64_t out_head_stride, int64_t rows,
                               int32_t heads, void *cuda_stream) {
  if (!a || !b || !out) return FOHO_E_NULL;
  if (rows <= 0 || heads <= 0 || lda % 8 || ldb % 8 || (out_head_stride != 0 && out_head_stride < rows)) return FOHO_E_SHAPE;
  k_rowdot<<<blocks_for(rows * heads * 8, 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(
      (const __half *)a, lda, (const __half *)b, ldb, out, out_head_stride ? out_head_stride : rows, rows, heads);
  FOHO_LAUNCH_CHECK();
  return 0;
}

extern "C" int foho_dec_gather_f32(const float *src, int64_t src_head_stride, const int32_t *idx, float *out, int64_t n, int32_t heads,
                                   void *cuda_stream) {
  if (!src || !idx || !out) return FOHO_E_NULL;
  if (n <= 0 || heads <= 0) return FOHO_E_SHAPE;
  k_gather_f32<<<blocks_for(n * heads, 256), 256, 0, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(src, src_head_stride, idx, out, n, heads);
  FOHO_LAUNCH_CHECK();
  return 0;
}
