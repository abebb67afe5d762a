// Fused optimiser + scheduler update of the guidance loop.
//
// Replaces, in one launch, `joint_optimizer.step()` -- torch.optim.AdamW(eps=1e-4) with
// torch defaults betas=(0.9,0.999), weight_decay=0.01 (reference
// third_party_patches/hy3dgen/shapegen/pipelines.py:1478,1601; phase 1 uses Adam, i.e.
// weight_decay=0, :1318) over the 6 scalar leaf groups + the velocity tensor
// (third_party/utilz/code_utils.py:57-78) -- and `scheduler.step_final`
// (third_party_patches/hy3dgen/shapegen/schedulers.py:470-484): x1 = x_t + (1-sigma) v.
//
// The update order mirrors torch's single-tensor AdamW:
//   p *= 1 - lr*wd ; m = lerp(m, g, 1-b1) ; v = b2*v + (1-b2) g*g ;
//   denom = sqrt(v)/sqrt(1-b2^t) + eps ; p -= (lr/(1-b1^t)) * m/denom
// HBM-bound elementwise stream: 5 reads + 4 writes of 4 B per velocity element.
#include "foho_common.cuh"
#include <cuda_fp16.h>
#include "foho_adamw.cuh"

namespace {

using AdamScalars = foho_adam_scalars_t;

// the 16 scalar leaves are float32 in every variant (pipelines.py:1208-1215): one thread per float
__device__ __forceinline__ void update_scalar_leaves(const foho_update_desc &d, const AdamScalars &s, int b, int k) {
  const int grp = k < 8 ? (k == 0 ? 0 : (k < 4 ? 1 : 2)) : (k == 8 ? 3 : (k < 12 ? 4 : 5));
  if ((d.theta_mask >> grp) & 1u) {
    const size_t o = (size_t)b * 16 + k;
    float p = d.theta[o], m = d.theta_m[o], v = d.theta_v[o];
    foho_adamw_one<false>(p, d.grad_theta[o], m, v, s.decay_theta[grp], s.neg_step_theta[grp], s);
    d.theta[o] = p; d.theta_m[o] = m; d.theta_v[o] = v;
  }
}

// NaN guard (pipelines.py:1442-1444,1590-1592): true = this sample's inner loop has been left, nothing of
// it is updated any more.  Every thread decides alike -- the flag only ever changes from 0 to non-zero, and
// only in a launch whose total is NaN, where the second test alone already says "skip".
__device__ __forceinline__ bool sample_halted(const foho_update_desc &d, int b) {
  if (!d.terms || !d.nan_flag) return false;
  const bool bad = isnan(d.terms[(size_t)b * FOHO_NUM_TERMS + FOHO_T_TOTAL]);
  const int flag = d.nan_flag[b];
  if (bad && flag == 0 && blockIdx.x == 0 && threadIdx.x == 0) d.nan_flag[b] = d.step;
  return bad || flag != 0;
}

__global__ void __launch_bounds__(256) k_update(foho_update_desc d, AdamScalars s) {
  const int b = blockIdx.y;
  if (sample_halted(d, b)) return;
  if (blockIdx.x == 0 && threadIdx.x < 16) update_scalar_leaves(d, s, b, threadIdx.x);
  if (!d.velocity) return;
  const float decay = s.decay_vel, neg_step = s.neg_step_vel;
  const float oms = 1.f - d.sigma;
  const size_t base = (size_t)b * d.L;
  const int L4 = d.L >> 2;
  float4 *P4 = reinterpret_cast<float4 *>(d.velocity + base);
  const float4 *G4 = reinterpret_cast<const float4 *>(d.grad_velocity + base);
  float4 *M4 = reinterpret_cast<float4 *>(d.vel_m + base);
  float4 *V4 = reinterpret_cast<float4 *>(d.vel_v + base);
  const float4 *X4 = d.x_t ? reinterpret_cast<const float4 *>(d.x_t + base) : nullptr;
  float4 *O4 = d.x1 ? reinterpret_cast<float4 *>(d.x1 + base) : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L4; i += gridDim.x * blockDim.x) {
    float4 p = P4[i], g = G4[i], m = M4[i], v = V4[i];
    foho_adamw_one<false>(p.x, g.x, m.x, v.x, decay, neg_step, s);
    foho_adamw_one<false>(p.y, g.y, m.y, v.y, decay, neg_step, s);
    foho_adamw_one<false>(p.z, g.z, m.z, v.z, decay, neg_step, s);
    foho_adamw_one<false>(p.w, g.w, m.w, v.w, decay, neg_step, s);
    P4[i] = p; M4[i] = m; V4[i] = v;
    if (X4 && O4) {
      float4 x = X4[i];
      // separately rounded product and sum, like torch's `sample + (1 - sigma) * model_output` (schedulers.py:481)
      O4[i] = make_float4(foho_step_final_one<false>(x.x, p.x, oms), foho_step_final_one<false>(x.y, p.y, oms),
                          foho_step_final_one<false>(x.z, p.z, oms), foho_step_final_one<false>(x.w, p.w, oms));
    }
  }
}

// fp16 velocity / latents (the reference's dtype: the leaf is a clone of the DiT's half output,
// code_utils.py:43-78; latents pipelines.py:1204).  torch's AdamW on a half parameter keeps the moments in
// half and rounds after EVERY elementwise op (each op computes in float and stores half) -- adamw_one<true>.
// 8 halves (16 B) per thread and access; L is a multiple of 8 (checked by the launcher).
struct __align__(16) Half8 { __half2 a, b, c, d; };

__device__ __forceinline__ void unpack8(const Half8 &h, float (&f)[8]) {
  float2 t;
  t = __half22float2(h.a); f[0] = t.x; f[1] = t.y;
  t = __half22float2(h.b); f[2] = t.x; f[3] = t.y;
  t = __half22float2(h.c); f[4] = t.x; f[5] = t.y;
  t = __half22float2(h.d); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ Half8 pack8(const float (&f)[8]) {
  Half8 h;
  h.a = __floats2half2_rn(f[0], f[1]); h.b = __floats2half2_rn(f[2], f[3]);
  h.c = __floats2half2_rn(f[4], f[5]); h.d = __floats2half2_rn(f[6], f[7]);
  return h;
}

__global__ void __launch_bounds__(256) k_update_f16(foho_update_desc d, AdamScalars s) {
  const int b = blockIdx.y;
  if (sample_halted(d, b)) return;
  if (blockIdx.x == 0 && threadIdx.x < 16) update_scalar_leaves(d, s, b, threadIdx.x);
  if (!d.velocity) return;
  const float decay = s.decay_vel, neg_step = s.neg_step_vel;
  // `(1 - sigma) * model_output`: the 0-dim fp32 factor is cast to half, the product is rounded to half,
  // the sum with the fp32 sample is taken in fp32 and cast back (schedulers.py:470-484; same rule as
  // k_sched_step_f16, which is pinned bit for bit by the reference scheduler's golden vectors)
  const float oms = 1.f - d.sigma;
  const size_t base = (size_t)b * d.L;
  const int L8 = d.L >> 3;
  Half8 *P8 = reinterpret_cast<Half8 *>(reinterpret_cast<__half *>(d.velocity) + base);
  const Half8 *G8 = reinterpret_cast<const Half8 *>(reinterpret_cast<const __half *>(d.grad_velocity) + base);
  Half8 *M8 = reinterpret_cast<Half8 *>(reinterpret_cast<__half *>(d.vel_m) + base);
  Half8 *V8 = reinterpret_cast<Half8 *>(reinterpret_cast<__half *>(d.vel_v) + base);
  const Half8 *X8 = d.x_t ? reinterpret_cast<const Half8 *>(reinterpret_cast<const __half *>(d.x_t) + base) : nullptr;
  Half8 *O8 = d.x1 ? reinterpret_cast<Half8 *>(reinterpret_cast<__half *>(d.x1) + base) : nullptr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L8; i += gridDim.x * blockDim.x) {
    float p[8], g[8], m[8], v[8];
    unpack8(P8[i], p); unpack8(G8[i], g); unpack8(M8[i], m); unpack8(V8[i], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) foho_adamw_one<true>(p[j], g[j], m[j], v[j], decay, neg_step, s);
    P8[i] = pack8(p); M8[i] = pack8(m); V8[i] = pack8(v);
    if (X8 && O8) {
      float x[8];
      unpack8(X8[i], x);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = foho_step_final_one<true>(x[j], p[j], oms);
      O8[i] = pack8(x);
    }
  }
}

__global__ void __launch_bounds__(256) k_sched_step(const float *__restrict__ x, const float *__restrict__ v,
                                                    float *__restrict__ prev, float *__restrict__ x1, long long n,
                                                    float dsig, float oms) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float xi = x[i], vi = v[i];
    // no FMA contraction: bit-equal to torch's separately rounded mul and add (schedulers.py:298,305)
    if (prev) prev[i] = __fadd_rn(xi, __fmul_rn(dsig, vi));
    if (x1) x1[i] = __fadd_rn(xi, __fmul_rn(oms, vi));
  }
}

// fp16 latents / model output (the reference's dtype, pipelines.py:1204): torch evaluates
// `sample.float() + (sigma_next - sigma) * model_output` with the 0-dim fp32 factor cast to half, the
// product rounded to half, the sum in fp32, and the result cast back to half.
__global__ void __launch_bounds__(256) k_sched_step_f16(const __half *__restrict__ x, const __half *__restrict__ v,
                                                        __half *__restrict__ prev, __half *__restrict__ x1, long long n,
                                                        float dsig, float oms) {
  const float dh = __half2float(__float2half_rn(dsig)), oh = __half2float(__float2half_rn(oms));
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float xi = __half2float(x[i]), vi = __half2float(v[i]);
    if (prev) prev[i] = __float2half_rn(__fadd_rn(xi, __half2float(__float2half_rn(__fmul_rn(dh, vi)))));
    if (x1) x1[i] = __float2half_rn(__fadd_rn(xi, __half2float(__float2half_rn(__fmul_rn(oh, vi)))));
  }
}

// LT = float | __half: the latents' dtype (the reference carries them in half, pipelines.py:1204; the decoded volume and its
// gradient are float like `latent2sdf(...).float()`, :309)
template <typename LT>
__global__ void __launch_bounds__(256) k_mock_dec_fwd(float *__restrict__ sdf, const float *__restrict__ sdf0,
                                                      const LT *__restrict__ x1, const long long *__restrict__ tap,
                                                      long long vol, int L, float alpha) {
  const int b = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) {
    const long long t = tap[j];
    sdf[(size_t)b * vol + t] = sdf0[(size_t)b * vol + t] + alpha * (float)x1[(size_t)b * L + j];
  }
}

template <typename LT>
__global__ void __launch_bounds__(256) k_mock_dec_bwd(const float *__restrict__ g, const long long *__restrict__ tap,
                                                      LT *__restrict__ gv, long long vol, int L, float scale) {
  const int b = blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x)
    gv[(size_t)b * L + j] = (LT)(scale * g[(size_t)b * vol + tap[j]]);
}

}  // namespace

template <typename LT>
static int mock_decoder_forward(float *sdf, const float *sdf0, const LT *x1, const int64_t *tap, int32_t B, int64_t vol, int32_t L,
                                float alpha, void *cuda_stream) {
  if (!sdf || !sdf0 || !x1 || !tap) return FOHO_E_NULL;
  if (B < 1 || vol < 1 || L < 1) return FOHO_E_SHAPE;
  int gx = (L + 255) / 256; if (gx > 296) gx = 296;
  k_mock_dec_fwd<LT><<<dim3(gx, B), 256, 0, (cudaStream_t)cuda_stream>>>(sdf, sdf0, x1, (const long long *)tap, vol, L, alpha);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

template <typename LT>
static int mock_decoder_backward(const float *grad_sdf, const int64_t *tap, LT *grad_velocity, int32_t B, int64_t vol, int32_t L,
                                 float scale, void *cuda_stream) {
  if (!grad_sdf || !tap || !grad_velocity) return FOHO_E_NULL;
  if (B < 1 || vol < 1 || L < 1) return FOHO_E_SHAPE;
  int gx = (L + 255) / 256; if (gx > 296) gx = 296;
  k_mock_dec_bwd<LT><<<dim3(gx, B), 256, 0, (cudaStream_t)cuda_stream>>>(grad_sdf, (const long long *)tap, grad_velocity, vol, L, scale);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

extern "C" int foho_mock_decoder_forward(float *sdf, const float *sdf0, const float *x1, const int64_t *tap, int32_t B,
                                         int64_t vol, int32_t L, float alpha, void *cuda_stream) {
  return mock_decoder_forward<float>(sdf, sdf0, x1, tap, B, vol, L, alpha, cuda_stream);
}
extern "C" int foho_mock_decoder_forward_f16(float *sdf, const float *sdf0, const void *x1, const int64_t *tap, int32_t B,
                                             int64_t vol, int32_t L, float alpha, void *cuda_stream) {
  return mock_decoder_forward<__half>(sdf, sdf0, (const __half *)x1, tap, B, vol, L, alpha, cuda_stream);
}
extern "C" int foho_mock_decoder_backward(const float *grad_sdf, const int64_t *tap, float *grad_velocity, int32_t B,
                                          int64_t vol, int32_t L, float scale, void *cuda_stream) {
  return mock_decoder_backward<float>(grad_sdf, tap, grad_velocity, B, vol, L, scale, cuda_stream);
}
extern "C" int foho_mock_decoder_backward_f16(const float *grad_sdf, const int64_t *tap, void *grad_velocity, int32_t B,
                                              int64_t vol, int32_t L, float scale, void *cuda_stream) {
  return mock_decoder_backward<__half>(grad_sdf, tap, (__half *)grad_velocity, B, vol, L, scale, cuda_stream);
}

static int update_checks(const foho_update_desc &d, int vec, int align) {
  if (!d.theta || !d.grad_theta || !d.theta_m || !d.theta_v) return FOHO_E_NULL;
  if (d.velocity && (!d.grad_velocity || !d.vel_m || !d.vel_v)) return FOHO_E_NULL;
  if ((d.terms == nullptr) != (d.nan_flag == nullptr)) return FOHO_E_NULL;      // the guard needs both
  if (d.B < 1 || d.step < 1 || (d.velocity && d.L < 1)) return FOHO_E_SHAPE;
  if (d.velocity && (d.L % vec) != 0) return FOHO_E_SHAPE;   // 16-byte vector stream
  if (d.velocity && (((uintptr_t)d.velocity | (uintptr_t)d.grad_velocity | (uintptr_t)d.vel_m | (uintptr_t)d.vel_v |
                      (uintptr_t)d.x_t | (uintptr_t)d.x1) & (uintptr_t)(align - 1)) != 0)
    return FOHO_E_ARG;
  return FOHO_OK;
}

static AdamScalars adam_scalars(const foho_update_desc &d) {
  return foho_adam_scalars(d.beta1, d.beta2, d.eps, d.weight_decay, d.lr_theta, d.lr_velocity, d.step);
}

extern "C" int foho_guidance_update(const foho_update_desc *dp, void *cuda_stream) {
  if (!dp) return FOHO_E_NULL;
  const foho_update_desc &d = *dp;
  if (int rc = update_checks(d, 4, 16)) return rc;
  int gx = 1;
  if (d.velocity) {
    gx = (d.L / 4 + 255) / 256;
    if (gx > 592) gx = 592;
    if (gx < 1) gx = 1;
  }
  k_update<<<dim3(gx, d.B), 256, 0, (cudaStream_t)cuda_stream>>>(d, adam_scalars(d));
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

extern "C" int foho_guidance_update_f16(const foho_update_desc *dp, void *cuda_stream) {
  if (!dp) return FOHO_E_NULL;
  const foho_update_desc &d = *dp;
  if (int rc = update_checks(d, 8, 16)) return rc;
  int gx = 1;
  if (d.velocity) {
    gx = (d.L / 8 + 255) / 256;
    if (gx > 592) gx = 592;
    if (gx < 1) gx = 1;
  }
  k_update_f16<<<dim3(gx, d.B), 256, 0, (cudaStream_t)cuda_stream>>>(d, adam_scalars(d));
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

extern "C" int foho_scheduler_step(const float *x_t, const float *velocity, float *prev_sample, float *pred_x1, int64_t n,
                                   float sigma, float sigma_next, void *cuda_stream) {
  if (!x_t || !velocity || (!prev_sample && !pred_x1)) return FOHO_E_NULL;
  if (n < 1) return FOHO_E_SHAPE;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  k_sched_step<<<(int)blocks, 256, 0, (cudaStream_t)cuda_stream>>>(x_t, velocity, prev_sample, pred_x1, n,
                                                                  sigma_next - sigma, 1.f - sigma);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

extern "C" int foho_scheduler_step_f16(const void *x_t, const void *velocity, void *prev_sample, void *pred_x1, int64_t n,
                                       float sigma, float sigma_next, void *cuda_stream) {
  if (!x_t || !velocity || (!prev_sample && !pred_x1)) return FOHO_E_NULL;
  if (n < 1) return FOHO_E_SHAPE;
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  k_sched_step_f16<<<(int)blocks, 256, 0, (cudaStream_t)cuda_stream>>>((const __half *)x_t, (const __half *)velocity,
                                                                      (__half *)prev_sample, (__half *)pred_x1, n,
                                                                      sigma_next - sigma, 1.f - sigma);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
