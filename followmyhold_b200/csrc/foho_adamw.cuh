// The arithmetic of the fused AdamW / step_final update, host + device.
// Kept __host__ __device__ (like foho_math.cuh) so tests/csrc_host_check.cpp can run the kernels' exact
// op sequence on the build box (no GPU there) against torch.optim.AdamW.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#endif

#if defined(__CUDACC__)
#define FOHO_AHD __host__ __device__ __forceinline__
#else
#define FOHO_AHD inline
#endif

// individually rounded IEEE single ops (never re-associated or contracted by the compiler)
#if defined(__CUDA_ARCH__)
#define FOHO_A_MUL(a, b) __fmul_rn((a), (b))
#define FOHO_A_ADD(a, b) __fadd_rn((a), (b))
#define FOHO_A_SUB(a, b) __fsub_rn((a), (b))
#define FOHO_A_DIV(a, b) __fdiv_rn((a), (b))
#define FOHO_A_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define FOHO_A_SQRT(a) __fsqrt_rn((a))
#define FOHO_A_RH(a) __half2float(__float2half_rn((a)))
#else
static inline float foho_a_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float foho_a_add(float a, float b) { volatile float r = a + b; return r; }
static inline float foho_a_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float foho_a_div(float a, float b) { volatile float r = a / b; return r; }
static inline float foho_a_rh(float a) { volatile _Float16 h = (_Float16)a; return (float)h; }   // RNE
#define FOHO_A_MUL(a, b) foho_a_mul((a), (b))
#define FOHO_A_ADD(a, b) foho_a_add((a), (b))
#define FOHO_A_SUB(a, b) foho_a_sub((a), (b))
#define FOHO_A_DIV(a, b) foho_a_div((a), (b))
#define FOHO_A_FMA(a, b, c) fmaf((a), (b), (c))
#define FOHO_A_SQRT(a) sqrtf((a))
#define FOHO_A_RH(a) foho_a_rh((a))
#endif

// Scalars exactly as torch forms them: python doubles, cast to the kernels' float opmath type.
struct foho_adam_scalars_t {
  float b2, one_m_b1, one_m_b2, eps;
  float bc2_sqrt;                            // (1 - b2^t) ** 0.5
  float decay_theta[6], neg_step_theta[6];   // 1 - lr*wd, -(lr / (1 - b1^t)) per leaf group
  float decay_vel, neg_step_vel;
};

// The descriptor carries the hyper-parameters as float; torch forms every derived scalar from the python
// doubles the user wrote (0.9, 0.999, 1e-4, 0.01, 1e-2 ...).  The shortest decimal that round-trips the float
// recovers that double (0.9f -> 0.9, not 0.89999997615814209), so 1-b1, 1-lr*wd, lr/bc1 come out as torch's.
static inline double foho_as_written(float f) {
  char buf[32];
  for (int prec = 1; prec <= 9; ++prec) {
    snprintf(buf, sizeof buf, "%.*g", prec, (double)f);
    if (strtof(buf, nullptr) == f) return strtod(buf, nullptr);
  }
  return (double)f;
}

static inline foho_adam_scalars_t foho_adam_scalars(float beta1, float beta2, float eps, float weight_decay,
                                                    const float *lr_theta6, float lr_velocity, int step) {
  const double b1 = foho_as_written(beta1), b2 = foho_as_written(beta2), wd = foho_as_written(weight_decay);
  foho_adam_scalars_t s;
  s.b2 = (float)b2;
  s.one_m_b1 = (float)(1.0 - b1);
  s.one_m_b2 = (float)(1.0 - b2);
  s.eps = (float)foho_as_written(eps);
  // bias corrections in double, as torch's python floats: 1 - beta**step, (1 - beta2**step) ** 0.5
  const double bc1 = 1.0 - pow(b1, (double)step);
  const double bc2 = 1.0 - pow(b2, (double)step);
  s.bc2_sqrt = (float)pow(bc2, 0.5);
  for (int g = 0; g < 6; ++g) {
    const double lr = foho_as_written(lr_theta6[g]);
    s.decay_theta[g] = (float)(1.0 - lr * wd);
    s.neg_step_theta[g] = (float)((lr / bc1) * -1.0);
  }
  const double lr = foho_as_written(lr_velocity);
  s.decay_vel = (float)(1.0 - lr * wd);
  s.neg_step_vel = (float)((lr / bc1) * -1.0);
  return s;
}

// One AdamW update in the op order and rounding of torch's multi-tensor (foreach) CUDA path -- the one the
// reference takes for CUDA parameters (torch/optim/adamw.py `_multi_tensor_adamw`; call sites
// pipelines.py:1318,1384,1478):
//   p.mul_(1-lr*wd); m.lerp_(g, 1-b1); v.mul_(b2); v.addcmul_(g, g, 1-b2);
//   den = v.sqrt(); den.div_(sqrt(bc2)); den.add_(eps); p.addcdiv_(m, den, -lr/bc1)
// Every op reads its operands in float and rounds its result to the tensors' dtype (identity for float,
// round-to-nearest-even half for fp16 leaves).  lerp / addcmul / addcdiv are `a + s*x` expressions that
// nvcc contracts to one FMA inside torch's kernels, hence the explicit fused multiply-adds here.
template <bool HALF>
FOHO_AHD float foho_rnd(float x) {
  if (HALF) return FOHO_A_RH(x);
  return x;
}

template <bool HALF>
FOHO_AHD void foho_adamw_one(float &p, float g, float &m, float &v, float decay, float neg_step,
                             const foho_adam_scalars_t &s) {
  p = foho_rnd<HALF>(FOHO_A_MUL(p, decay));
  m = foho_rnd<HALF>(FOHO_A_FMA(s.one_m_b1, FOHO_A_SUB(g, m), m));
  v = foho_rnd<HALF>(FOHO_A_MUL(v, s.b2));
  v = foho_rnd<HALF>(FOHO_A_FMA(s.one_m_b2, FOHO_A_MUL(g, g), v));
  float den = foho_rnd<HALF>(FOHO_A_SQRT(v));
  den = foho_rnd<HALF>(FOHO_A_DIV(den, s.bc2_sqrt));
  den = foho_rnd<HALF>(FOHO_A_ADD(den, s.eps));
  p = foho_rnd<HALF>(FOHO_A_FMA(neg_step, FOHO_A_DIV(m, den), p));
}

// x1 = x_t + (1 - sigma) * v  (schedulers.py:470-484).  float: separately rounded product and sum.
// half: the 0-dim fp32 factor is cast to half, the product rounded to half, the sum with the fp32-upcast
// sample taken in fp32 and cast back (the rule k_sched_step_f16 follows, pinned by the reference
// scheduler's golden vectors).
template <bool HALF>
FOHO_AHD float foho_step_final_one(float x, float v, float one_minus_sigma) {
  if (HALF) return FOHO_A_RH(FOHO_A_ADD(x, FOHO_A_RH(FOHO_A_MUL(FOHO_A_RH(one_minus_sigma), v))));
  return FOHO_A_ADD(x, FOHO_A_MUL(one_minus_sigma, v));
}
