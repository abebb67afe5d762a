// Hand-written tcgen05 GEMM for the latent -> SDF decoder (row f1; the dense contractions of
// third_party_patches/hy3dgen/shapegen/pipelines.py:292-312 `latent2sdf`: ShapeVAE transformer + geo_decoder).
//
//   C[b][m][n] = epilogue( alpha * sum_k A[b](m,k) * B[b](n,k) )        fp16 operands, fp32 accumulation in TMEM
//
// Each operand is either K-major (rows of k, the nn.Linear weight layout [out,in] and activations [tokens,in]) or
// MN-major (k-rows of m / n: lets P^T, dS^T, V, K enter the attention-gradient products without a transpose).
//
// One persistent CTA per SM, warp-specialised:
//   warp 0   TMA producer: cp.async.bulk.tensor (128-byte swizzle) into a ring of kStages {A,B} tiles, mbarrier full/empty
//   warp 1   one thread issues tcgen05.mma (M=128, N=BN, K=16) into one of two TMEM accumulators, tcgen05.commit frees slots
//   warp 2   TMEM allocator
//   warps 4-11 epilogue (two warpgroups, half of the tile's columns each): tcgen05.ld -> bias / GELU / GELU' / residual -> global,
//             overlapped with the next tile's mainloop
#include "foho_common.cuh"
#include "foho_tc.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;

struct GemmParams {
  int M, N, K, batch;
  int tiles_m, tiles_n;
  int a_bcast, b_bcast;          // operand shared by every batch (batch stride 0)
  // epilogue
  void *C; long long ldc, bsc; int c_f32;
  const float *bias;
  const void *res; long long ldr, bsr; int res_f32;
  const __half *aux_in; __half *aux_out; long long ldaux, bsaux;
  float alpha; int act;
  int c_fast, aux_fast;          // 16-byte aligned rows: outputs leave through the coalescing shared-memory stage
  const float *row_vec; long long bs_rowvec;     // act 3 / 4: one float per output row
};

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_ROW = 144;                       // bytes per staged output row (128 + 16: conflict-free 16-byte writes)
  static constexpr int STG_BYTES = 8 * 16 * STG_ROW;         // eight epilogue warps x 16 rows (a chunk leaves in two halves)
#ifdef FOHO_GEMM_STAGES
  static constexpr int STAGES = FOHO_GEMM_STAGES;
#else
  static constexpr int STAGES = (192 * 1024 / STAGE_BYTES) > 8 ? 8 : (192 * 1024 / STAGE_BYTES);
#endif
  static constexpr int SMEM = STAGES * STAGE_BYTES + STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
};

// erf by Abramowitz & Stegun 7.1.28, 1 - (1 + a1 x + ... + a6 x^6)^-16 (|error| <= 3e-7, far below the fp16 rounding of the
// stored activation; 2e-6 in float arithmetic): six FMAs, four squarings and ONE special-function instruction (the
// reciprocal) -- the GELU epilogue of the 1024 -> 4096 GEMM is what the tile's 256 x 128 outputs spend their time on.
// (7.1.26, used before, needs a reciprocal AND an exponential: 855 -> 869 TFLOP/s for that GEMM.)  A power that overflows
// gives 1/inf = 0: erf = 1.
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  float q = fmaf(0.0000430638f, ax, 0.0002765672f);
  q = fmaf(q, ax, 0.0001520143f);
  q = fmaf(q, ax, 0.0092705272f);
  q = fmaf(q, ax, 0.0422820123f);
  q = fmaf(q, ax, 0.0705230784f);
  q = fmaf(q, ax, 1.0f);
  q *= q; q *= q; q *= q; q *= q;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(q));
  return copysignf(1.0f - r, x);
}
// The same for two elements with packed arithmetic: 20 instructions per pair against 19 per element.
__device__ __forceinline__ float2 gelu2(float2 x) {
  const float2 y = tc::fmul2(x, tc::splat2(0.70710678118654752f));
  const float2 ax = make_float2(fabsf(y.x), fabsf(y.y));
  float2 q = tc::ffma2(tc::splat2(0.0000430638f), ax, tc::splat2(0.0002765672f));
  q = tc::ffma2(q, ax, tc::splat2(0.0001520143f));
  q = tc::ffma2(q, ax, tc::splat2(0.0092705272f));
  q = tc::ffma2(q, ax, tc::splat2(0.0422820123f));
  q = tc::ffma2(q, ax, tc::splat2(0.0705230784f));
  q = tc::ffma2(q, ax, tc::splat2(1.0f));
  q = tc::fmul2(q, q); q = tc::fmul2(q, q); q = tc::fmul2(q, q); q = tc::fmul2(q, q);
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(q.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(q.y));
  float2 e = tc::ffma2(r, tc::splat2(-1.0f), tc::splat2(1.0f));
  e.x = copysignf(e.x, x.x); e.y = copysignf(e.y, x.y);
  const float2 hx = tc::fmul2(x, tc::splat2(0.5f));
  return tc::ffma2(hx, e, hx);
}
__device__ __forceinline__ float dgelu_f(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-0.72134752044448170f * x * x));
  return 0.5f * (1.0f + erf_fast(x * 0.70710678118654752f)) + x * 0.3989422804014327f * e;
}

// One warp's 32 x 32 chunk of outputs (lane = row) -> global memory in full 16-byte x 4 (fp16) / x 8 (fp32) row segments:
// each lane parks its row in shared memory, then the warp writes 8 (fp16) or 4 (fp32) rows per instruction, so an
// instruction touches 8 / 4 cache lines instead of 32 (row-per-lane stores were what bounded every K <= 1024 product).
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
// `stg` is the stage's SHARED-space address: through a generic pointer the compiler emitted generic LD.E / ST.E for the stage
// (address-space resolution in the load-store unit, long-scoreboard waits in the epilogue that bounds the bias + GELU tile).
template <bool F32>
__device__ __forceinline__ void staged_store(uint32_t stg, int lane, const float (&f)[32], void *C, long long row0_off, long long ld,
                                             int rows_valid, int cols_valid) {
  constexpr int SEG = F32 ? 8 : 4;            // 16-byte segments per row
  constexpr int RPI = 32 / SEG;               // rows per instruction
  constexpr int EPS = F32 ? 4 : 8;            // elements per segment
  const int seg = lane % SEG, rsub = lane / SEG;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {            // rows 0..15, then 16..31: the stage holds 16 rows per warp
    if ((lane >> 4) == hf) {
      const uint32_t mine = stg + (lane & 15) * 144;
      if (F32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          sts128(mine + j * 4, __float_as_uint(f[j]), __float_as_uint(f[j + 1]), __float_as_uint(f[j + 2]), __float_as_uint(f[j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t h[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const __half2 hh = __floats2half2_rn(f[j + 2 * t], f[j + 2 * t + 1]);
            h[t] = *reinterpret_cast<const uint32_t *>(&hh);
          }
          sts128(mine + j * 2, h[0], h[1], h[2], h[3]);
        }
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16 / RPI; ++i) {
      const int rl = i * RPI + rsub, r = hf * 16 + rl;
      if (r < rows_valid && seg * EPS < cols_valid) {
        const uint4 v = lds128(stg + rl * 144 + seg * 16);
        uint8_t *dst = reinterpret_cast<uint8_t *>(C) + (row0_off + (long long)r * ld + seg * EPS) * (F32 ? 4 : 2);
        *reinterpret_cast<uint4 *>(dst) = v;
      }
    }
    __syncwarp();
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(384, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using C_ = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *stg_base = smem + C_::STAGES * C_::STAGE_BYTES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(stg_base + C_::STG_BYTES);
  uint64_t *full = bars, *empty = bars + C_::STAGES, *tfull = bars + 2 * C_::STAGES, *tempty = tfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tempty + 2);

  const int warp = tc::warp_idx_uniform(), lane = threadIdx.x & 31;   // provably warp-uniform: see foho_tc.cuh
  const int nkb = (p.K + BK - 1) / BK;
  const int tiles_per_batch = p.tiles_m * p.tiles_n;
  const int num_tiles = tiles_per_batch * p.batch;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C_::STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 256); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<C_::TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: the whole warp walks the loop (uniform
    // control flow keeps coordinates and barrier addresses in uniform registers), one elected lane issues
    const bool leader = tc::elect_one();
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_batch, r = tile - b * tiles_per_batch;
      const int m0 = (r / p.tiles_n) * BM, n0 = (r % p.tiles_n) * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % C_::STAGES, ph = (it / C_::STAGES) & 1;
        tc::mbar_wait(&empty[s], ph ^ 1);
        uint8_t *sa = smem + s * C_::STAGE_BYTES, *sb = sa + C_::A_BYTES;
        const int k0 = kb * BK;
        if (leader) {
          tc::mbar_expect_tx(&full[s], C_::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tc::tma_load_3d(sa + j * (BK * 128), &tmA, &full[s], m0 + 64 * j, k0, p.a_bcast ? 0 : b);
          } else {
            tc::tma_load_3d(sa, &tmA, &full[s], k0, m0, p.a_bcast ? 0 : b);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tc::tma_load_3d(sb + j * (BK * 128), &tmB, &full[s], n0 + 64 * j, k0, p.b_bcast ? 0 : b);
          } else {
            tc::tma_load_3d(sb, &tmB, &full[s], k0, n0, p.b_bcast ? 0 : b);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues: the
    // descriptors are built in uniform registers and the tcgen05.mma instructions leave back to back)
    const bool leader = tc::elect_one();
    constexpr uint32_t idesc = tc::idesc_f16(BM, BN, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
    uint32_t it = 0, acc_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
      const uint32_t acc = acc_it & 1, acc_ph = (acc_it >> 1) & 1;
      tc::mbar_wait(&tempty[acc], acc_ph ^ 1);
      tc::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % C_::STAGES, ph = (it / C_::STAGES) & 1;
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + s * C_::STAGE_BYTES), sb = sa + C_::A_BYTES;
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 16 halves = 32 B further along the 128-B swizzled row; MN-major: 16 k-rows = 2048 B further
            const uint64_t ad = A_MN ? tc::smem_desc_sw128(sa + k * (UMMA_K * 128), BK * 128, 1024)
                                     : tc::smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
            const uint64_t bd = B_MN ? tc::smem_desc_sw128(sb + k * (UMMA_K * 128), BK * 128, 1024)
                                     : tc::smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
            tc::mma_f16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          tc::mma_commit(&empty[s]);      // slot free once these MMAs have read it
          if (kb + 1 == nkb) tc::mma_commit(&tfull[acc]);      // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (256 threads: thread = one row of the tile, the two
    // warpgroups take the lower / upper half of its columns)
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;   // which half of the tile's columns
    uint32_t acc_it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++acc_it) {
      const int b = tile / tiles_per_batch, r = tile - b * tiles_per_batch;
      const int m0 = (r / p.tiles_n) * BM, n0 = (r % p.tiles_n) * BN;
      const uint32_t acc = acc_it & 1, acc_ph = (acc_it >> 1) & 1;
      tc::mbar_wait(&tfull[acc], acc_ph);
      tc::tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const bool row_ok = m < p.M;
      const float rv = (p.act >= 3 && row_ok) ? __ldg(p.row_vec + (long long)b * p.bs_rowvec + m) : 0.f;
      // Measured and dropped: the next chunk's tcgen05.ld in flight while this one is processed, the accumulator handed back as
      // soon as its last chunk is in registers (the MMA warp spins ~110 times per tile on `tempty` with the GELU epilogue):
      // 32 more live registers, spills and copies -- 1 027 / 1 087 / 933 against 1 173 / 1 193 / 1 062 TFLOP/s.
#pragma unroll 1
      for (int c = half * (BN / 64); c < (half + 1) * (BN / 64); ++c) {
        uint32_t v[32];
        tc::tmem_ld32(tmem_base + acc * BN + c * 32 + ((uint32_t)(q * 32) << 16), v);
        const int n = n0 + c * 32;
        const bool full_chunk = (n + 32 <= p.N);
        // the bias row of the chunk is fetched under the latency of the accumulator load
        const bool fast_bias = p.bias && full_chunk && ((reinterpret_cast<uintptr_t>(p.bias + n) & 15) == 0);
        float4 bb[8];
        if (fast_bias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bb[j] = __ldg(reinterpret_cast<const float4 *>(p.bias + n) + j);
        }
        // ... and so is this thread's 64-byte piece of a half-precision residual row (c_proj + x, fc2 + x: an L2 round trip
        // that the chunk otherwise waits for after its arithmetic)
        const __half *res_h = reinterpret_cast<const __half *>(p.res) + (long long)b * p.bsr + (long long)m * p.ldr + n;
        const bool fast_res = p.res && !p.res_f32 && row_ok && full_chunk && ((reinterpret_cast<uintptr_t>(res_h) & 15) == 0);
        // (the same registers serve the saved pre-activations of the GELU' epilogue: that product has no residual)
        const __half *aux_h = p.aux_in + (long long)b * p.bsaux + (long long)m * p.ldaux + n;
        const bool fast_aux = p.act == 2 && !p.res && row_ok && full_chunk && ((reinterpret_cast<uintptr_t>(aux_h) & 15) == 0);
        uint4 rr[4];
        if (fast_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = *(reinterpret_cast<const uint4 *>(res_h) + j);
        } else if (fast_aux) {
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = __ldg(reinterpret_cast<const uint4 *>(aux_h) + j);
        }
        tc::tmem_ld_wait();
        if (n >= p.N) continue;                           // warp-uniform: whole chunk outside the matrix
        const int rows_valid = min(32, p.M - (m0 + q * 32)), cols_valid = min(32, p.N - n);
        const uint32_t stg = tc::smem_u32(stg_base) + (warp - 4) * (16 * C_::STG_ROW);
        const long long row0 = (long long)(m0 + q * 32);
        float f[32];
        float pre[32];
        if (row_ok) {
          if (fast_bias) {
            // alpha * acc + bias in packed pairs, the bias row in 16-byte loads (the same address in every lane)
            const float2 al = tc::splat2(p.alpha);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = bb[j >> 2];
              const float2 lo = tc::ffma2(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), al, make_float2(b4.x, b4.y));
              const float2 hi = tc::ffma2(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), al, make_float2(b4.z, b4.w));
              f[j] = lo.x; f[j + 1] = lo.y; f[j + 2] = hi.x; f[j + 3] = hi.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (full_chunk || n + j < p.N) f[j] += __ldg(p.bias + n + j);
            }
          }
          if (p.aux_out) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pre[j] = f[j];
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 g2 = gelu2(make_float2(f[j], f[j + 1]));
              f[j] = g2.x; f[j + 1] = g2.y;
            }
          } else if (p.act == 3) {            // softmax probabilities from saved log-sum-exps: exp2(alpha acc - lse2[m])
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = exp2f(f[j] - rv);
          } else if (p.act == 4) {            // softmax backward: P o (dP - delta[m]) (alpha carries the score scale)
            const __half *ai = p.aux_in + (long long)b * p.bsaux + (long long)m * p.ldaux + n;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(ai) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u = __ldg(reinterpret_cast<const uint4 *>(ai + j));
                const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  float2 x = __half22float2(h[t]);
                  f[j + 2 * t] = x.x * (f[j + 2 * t] - p.alpha * rv);
                  f[j + 2 * t + 1] = x.y * (f[j + 2 * t + 1] - p.alpha * rv);
                }
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n + j < p.N) f[j] = __half2float(ai[j]) * (f[j] - p.alpha * rv);
            }
          } else if (p.act == 2) {
            const __half *ai = aux_h;
            if (full_chunk && ((reinterpret_cast<uintptr_t>(ai) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 u = fast_aux ? rr[j >> 3] : __ldg(reinterpret_cast<const uint4 *>(ai + j));
                const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  float2 x = __half22float2(h[t]);
                  f[j + 2 * t] *= dgelu_f(x.x);
                  f[j + 2 * t + 1] *= dgelu_f(x.y);
                }
              }
            } else {
              _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n + j < p.N) f[j] *= dgelu_f(__half2float(ai[j]));
            }
          }
          if (p.res) {
            if (p.res_f32) {
              const float *rp = reinterpret_cast<const float *>(p.res) + (long long)b * p.bsr + (long long)m * p.ldr + n;
              if (full_chunk && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  float4 x = *reinterpret_cast<const float4 *>(rp + j);
                  f[j] += x.x; f[j + 1] += x.y; f[j + 2] += x.z; f[j + 3] += x.w;
                }
              } else {
                _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n + j < p.N) f[j] += rp[j];
              }
            } else {
              const __half *rp = res_h;
              if (fast_res) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  const uint4 u = rr[j >> 3];
                  const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
                  for (int t = 0; t < 4; ++t) {
                    float2 x = __half22float2(h[t]);
                    f[j + 2 * t] += x.x; f[j + 2 * t + 1] += x.y;
                  }
                }
              } else {
                _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n + j < p.N) f[j] += __half2float(rp[j]);
              }
            }
          }
        }
        if (p.aux_out) {
          if (p.aux_fast) {
            staged_store<false>(stg, lane, pre, p.aux_out, (long long)b * p.bsaux + row0 * p.ldaux + n, p.ldaux, rows_valid, cols_valid);
          } else if (row_ok) {
            __half *ao = p.aux_out + (long long)b * p.bsaux + (long long)m * p.ldaux + n;
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n + j < p.N) ao[j] = __float2half_rn(pre[j]);
          }
        }
        if (p.c_fast) {
          if (p.c_f32) staged_store<true>(stg, lane, f, p.C, (long long)b * p.bsc + row0 * p.ldc + n, p.ldc, rows_valid, cols_valid);
          else staged_store<false>(stg, lane, f, p.C, (long long)b * p.bsc + row0 * p.ldc + n, p.ldc, rows_valid, cols_valid);
        } else if (row_ok) {
          if (p.c_f32) {
            float *cp = reinterpret_cast<float *>(p.C) + (long long)b * p.bsc + (long long)m * p.ldc + n;
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n + j < p.N) cp[j] = f[j];
          } else {
            __half *cp = reinterpret_cast<__half *>(p.C) + (long long)b * p.bsc + (long long)m * p.ldc + n;
            _Pragma("unroll") for (int j = 0; j < 32; ++j) if (n + j < p.N) cp[j] = __float2half_rn(f[j]);
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tempty[acc]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<C_::TMEM_COLS>(tmem_base);
  }
}

template <int BN, bool A_MN, bool B_MN>
int launch_gemm(const foho_gemm_desc *d, cudaStream_t st) {
  using C_ = Cfg<BN>;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: dims {K, rows, batch}, box {64, BM|BN}; MN-major: dims {rows, K, batch}, box {64, BK}
  const int a_bcast = d->batch > 1 && d->bsa == 0, b_bcast = d->batch > 1 && d->bsb == 0;
  const int nba = a_bcast ? 1 : d->batch, nbb = b_bcast ? 1 : d->batch;
  if (A_MN) rc = tc::make_tmap_f16(&tmA, d->A, d->M, d->K, nba, d->lda, d->bsa, BK);
  else      rc = tc::make_tmap_f16(&tmA, d->A, d->K, d->M, nba, d->lda, d->bsa, BM);
  if (rc) return rc;
  if (B_MN) rc = tc::make_tmap_f16(&tmB, d->B, d->N, d->K, nbb, d->ldb, d->bsb, BK);
  else      rc = tc::make_tmap_f16(&tmB, d->B, d->K, d->N, nbb, d->ldb, d->bsb, BN);
  if (rc) return rc;
  GemmParams p;
  p.a_bcast = a_bcast; p.b_bcast = b_bcast;
  p.M = d->M; p.N = d->N; p.K = d->K; p.batch = d->batch;
  p.tiles_m = (d->M + BM - 1) / BM; p.tiles_n = (d->N + BN - 1) / BN;
  p.C = d->C; p.ldc = d->ldc; p.bsc = d->bsc; p.c_f32 = d->c_f32;
  p.bias = d->bias;
  p.res = d->res; p.ldr = d->ldr; p.bsr = d->bsr; p.res_f32 = d->res_f32;
  p.aux_in = reinterpret_cast<const __half *>(d->aux_in); p.aux_out = reinterpret_cast<__half *>(d->aux_out);
  p.ldaux = d->ldaux; p.bsaux = d->bsaux;
  p.alpha = d->alpha; p.act = d->act;
  {
    const int es = d->c_f32 ? 4 : 2;
    p.c_fast = ((reinterpret_cast<uintptr_t>(d->C) & 15) == 0) && (d->ldc * es) % 16 == 0 && (d->bsc * es) % 16 == 0 && d->N % (16 / es) == 0;
    p.aux_fast = d->aux_out && ((reinterpret_cast<uintptr_t>(d->aux_out) & 15) == 0) && (d->ldaux * 2) % 16 == 0 && (d->bsaux * 2) % 16 == 0 &&
                 d->N % 8 == 0;
  }
  p.row_vec = d->row_vec; p.bs_rowvec = d->bs_rowvec;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    FOHO_CUDA_TRY(cudaGetDevice(&dev));
    FOHO_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  FOHO_CUDA_TRY(cudaFuncSetAttribute(k_gemm_tc<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C_::SMEM));
  long long tiles = (long long)p.tiles_m * p.tiles_n * p.batch;
  int grid = (int)(tiles < sm_count ? tiles : sm_count);
  if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
  k_gemm_tc<BN, A_MN, B_MN><<<grid, 384, C_::SMEM, st>>>(tmA, tmB, p);
  FOHO_LAUNCH_CHECK();
  return 0;
}

template <int BN>
int dispatch_major(const foho_gemm_desc *d, cudaStream_t st) {
  if (d->a_mn_major) return d->b_mn_major ? launch_gemm<BN, true, true>(d, st) : launch_gemm<BN, true, false>(d, st);
  return d->b_mn_major ? launch_gemm<BN, false, true>(d, st) : launch_gemm<BN, false, false>(d, st);
}

}  // namespace

extern "C" int foho_tc_gemm(const foho_gemm_desc *d, void *cuda_stream) {
  if (!d || !d->A || !d->B || !d->C) return FOHO_E_NULL;
  if (d->M <= 0 || d->N <= 0 || d->K <= 0 || d->batch <= 0) return FOHO_E_ARG;
  if (d->act < 0 || d->act > 4 || ((d->act == 2 || d->act == 4) && !d->aux_in) || (d->act >= 3 && !d->row_vec)) return FOHO_E_ARG;
  if (d->K % 8 || d->lda % 8 || d->ldb % 8) return FOHO_E_ARG;   // 16-byte TMA strides
  if ((d->a_mn_major && d->M % 8) || (d->b_mn_major && d->N % 8)) return FOHO_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  int bn = d->block_n;
  if (bn == 0) bn = d->N > 128 ? 256 : (d->N > 64 ? 128 : 64);
  switch (bn) {
    case 64: return dispatch_major<64>(d, st);
    case 128: return dispatch_major<128>(d, st);
    case 256: return dispatch_major<256>(d, st);
  }
  return FOHO_E_ARG;
}
