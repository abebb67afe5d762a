// Hand-written tcgen05 attention forward for the latent -> SDF decoder (row f1): the cross attention of the
// lattice queries onto the 3072 latent tokens (hy3dgen geo_decoder, called at
// third_party_patches/hy3dgen/shapegen/pipelines.py:304) and the self attention of the ShapeVAE transformer
// (pipelines.py:299).  Head dimension 64, no mask, softmax scale 1/8.
//
//   O[i][q][h] = softmax_k( Q[q][h] . K[i][k][h] / 8 ) V[i][k][h]
//
// One persistent CTA per SM walks work items (image, head, tile of 128 queries); per item the 3072 keys stream
// through in blocks of 128:
//   warp 0     TMA producer: Q tile once, then a ring of {K block, V block} stages (128-byte swizzle)
//   warp 1     one thread issues tcgen05.mma: S_j = Q K_j^T (M128 N128 K64) into one of two TMEM buffers,
//              O += P_j V_j (M128 N64 K128; P from shared memory, V as an MN-major operand -- no transpose)
//   warp 2     TMEM allocator (512 columns: S0 | S1 | O)
//   warps 4-7  online softmax, one query row per thread: tcgen05.ld S_j -> exp2 in registers -> P_j (fp16) into
//              swizzled shared memory; rescale O in TMEM (tcgen05.ld/st) only when the running maximum moved by
//              more than 2^8; final 1/l normalisation and the fp16 store
// S_{j+1} is issued before P_j is consumed, so the tensor core works on the next scores while the softmax runs.
#include "foho_common.cuh"
#include "foho_tc.cuh"
#include <cuda_fp16.h>

namespace {

constexpr int AQ = 128;          // queries per tile
constexpr int AK = 128;          // keys per block
constexpr int HD = 64;           // head dimension
constexpr int KV_STAGES = 3;
constexpr int Q_BYTES = AQ * HD * 2, K_BYTES = AK * HD * 2, V_BYTES = AK * HD * 2, P_BYTES = AQ * AK * 2;
constexpr int ATT_SMEM = Q_BYTES + KV_STAGES * (K_BYTES + V_BYTES) + 2 * P_BYTES + 1024 + 256;
constexpr uint32_t TM_S0 = 0, TM_S1 = 128, TM_O = 256, TM_COLS = 512;

struct AttnParams {
  int n_img, heads, n_q, n_k;          // queries per image, keys per image (multiple of 128)
  int q_tiles;
  long long q_rows_per_img;            // row offset of image i's queries in the Q tensor map (0: queries shared)
  __half *out; long long ldo, out_img_stride;   // O[i][q][h*64 + d]
  float scale_log2;                    // softmax scale * log2(e)
  float *lse2;                         // optional [n_img][heads][lse_stride]: m + log2(l) of the scaled scores (variant 0 only)
  long long lse_stride;
};

__global__ void __launch_bounds__(256, 1)
k_attn_fwd(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
           const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem;
  uint8_t *sKV = sQ + Q_BYTES;                                  // stage s: K at s*(K+V), V right after
  uint8_t *sP = sKV + KV_STAGES * (K_BYTES + V_BYTES);          // two P buffers
  uint64_t *bars = reinterpret_cast<uint64_t *>(sP + 2 * P_BYTES);
  uint64_t *q_full = bars, *q_empty = bars + 1, *o_empty = bars + 2;
  uint64_t *k_full = bars + 3, *v_full = k_full + KV_STAGES, *kv_empty = v_full + KV_STAGES;
  uint64_t *s_full = kv_empty + KV_STAGES, *s_empty = s_full + 2, *p_full = s_empty + 2, *p_empty = p_full + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(p_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.n_k / AK;
  const int items_per_img = p.heads * p.q_tiles;
  const int n_items = p.n_img * items_per_img;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(q_full, 1); tc::mbar_init(q_empty, 1); tc::mbar_init(o_empty, 128);
    for (int i = 0; i < KV_STAGES; ++i) { tc::mbar_init(&k_full[i], 1); tc::mbar_init(&v_full[i], 1); tc::mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 128);
      tc::mbar_init(&p_full[i], 128); tc::mbar_init(&p_empty[i], 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<TM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    uint32_t g = 0, w = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
      const int img = item / items_per_img, r = item - img * items_per_img;
      const int h = r / p.q_tiles, qt = r - h * p.q_tiles;
      tc::mbar_wait(q_empty, (w & 1) ^ 1);
      tc::mbar_expect_tx(q_full, Q_BYTES);
      tc::tma_load_3d(sQ, &tmQ, q_full, 0, (int)(img * p.q_rows_per_img) + qt * AQ, h);
      for (int j = 0; j < nblk; ++j, ++g) {
        const uint32_t s = g % KV_STAGES, ph = (g / KV_STAGES) & 1;
        tc::mbar_wait(&kv_empty[s], ph ^ 1);
        uint8_t *sk = sKV + s * (K_BYTES + V_BYTES), *sv = sk + K_BYTES;
        tc::mbar_expect_tx(&k_full[s], K_BYTES);
        tc::tma_load_3d(sk, &tmK, &k_full[s], 0, img * p.n_k + j * AK, h);
        tc::mbar_expect_tx(&v_full[s], V_BYTES);
        tc::tma_load_3d(sv, &tmV, &v_full[s], 0, img * p.n_k + j * AK, h);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = tc::idesc_f16(AQ, AK, 0, 0);       // S = Q K^T : both K-major
    constexpr uint32_t idesc_o = tc::idesc_f16(AQ, HD, 0, 1);       // O += P V  : P K-major, V MN-major
    const uint32_t q_addr = tc::smem_u32(sQ);
    uint32_t g = 0, w = 0;
    auto issue_s = [&](uint32_t gg) {
      const uint32_t s = gg % KV_STAGES, ph = (gg / KV_STAGES) & 1, u = gg & 1, n = gg >> 1;
      tc::mbar_wait(&k_full[s], ph);
      tc::mbar_wait(&s_empty[u], (n & 1) ^ 1);
      tc::tc_fence_after();
      const uint32_t k_addr = tc::smem_u32(sKV + s * (K_BYTES + V_BYTES));
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        tc::mma_f16_ss(tmem_base + (u ? TM_S1 : TM_S0), tc::smem_desc_sw128(q_addr + k * 32, 16, 1024),
                       tc::smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0);
      tc::mma_commit(&s_full[u]);
    };
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
      tc::mbar_wait(q_full, w & 1);
      issue_s(g);
      for (int j = 0; j < nblk; ++j, ++g) {
        if (j + 1 < nblk) issue_s(g + 1);
        if (j + 1 == nblk) tc::mma_commit(q_empty);          // every S product of this item has been issued
        const uint32_t s = g % KV_STAGES, ph = (g / KV_STAGES) & 1, u = g & 1, n = g >> 1;
        tc::mbar_wait(&p_full[u], n & 1);
        tc::mbar_wait(&v_full[s], ph);
        if (j == 0) tc::mbar_wait(o_empty, (w & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t p_addr = tc::smem_u32(sP + u * P_BYTES);
        const uint32_t v_addr = tc::smem_u32(sKV + s * (K_BYTES + V_BYTES) + K_BYTES);
#pragma unroll
        for (int k = 0; k < AK / 16; ++k)
          tc::mma_f16_ss(tmem_base + TM_O, tc::smem_desc_sw128(p_addr + (k >> 2) * (AQ * 128) + (k & 3) * 32, 16, 1024),
                         tc::smem_desc_sw128(v_addr + k * 2048, AK * 128, 1024), idesc_o, (j | k) != 0);
        tc::mma_commit(&kv_empty[s]);
        tc::mma_commit(&p_empty[u]);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax / correction / epilogue
    const int q = warp & 3;
    const int row = q * 32 + lane;                                  // query row of this thread inside the tile
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    uint32_t g = 0, w = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
      const int img = item / items_per_img, r = item - img * items_per_img;
      const int h = r / p.q_tiles, qt = r - h * p.q_tiles;
      float m_used = 0.f, l = 0.f;
      for (int j = 0; j < nblk; ++j, ++g) {
        const uint32_t u = g & 1, n = g >> 1;
        tc::mbar_wait(&s_full[u], n & 1);
        tc::tc_fence_after();
        uint32_t sv[4][32];
        const uint32_t s_addr = tmem_base + (u ? TM_S1 : TM_S0) + lane_off;
#pragma unroll
        for (int c = 0; c < 4; ++c) tc::tmem_ld32(s_addr + c * 32, sv[c]);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&s_empty[u]);                               // scores are in registers: buffer free
        float bm = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) bm = fmaxf(bm, __uint_as_float(sv[c][i]));
        bm *= p.scale_log2;
        float corr = 1.f;
        bool need = false;
        if (j == 0) {
          m_used = bm;
        } else if (bm > m_used + 8.f) {                             // stale maximum is fine while p <= 2^8
          corr = exp2f(m_used - bm);
          m_used = bm;
          need = true;
        }
        float sum = 0.f;
        uint32_t ph[4][16];                                         // P row as packed halves
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float a = exp2f(fmaf(__uint_as_float(sv[c][i]), p.scale_log2, -m_used));
            const float b = exp2f(fmaf(__uint_as_float(sv[c][i + 1]), p.scale_log2, -m_used));
            const __half2 hh = __floats2half2_rn(a, b);
            // the sum uses the rounded values the tensor core will see, so rows of P V / l sum to one
            const float2 rr = __half22float2(hh);
            sum += rr.x + rr.y;
            ph[c][i >> 1] = *reinterpret_cast<const uint32_t *>(&hh);
          }
        l = l * corr + sum;
        if (__any_sync(0xffffffffu, need)) {
          // O must not change under a running P V product: wait for the previous block's
          const uint32_t gp = g - 1;
          tc::mbar_wait(&p_empty[gp & 1], (gp >> 1) & 1);
          tc::tc_fence_after();
          uint32_t ov[2][32];
#pragma unroll
          for (int c = 0; c < 2; ++c) tc::tmem_ld32(tmem_base + TM_O + lane_off + c * 32, ov[c]);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[c][i] = __float_as_uint(__uint_as_float(ov[c][i]) * corr);
#pragma unroll
          for (int c = 0; c < 2; ++c) tc::tmem_st32(tmem_base + TM_O + lane_off + c * 32, ov[c]);
          tc::tmem_st_wait();
        }
        tc::mbar_wait(&p_empty[u], (n & 1) ^ 1);                    // P buffer free (product of two blocks ago done)
        uint8_t *pb = sP + u * P_BYTES + row * 128;
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {                           // 16-byte chunk cc of the row = keys 8cc .. 8cc+7
          const int blk = cc >> 3, c = cc & 7;
          uint4 val = make_uint4(ph[cc >> 2][(cc & 3) * 4], ph[cc >> 2][(cc & 3) * 4 + 1], ph[cc >> 2][(cc & 3) * 4 + 2],
                                 ph[cc >> 2][(cc & 3) * 4 + 3]);
          *reinterpret_cast<uint4 *>(pb + blk * (AQ * 128) + ((c ^ (row & 7)) << 4)) = val;
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        tc::mbar_arrive(&p_full[u]);
      }
      // ---- epilogue: O / l
      {
        const uint32_t gp = g - 1;
        tc::mbar_wait(&p_empty[gp & 1], (gp >> 1) & 1);
        tc::tc_fence_after();
        uint32_t ov[2][32];
#pragma unroll
        for (int c = 0; c < 2; ++c) tc::tmem_ld32(tmem_base + TM_O + lane_off + c * 32, ov[c]);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(o_empty);
        const int qrow = qt * AQ + row;
        if (qrow < p.n_q) {
          const float inv = 1.f / l;
          __half *op = p.out + (long long)img * p.out_img_stride + (long long)qrow * p.ldo + h * HD;
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              __align__(16) __half2 hh[4];
#pragma unroll
              for (int t = 0; t < 4; ++t)
                hh[t] = __floats2half2_rn(__uint_as_float(ov[c][i + 2 * t]) * inv, __uint_as_float(ov[c][i + 2 * t + 1]) * inv);
              *reinterpret_cast<uint4 *>(op + c * 32 + i) = *reinterpret_cast<uint4 *>(hh);
            }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<TM_COLS>(tmem_base);
  }
}


// ================================================================================================
// Variant 2 (default): two query tiles per CTA in ping-pong.  With one softmax warp per scheduler the exp2 / pack /
// store chain of a tile cannot hide its own latencies and the tensor core idles for two thirds of the time (384
// TFLOP/s); a second softmax warpgroup on a second tile fills those slots.  One S buffer per tile (TMEM: S_A | S_B |
// O_A | O_B), registers moved from the producer / MMA warps to the softmax warps with setmaxnreg.
// Measured and dropped (profiles/r02_tc_attention_probe.json history in DESIGN.md section 8): packed-half exponentials
// (ex2.approx.f16x2 compiles to two MUFU.EX2.F16 + PRMT: 205 TFLOP/s), and row sums on the tensor core (P V made 80 wide
// with a tile of ones) with the max pass folded into the exponential loop (545 vs 614 TFLOP/s).
//   warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 softmax of tile A, warps 8-11 softmax of tile B
//   MMA order per key block j:  S_A(j+1), S_B(j+1) (as soon as the groups hold block j in registers), then
//   [P_A(j) ready -> O_A += P_A V_j]  [P_B(j) ready -> O_B += P_B V_j]
// ================================================================================================
// PT = 1 (default since the P-in-TMEM rewrite): P never touches shared memory.  The softmax warps write the fp16
// probabilities back into TMEM (tcgen05.st, two halves per 32-bit column: lane = query row, column c = keys 2c, 2c+1) and
// the product O += P V takes its A operand from there (tcgen05.mma [d], [a_tmem], b_desc).  With P in shared memory
// the P V product reads 6 KB of operands per 32-cycle instruction -- more than the 128 B/clk the shared memory
// delivers -- while the sixteen 16-byte stores per row and block (2.4 wavefronts each: bank conflicts between the four
// quarter-warps) compete for the same port; from TMEM the product reads only V, and the 64 KB of P buffers become two
// more K/V stages.  TMEM: S_A | S_B | O_A | O_B | P_A | P_B = 128 + 128 + 64 + 64 + 64 + 64 = 512 columns.
#ifdef FOHO_ATTN_TRACE
// Debug build only (-DFOHO_ATTN_TRACE): SM-clock stamps of CTA 0's second work item, [role][block][event]; role 0 / 1 =
// lane 0 of the first warp of softmax group A / B, role 2 = the MMA issuer.  Read back with foho_debug_attn_trace.
__device__ long long g_attn_trace[9][32][8];   // roles 0-3: group A warps, 4-7: group B warps, 8: MMA
#define ATT_TRACE(role, blk, ev) do { if (blockIdx.x == 0 && w == 1 && (blk) < 32) g_attn_trace[role][blk][ev] = clock64(); } while (0)
#else
#define ATT_TRACE(role, blk, ev) do { } while (0)
#endif
template <int PT> struct Att2Cfg {
  static constexpr int STAGES = PT ? 5 : KV_STAGES;
  static constexpr int SMEM = 2 * Q_BYTES + STAGES * (K_BYTES + V_BYTES) + (PT ? 0 : 2 * P_BYTES) + 1024 + 256;
};
constexpr uint32_t T2_S = 0, T2_O = 256, T2_P = 384;          // S_g at g*128, O_g at 256 + g*64, P_g at 384 + g*64

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x (K/2) columns of packed halves
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int PT>
__global__ void __launch_bounds__(384, 1)
k_attn_fwd2(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
            const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem;                                           // tile A, tile B
  uint8_t *sKV = sQ + 2 * Q_BYTES;
  constexpr int KV_STAGES = Att2Cfg<PT>::STAGES;                // shadows the file-scope constant
  uint8_t *sP = sKV + KV_STAGES * (K_BYTES + V_BYTES);          // P_A, P_B (PT = 0 only)
  uint64_t *bars = reinterpret_cast<uint64_t *>(sP + (PT ? 0 : 2 * P_BYTES));
  uint64_t *q_full = bars, *q_empty = bars + 1;
  uint64_t *k_full = bars + 2, *v_full = k_full + KV_STAGES, *kv_empty = v_full + KV_STAGES;
  uint64_t *s_full = kv_empty + KV_STAGES, *s_empty = s_full + 2, *p_full = s_empty + 2, *p_empty = p_full + 2, *o_empty = p_empty + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(o_empty + 2);

  const int warp = tc::warp_idx_uniform(), lane = threadIdx.x & 31;
  const int nblk = p.n_k / AK;
  const int pair_tiles = (p.q_tiles + 1) / 2;
  const int items_per_img = p.heads * pair_tiles;
  const int n_items = p.n_img * items_per_img;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(q_full, 1); tc::mbar_init(q_empty, 1);
    for (int i = 0; i < KV_STAGES; ++i) { tc::mbar_init(&k_full[i], 1); tc::mbar_init(&v_full[i], 1); tc::mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1); tc::mbar_init(&s_empty[i], 128);
      tc::mbar_init(&p_full[i], 128); tc::mbar_init(&p_empty[i], 1); tc::mbar_init(&o_empty[i], 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<TM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // provably warp-uniform

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    if (warp == 0) {
      // ---------------------------------------------------------- TMA producer (whole warp walks the loop, one lane issues)
      const bool leader = tc::elect_one();
      uint32_t g = 0, w = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
        const int img = item / items_per_img, r = item - img * items_per_img;
        const int h = r / pair_tiles, pt = r - h * pair_tiles;
        tc::mbar_wait(q_empty, (w & 1) ^ 1);
        const int qrow0 = (int)(img * p.q_rows_per_img) + pt * 2 * AQ;
        if (leader) {
          tc::mbar_expect_tx(q_full, 2 * Q_BYTES);
          tc::tma_load_3d(sQ, &tmQ, q_full, 0, qrow0, h);
          tc::tma_load_3d(sQ + Q_BYTES, &tmQ, q_full, 0, qrow0 + AQ, h);      // beyond the last row: zero filled
        }
        __syncwarp();
        for (int j = 0; j < nblk; ++j, ++g) {
          const uint32_t s = g % KV_STAGES, ph = (g / KV_STAGES) & 1;
          tc::mbar_wait(&kv_empty[s], ph ^ 1);
          uint8_t *sk = sKV + s * (K_BYTES + V_BYTES), *sv = sk + K_BYTES;
          if (leader) {
            tc::mbar_expect_tx(&k_full[s], K_BYTES);
            tc::tma_load_3d(sk, &tmK, &k_full[s], 0, img * p.n_k + j * AK, h);
            tc::mbar_expect_tx(&v_full[s], V_BYTES);
            tc::tma_load_3d(sv, &tmV, &v_full[s], 0, img * p.n_k + j * AK, h);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer (whole warp walks the loop, one lane issues)
      const bool leader = tc::elect_one();
      constexpr uint32_t idesc_s = tc::idesc_f16(AQ, AK, 0, 0);
      constexpr uint32_t idesc_o = tc::idesc_f16(AQ, HD, 0, 1);
      uint32_t g = 0, w = 0;
      auto issue_s = [&](uint32_t gg, int grp) {          // S_grp(gg) = Q_grp K_gg^T
        const uint32_t s = gg % KV_STAGES, ph = (gg / KV_STAGES) & 1;
        tc::mbar_wait(&k_full[s], ph);
        tc::mbar_wait(&s_empty[grp], (gg & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t q_addr = tc::smem_u32(sQ + grp * Q_BYTES);
        const uint32_t k_addr = tc::smem_u32(sKV + s * (K_BYTES + V_BYTES));
        if (leader) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            tc::mma_f16_ss(tmem_base + T2_S + grp * 128, tc::smem_desc_sw128(q_addr + k * 32, 16, 1024),
                           tc::smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0);
          tc::mma_commit(&s_full[grp]);
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
        tc::mbar_wait(q_full, w & 1);
        issue_s(g, 0);
        issue_s(g, 1);
        for (int j = 0; j < nblk; ++j, ++g) {
          const uint32_t s = g % KV_STAGES, ph = (g / KV_STAGES) & 1;
          const uint32_t v_addr = tc::smem_u32(sKV + s * (K_BYTES + V_BYTES) + K_BYTES);
          // the next scores first: a softmax group frees its S buffer as soon as the scores are in its registers,
          // so S(j+1) is computed while the group is still exponentiating block j
          ATT_TRACE(8, j, 0);
          if (j + 1 < nblk) { issue_s(g + 1, 0); ATT_TRACE(8, j, 1); issue_s(g + 1, 1); }
          else { if (leader) tc::mma_commit(q_empty); __syncwarp(); }
          ATT_TRACE(8, j, 2);
#pragma unroll
          for (int grp = 0; grp < 2; ++grp) {
            tc::mbar_wait(&p_full[grp], g & 1);
            ATT_TRACE(8, j, 3 + 2 * grp);
            if (grp == 0) tc::mbar_wait(&v_full[s], ph);
            if (j == 0) tc::mbar_wait(&o_empty[grp], (w & 1) ^ 1);
            tc::tc_fence_after();
            if (leader) {
              if (PT) {
#pragma unroll
                for (int k = 0; k < AK / 16; ++k)
                  mma_f16_ts(tmem_base + T2_O + grp * 64, tmem_base + T2_P + grp * 64 + k * 8,
                             tc::smem_desc_sw128(v_addr + k * 2048, AK * 128, 1024), idesc_o, (j | k) != 0);
              } else {
                const uint32_t p_addr = tc::smem_u32(sP + grp * P_BYTES);
#pragma unroll
                for (int k = 0; k < AK / 16; ++k)
                  tc::mma_f16_ss(tmem_base + T2_O + grp * 64, tc::smem_desc_sw128(p_addr + (k >> 2) * (AQ * 128) + (k & 3) * 32, 16, 1024),
                                 tc::smem_desc_sw128(v_addr + k * 2048, AK * 128, 1024), idesc_o, (j | k) != 0);
              }
              tc::mma_commit(&p_empty[grp]);
              if (grp == 1) tc::mma_commit(&kv_empty[s]);
            }
            __syncwarp();
            ATT_TRACE(8, j, 4 + 2 * grp);
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ------------------------------------------------------------ softmax group grp on its query tile
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t s_addr = tmem_base + T2_S + grp * 128 + lane_off;
    const uint32_t o_addr = tmem_base + T2_O + grp * 64 + lane_off;
    const uint32_t pb_s = tc::smem_u32(sP + (PT ? 0 : grp * P_BYTES) + row * 128);
    const uint32_t p_addr = tmem_base + T2_P + grp * 64 + lane_off;
    uint32_t g = 0, w = 0;
    // The exponential phase is bound by the MUFU pipe (128 ex2 per thread and block at 16 per clock and SM: 1 050 cycles
    // per warp and phase with two warps on a scheduler, 1 240 with one -- scripts/ubench/exp_phase.cu), the rest of a
    // block (tcgen05.ld, max pass, barrier round trips) takes ~550 cycles.  clock64 traces of both groups
    // (scripts/attn_trace.py, -DFOHO_ATTN_TRACE): ~2 900 cycles per block pair with the groups in phase.  Measured and
    // dropped on top of this version (all within +-5 % of it): named-barrier turn taking on the
    // phase (strict, arrival at the last exponential, arrival at the middle pinned with a volatile load),
    // FlashAttention-4's polynomial exp2 on the FMA pipe for 1-4 of every 8 key pairs; and speculating on the running
    // maximum (exponentials with the maximum of the blocks before, the block maximum reduced inside the pass, the pass
    // repeated when it exceeds the lazy-rescale bound): 704 against 827 TFLOP/s.
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
      const int img = item / items_per_img, r = item - img * items_per_img;
      const int h = r / pair_tiles, pt = r - h * pair_tiles;
      float m_used = 0.f, l = 0.f;
      for (int j = 0; j < nblk; ++j, ++g) {
        const bool tr = (lane == 0);
        if (tr) ATT_TRACE(grp * 4 + q, j, 0);
        tc::mbar_wait(&s_full[grp], g & 1);
        if (tr) ATT_TRACE(grp * 4 + q, j, 1);
        tc::tc_fence_after();
        uint32_t sv[4][32];
        // the second half of the row is in flight while the maximum of the first half is reduced (eight chains)
        tc::tmem_ld32(s_addr, sv[0]); tc::tmem_ld32(s_addr + 32, sv[1]);
        tc::tmem_ld_wait();
        tc::tmem_ld32(s_addr + 64, sv[2]); tc::tmem_ld32(s_addr + 96, sv[3]);
        float bm0 = -INFINITY, bm1 = -INFINITY, bm2 = -INFINITY, bm3 = -INFINITY, bm4 = -INFINITY, bm5 = -INFINITY, bm6 = -INFINITY, bm7 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          bm0 = fmaxf(bm0, fmaxf(__uint_as_float(sv[0][i]), __uint_as_float(sv[0][i + 1])));
          bm1 = fmaxf(bm1, fmaxf(__uint_as_float(sv[0][i + 2]), __uint_as_float(sv[0][i + 3])));
          bm2 = fmaxf(bm2, fmaxf(__uint_as_float(sv[1][i]), __uint_as_float(sv[1][i + 1])));
          bm3 = fmaxf(bm3, fmaxf(__uint_as_float(sv[1][i + 2]), __uint_as_float(sv[1][i + 3])));
        }
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(&s_empty[grp]);
        if (tr) ATT_TRACE(grp * 4 + q, j, 2);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          bm4 = fmaxf(bm4, fmaxf(__uint_as_float(sv[2][i]), __uint_as_float(sv[2][i + 1])));
          bm5 = fmaxf(bm5, fmaxf(__uint_as_float(sv[2][i + 2]), __uint_as_float(sv[2][i + 3])));
          bm6 = fmaxf(bm6, fmaxf(__uint_as_float(sv[3][i]), __uint_as_float(sv[3][i + 1])));
          bm7 = fmaxf(bm7, fmaxf(__uint_as_float(sv[3][i + 2]), __uint_as_float(sv[3][i + 3])));
        }
        const float bm = fmaxf(fmaxf(fmaxf(bm0, bm1), fmaxf(bm2, bm3)), fmaxf(fmaxf(bm4, bm5), fmaxf(bm6, bm7))) * p.scale_log2;
        float corr = 1.f;
        bool need = false;
        if (j == 0) {
          m_used = bm;
        } else if (bm > m_used + 8.f) {
          corr = ex2_approx(m_used - bm);
          m_used = bm;
          need = true;
        }
        // the previous product of this tile must be complete before O is rescaled or P overwritten
        if (tr) ATT_TRACE(grp * 4 + q, j, 3);
        tc::mbar_wait(&p_empty[grp], (g & 1) ^ 1);
        if (tr) ATT_TRACE(grp * 4 + q, j, 4);
        if (__any_sync(0xffffffffu, need)) {
          tc::tc_fence_after();
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tc::tmem_ld32(o_addr + c * 32, ov);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * corr);
            tc::tmem_st32(o_addr + c * 32, ov);
          }
          tc::tmem_st_wait();
        }
        float sum0 = 0.f, sum1 = 0.f;
        const float neg_m = -m_used;
        if (tr) ATT_TRACE(grp * 4 + q, j, 7);
        if (PT) {
          const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
          const float2 nm2 = make_float2(neg_m, neg_m);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < 2; ++c) {                              // 64 keys = 32 packed columns per store
            uint32_t pk[32];
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const int e = (2 * t) & 31;
              const float2 x = tc::ffma2(make_float2(__uint_as_float(sv[c * 2 + (t >> 4)][e]), __uint_as_float(sv[c * 2 + (t >> 4)][e + 1])), sc2, nm2);
              float2 ab;
              ab.x = ex2_approx(x.x); ab.y = ex2_approx(x.y);
              if (t & 1) acc1 = tc::fadd2(acc1, ab); else acc0 = tc::fadd2(acc0, ab);
              const __half2 hh = __floats2half2_rn(ab.x, ab.y);
              pk[t] = *reinterpret_cast<const uint32_t *>(&hh);
            }
            tc::tmem_st32(p_addr + c * 32, pk);
          }
          sum0 = acc0.x + acc0.y; sum1 = acc1.x + acc1.y;
          if (tr) ATT_TRACE(grp * 4 + q, j, 5);
          tc::tmem_st_wait();
        } else {
#pragma unroll
        for (int cc = 0; cc < 16; ++cc) {                          // 16-byte chunk cc = keys 8cc .. 8cc+7
          uint32_t pk[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int e = (cc & 3) * 8 + 2 * t;
            const float a = ex2_approx(fmaf(__uint_as_float(sv[cc >> 2][e]), p.scale_log2, neg_m));
            const float b = ex2_approx(fmaf(__uint_as_float(sv[cc >> 2][e + 1]), p.scale_log2, neg_m));
            sum0 += a; sum1 += b;
            const __half2 hh = __floats2half2_rn(a, b);
            pk[t] = *reinterpret_cast<const uint32_t *>(&hh);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pb_s + (cc >> 3) * (AQ * 128) + (((cc & 7) ^ (row & 7)) << 4)),
                       "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                       : "memory");
        }
        }
        l = l * corr + (sum0 + sum1);
        if (!PT) tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        tc::mbar_arrive(&p_full[grp]);
        if (tr) ATT_TRACE(grp * 4 + q, j, 6);
      }
      // ---- epilogue: O / l
      tc::mbar_wait(&p_empty[grp], (g & 1) ^ 1);                    // the last product of this item (block g-1)
      tc::tc_fence_after();
      const int qrow = (pt * 2 + grp) * AQ + row;
      const float inv = 1.f / l;
      if (p.lse2 && qrow < p.n_q) p.lse2[((long long)img * p.heads + h) * p.lse_stride + qrow] = m_used + log2f(l);
      __half *op = p.out + (long long)img * p.out_img_stride + (long long)qrow * p.ldo + h * HD;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ov[32];
        tc::tmem_ld32(o_addr + c * 32, ov);
        tc::tmem_ld_wait();
        if (qrow < p.n_q) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            __align__(16) __half2 hh[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              hh[t] = __floats2half2_rn(__uint_as_float(ov[i + 2 * t]) * inv, __uint_as_float(ov[i + 2 * t + 1]) * inv);
            *reinterpret_cast<uint4 *>(op + c * 32 + i) = *reinterpret_cast<uint4 *>(hh);
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&o_empty[grp]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<TM_COLS>(tmem_base);
  }
}


}  // namespace

#ifdef FOHO_ATTN_TRACE
extern "C" int foho_debug_attn_trace(long long *host_dst) {
  return cudaMemcpyFromSymbol(host_dst, g_attn_trace, sizeof(g_attn_trace)) == cudaSuccess ? 0 : -1;
}
#endif

extern "C" int foho_tc_attention(const foho_attn_desc *d, void *cuda_stream) {
  if (!d || !d->q || !d->k || !d->v || !d->out) return FOHO_E_NULL;
  if (d->n_img <= 0 || d->heads <= 0 || d->n_q <= 0 || d->n_k <= 0 || d->n_k % AK) return FOHO_E_SHAPE;
  if (d->ldq % 8 || d->ldk % 8 || d->ldv % 8 || d->ldo % 8 || d->hsq % 8 || d->hsk % 8 || d->hsv % 8) return FOHO_E_ARG;
  if (reinterpret_cast<uintptr_t>(d->out) & 15) return FOHO_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t q_rows = d->q_shared ? (uint64_t)d->n_q : (uint64_t)d->n_q * d->n_img;
  int rc = tc::make_tmap_f16(&tmQ, d->q, HD, q_rows, d->heads, d->ldq, d->hsq, AQ);
  if (rc) return rc;
  rc = tc::make_tmap_f16(&tmK, d->k, HD, (uint64_t)d->n_k * d->n_img, d->heads, d->ldk, d->hsk, AK);
  if (rc) return rc;
  rc = tc::make_tmap_f16(&tmV, d->v, HD, (uint64_t)d->n_k * d->n_img, d->heads, d->ldv, d->hsv, AK);
  if (rc) return rc;
  AttnParams p;
  p.n_img = d->n_img; p.heads = d->heads; p.n_q = d->n_q; p.n_k = d->n_k;
  p.q_tiles = (d->n_q + AQ - 1) / AQ;
  p.q_rows_per_img = d->q_shared ? 0 : d->n_q;
  p.out = reinterpret_cast<__half *>(d->out); p.ldo = d->ldo; p.out_img_stride = d->out_img_stride;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.lse2 = d->lse2;
  p.lse_stride = d->lse2_stride > 0 ? d->lse2_stride : d->n_q;
  const int variant = d->variant & 3;
  if (d->lse2 && variant == 1) return FOHO_E_ARG;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    FOHO_CUDA_TRY(cudaGetDevice(&dev));
    FOHO_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  if (variant == 1) {     // one query tile per CTA (the first version; kept for A/B measurements)
    FOHO_CUDA_TRY(cudaFuncSetAttribute(k_attn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    long long items = (long long)p.n_img * p.heads * p.q_tiles;
    int grid = (int)(items < sm_count ? items : sm_count);
    if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
    k_attn_fwd<<<grid, 256, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
  } else {
    long long items = (long long)p.n_img * p.heads * ((p.q_tiles + 1) / 2);
    int grid = (int)(items < sm_count ? items : sm_count);
    if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
    if (variant == 2) {   // P through shared memory (the version before the P-in-TMEM rewrite; kept for A/B measurements)
      FOHO_CUDA_TRY(cudaFuncSetAttribute(k_attn_fwd2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att2Cfg<0>::SMEM));
      k_attn_fwd2<0><<<grid, 384, Att2Cfg<0>::SMEM, st>>>(tmQ, tmK, tmV, p);
    } else {
      FOHO_CUDA_TRY(cudaFuncSetAttribute(k_attn_fwd2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att2Cfg<1>::SMEM));
      k_attn_fwd2<1><<<grid, 384, Att2Cfg<1>::SMEM, st>>>(tmQ, tmK, tmV, p);
    }
  }
  FOHO_LAUNCH_CHECK();
  return 0;
}
