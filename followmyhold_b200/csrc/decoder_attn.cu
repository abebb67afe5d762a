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

}  // namespace

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
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    FOHO_CUDA_TRY(cudaGetDevice(&dev));
    FOHO_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  FOHO_CUDA_TRY(cudaFuncSetAttribute(k_attn_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
  long long items = (long long)p.n_img * p.heads * p.q_tiles;
  int grid = (int)(items < sm_count ? items : sm_count);
  if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
  k_attn_fwd<<<grid, 256, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
  FOHO_LAUNCH_CHECK();
  return 0;
}
