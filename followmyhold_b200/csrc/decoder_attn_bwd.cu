// Hand-written tcgen05 attention ADJOINT for the latent -> SDF decoder (row f1): what `loss.backward()` runs through the
// self attention of the ShapeVAE transformer and the cross attention of the lattice queries
// (third_party_patches/hy3dgen/shapegen/pipelines.py:299,304 forward; :1590-1600 the backward), without ever writing
// a score matrix to memory.  Head dimension 64, no mask.
//
//   given  Q, K, V, dO, lse2 = log2 sum_k exp2(S) per query (kept by the forward kernel), delta = dO . O per query
//   S = scale log2(e) Q K^T     P = exp2(S - lse2)     dP = dO V^T     dS = P o (dP - delta)
//   dV = P^T dO        dK = scale dS^T Q        dQ = scale dS K
//
// Two kinds of work item share one persistent kernel, so nothing is accumulated across CTAs (no atomics; bit-identical
// from run to run):
//   kind 0 (dK, dV of one block of 128 keys):  (K, V) fixed, streams the query blocks (Q, dO)
//   kind 1 (dQ of one block of 128 queries):   (Q, dO) fixed, streams the key blocks (K, V)
// Both compute, per streamed block, T1 = Q K^T (scores) and T2 = dO V^T (dP) into TMEM, rows = queries, so the
// statistics are one (lse2, delta) pair per thread; 256 threads turn them into the fp16 tiles dS (and P for kind 0) in
// swizzled shared memory (two buffers: the products of block b read one while block b+1 is written), and the
// accumulators take  kind 1: dQ += dS K   |   kind 0: dK += dS^T Q, dV += P^T dO  -- the transposes are MN-major
// operand descriptors on the same tiles, nothing is moved.
//   warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
//   warps 4-7 columns 0-63 and warps 8-11 columns 64-127 of the 128 x 128 tile (one row per thread)
// T(block+1) is issued as soon as the threads hold T(block) in registers, so the tensor core computes the next scores
// while the exponentials of this block run.
#include "foho_common.cuh"
#include "foho_tc.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace {

constexpr int BT = 128;                    // rows of a tile (queries or keys)
constexpr int HD = 64;                     // head dimension
constexpr int TILE_BYTES = BT * HD * 2;    // 16 KB
constexpr int PD_BYTES = BT * BT * 2;      // 32 KB
// F1, F2 | streamed stages | P / dS buffers: 2 buffers leave room for 2 stages, 1 buffer for 4 (both 224 KB)
constexpr int BWD_SMEM = 2 * TILE_BYTES + 2 * 2 * TILE_BYTES + 2 * 2 * PD_BYTES + 1024 + 256;
static_assert(BWD_SMEM <= 232448, "shared memory of k_attn_bwd");
constexpr uint32_t TB_T1 = 0, TB_T2 = 128, TB_A = 256, TB_B = 320, TB_COLS = 512;

struct BwdParams {
  int n_img, heads, n_q, n_k, q_tiles, k_tiles;
  float scale_log2, scale;
  const float *lse2; long long lse_stride;       // [n_img][heads][lse_stride]
  const float *delta; long long delta_stride;    // [n_img][heads][delta_stride]
  __half *dq, *dk, *dv;                          // [rows][heads][64] views: row stride, head stride (elements)
  long long lddq, hsdq, lddk, hsdk, lddv, hsdv;
  int dbg;                                       // measurement only (FOHO_ATTN_BWD_DBG): 1 skip the exponentials, 2 skip the accumulate products, 4 skip the score products
};

struct Item { int kind, img, h, tile, nsteps; };

__device__ __forceinline__ Item decode_item(const BwdParams &p, int item) {
  Item it;
  const int n0 = p.n_img * p.heads * p.k_tiles;
  if (item < n0) {
    it.kind = 0;
    it.img = item / (p.heads * p.k_tiles);
    const int r = item - it.img * p.heads * p.k_tiles;
    it.h = r / p.k_tiles; it.tile = r - it.h * p.k_tiles;
    it.nsteps = p.q_tiles;
  } else {
    item -= n0;
    it.kind = 1;
    it.img = item / (p.heads * p.q_tiles);
    const int r = item - it.img * p.heads * p.q_tiles;
    it.h = r / p.q_tiles; it.tile = r - it.h * p.q_tiles;
    it.nsteps = p.k_tiles;
  }
  return it;
}

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// NG element-wise warpgroups: each takes 128 / NG columns of the tile (2: 384 threads, 4: 640 threads); PDB buffers of the
// P / dS tiles (1: the tile is written after the products of the previous block have read it, four stages of streamed
// operands; 2: two stages)
template <int NG, int PDB>
__global__ void __launch_bounds__(128 + 128 * NG, 1)
k_attn_bwd(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
           const __grid_constant__ CUtensorMap tmdO, const BwdParams p) {
  constexpr int Y_STAGES = PDB == 2 ? 2 : 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sF = smem;                                        // F1, F2
  uint8_t *sY = sF + 2 * TILE_BYTES;                         // stage s: Y1 at s*2*TILE, Y2 right after
  uint8_t *sPD = sY + Y_STAGES * 2 * TILE_BYTES;             // buffer u: P tile (kind 0) at u*2*PD, dS tile right after
  uint64_t *bars = reinterpret_cast<uint64_t *>(sPD + PDB * 2 * PD_BYTES);
  uint64_t *f_full = bars, *f_empty = bars + 1, *t_full = bars + 2, *t_empty = bars + 3, *acc_full = bars + 4, *acc_empty = bars + 5;
  uint64_t *pd_full = bars + 6, *pd_empty = bars + 8;
  uint64_t *y_full = bars + 10, *y_empty = y_full + Y_STAGES;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(y_empty + Y_STAGES);

  const int warp = tc::warp_idx_uniform(), lane = threadIdx.x & 31;   // provably warp-uniform: see foho_tc.cuh
  const int n_items = p.n_img * p.heads * (p.k_tiles + (p.dq ? p.q_tiles : 0));

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV); tc::tma_prefetch_desc(&tmdO);
  }
  if (warp == 1 && lane == 0) {
    tc::mbar_init(f_full, 1); tc::mbar_init(f_empty, 1);
    tc::mbar_init(t_full, 1); tc::mbar_init(t_empty, 128 * NG);
    for (int i = 0; i < PDB; ++i) { tc::mbar_init(&pd_full[i], 128 * NG); tc::mbar_init(&pd_empty[i], 1); }
    tc::mbar_init(acc_full, 1); tc::mbar_init(acc_empty, 128 * NG);
    for (int i = 0; i < Y_STAGES; ++i) { tc::mbar_init(&y_full[i], 1); tc::mbar_init(&y_empty[i], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc<TB_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp < 4) {
    // NG == 4: 640 threads x 96 registers is the whole file already and the element-wise code fits in 96
    if (NG == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ---------------------------------------------------------- TMA producer (whole warp walks the loop, one lane issues)
      const bool leader = tc::elect_one();
      uint32_t g = 0, w = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
        const Item it = decode_item(p, item);
        const int frow = it.kind == 0 ? it.img * p.n_k + it.tile * BT : it.img * p.n_q + it.tile * BT;
        tc::mbar_wait(f_empty, (w & 1) ^ 1);
        if (leader) {
          tc::mbar_expect_tx(f_full, 2 * TILE_BYTES);
          tc::tma_load_3d(sF, it.kind == 0 ? &tmK : &tmQ, f_full, 0, frow, it.h);
          tc::tma_load_3d(sF + TILE_BYTES, it.kind == 0 ? &tmV : &tmdO, f_full, 0, frow, it.h);
        }
        __syncwarp();
        for (int s = 0; s < it.nsteps; ++s, ++g) {
          const uint32_t st = g % Y_STAGES, ph = (g / Y_STAGES) & 1;
          const int yrow = it.kind == 0 ? it.img * p.n_q + s * BT : it.img * p.n_k + s * BT;
          tc::mbar_wait(&y_empty[st], ph ^ 1);
          uint8_t *y1 = sY + st * 2 * TILE_BYTES;
          if (leader) {
            tc::mbar_expect_tx(&y_full[st], 2 * TILE_BYTES);
            tc::tma_load_3d(y1, it.kind == 0 ? &tmQ : &tmK, &y_full[st], 0, yrow, it.h);
            tc::tma_load_3d(y1 + TILE_BYTES, it.kind == 0 ? &tmdO : &tmV, &y_full[st], 0, yrow, it.h);
          }
          __syncwarp();
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------- MMA issuer (whole warp walks the loop, one lane issues)
      const bool leader = tc::elect_one();
      constexpr uint32_t idesc_t = tc::idesc_f16(BT, BT, 0, 0);      // T = A B^T        : both K-major
      constexpr uint32_t idesc_q = tc::idesc_f16(BT, HD, 0, 1);      // dQ += dS K       : dS K-major, K MN-major
      constexpr uint32_t idesc_k = tc::idesc_f16(BT, HD, 1, 1);      // dK += dS^T Q ... : the tile and Q / dO MN-major
      const uint32_t f_addr = tc::smem_u32(sF), pd_addr = tc::smem_u32(sPD);
      uint32_t g = 0, w = 0;
      auto issue_t = [&](uint32_t gg, int kind) {
        const uint32_t st = gg % Y_STAGES, ph = (gg / Y_STAGES) & 1;
        tc::mbar_wait(&y_full[st], ph);
        tc::mbar_wait(t_empty, (gg & 1) ^ 1);
        tc::tc_fence_after();
        const uint32_t y_addr = tc::smem_u32(sY + st * 2 * TILE_BYTES);
        // rows = queries: kind 0 streams them (A = Y), kind 1 holds them (A = F)
        const uint32_t a1 = kind == 0 ? y_addr : f_addr, b1 = kind == 0 ? f_addr : y_addr;
        if (leader) {
          if (!(p.dbg & 4)) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              tc::mma_f16_ss(tmem_base + TB_T1, tc::smem_desc_sw128(a1 + k * 32, 16, 1024), tc::smem_desc_sw128(b1 + k * 32, 16, 1024),
                             idesc_t, k != 0);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k)
              tc::mma_f16_ss(tmem_base + TB_T2, tc::smem_desc_sw128(a1 + TILE_BYTES + k * 32, 16, 1024),
                             tc::smem_desc_sw128(b1 + TILE_BYTES + k * 32, 16, 1024), idesc_t, k != 0);
          }
          tc::mma_commit(t_full);
        }
        __syncwarp();
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
        const Item it = decode_item(p, item);
        tc::mbar_wait(f_full, w & 1);
        issue_t(g, it.kind);
        for (int s = 0; s < it.nsteps; ++s, ++g) {
          if (s + 1 < it.nsteps) issue_t(g + 1, it.kind);
          else { if (leader) tc::mma_commit(f_empty); __syncwarp(); }   // every product that reads the fixed tiles has been issued
          const uint32_t st = g % Y_STAGES, u = PDB == 2 ? (g & 1) : 0, pph = PDB == 2 ? ((g >> 1) & 1) : (g & 1);
          tc::mbar_wait(&pd_full[u], pph);
          if (s == 0) tc::mbar_wait(acc_empty, (w & 1) ^ 1);
          tc::tc_fence_after();
          const uint32_t y_addr = tc::smem_u32(sY + st * 2 * TILE_BYTES);
          const uint32_t p_addr = pd_addr + u * 2 * PD_BYTES, d_addr = p_addr + PD_BYTES;
          if (leader) {
          if (p.dbg & 2) {
          } else if (it.kind == 1) {
#pragma unroll
            for (int k = 0; k < BT / 16; ++k)
              tc::mma_f16_ss(tmem_base + TB_A, tc::smem_desc_sw128(d_addr + (k >> 2) * (BT * 128) + (k & 3) * 32, 16, 1024),
                             tc::smem_desc_sw128(y_addr + k * 2048, BT * 128, 1024), idesc_q, (s | k) != 0);
          } else {
#pragma unroll
            for (int k = 0; k < BT / 16; ++k)
              tc::mma_f16_ss(tmem_base + TB_A, tc::smem_desc_sw128(d_addr + k * 2048, BT * 128, 1024),
                             tc::smem_desc_sw128(y_addr + k * 2048, BT * 128, 1024), idesc_k, (s | k) != 0);
#pragma unroll
            for (int k = 0; k < BT / 16; ++k)
              tc::mma_f16_ss(tmem_base + TB_B, tc::smem_desc_sw128(p_addr + k * 2048, BT * 128, 1024),
                             tc::smem_desc_sw128(y_addr + TILE_BYTES + k * 2048, BT * 128, 1024), idesc_k, (s | k) != 0);
          }
          tc::mma_commit(&y_empty[st]);
          tc::mma_commit(&pd_empty[u]);
          if (s + 1 == it.nsteps) tc::mma_commit(acc_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    if (NG == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ------------------------------------------------------------ 128 NG threads: scores -> P, dS; epilogue
    constexpr int COLS = BT / NG;                    // columns of the tile per thread: 64 | 32
    constexpr int ACOLS = HD / NG;                   // columns of an accumulator per thread: 32 | 16
    const int grp = (warp - 4) >> 2;                 // column group
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t t1_addr = tmem_base + TB_T1 + grp * COLS + lane_off, t2_addr = tmem_base + TB_T2 + grp * COLS + lane_off;
    // columns grp*COLS .. : 64-column block (grp*COLS)/64, 16-byte chunk ((grp*COLS)%64)/8 + cc of the 128-byte row
    const uint32_t pd_row = tc::smem_u32(sPD) + ((grp * COLS) >> 6) * (BT * 128) + row * 128;
    const int chunk0 = ((grp * COLS) & 63) >> 3;
    uint32_t g = 0, w = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++w) {
      const Item it = decode_item(p, item);
      const float *lse_p = p.lse2 + ((long long)it.img * p.heads + it.h) * p.lse_stride;
      const float *dl_p = p.delta + ((long long)it.img * p.heads + it.h) * p.delta_stride;
      // the statistics of this thread's query row: fixed for kind 1, one per streamed block for kind 0 (loaded one block
      // ahead).  Rows past the last query get lse2 = +inf: their P and dS are exactly zero.
      float lse_r = INFINITY, dl_r = 0.f, lse_n = INFINITY, dl_n = 0.f;
      {
        const int qrow = (it.kind == 1 ? it.tile * BT : 0) + row;
        if (qrow < p.n_q) { lse_r = __ldg(lse_p + qrow); dl_r = __ldg(dl_p + qrow); }
      }
      for (int s = 0; s < it.nsteps; ++s, ++g) {
        if (it.kind == 0) {
          const int qn = (s + 1) * BT + row;
          lse_n = INFINITY; dl_n = 0.f;
          if (s + 1 < it.nsteps && qn < p.n_q) { lse_n = __ldg(lse_p + qn); dl_n = __ldg(dl_p + qn); }
        }
        const uint32_t u = PDB == 2 ? (g & 1) : 0, pph = PDB == 2 ? ((g >> 1) & 1) : (g & 1);
        tc::mbar_wait(t_full, g & 1);
        tc::tc_fence_after();
        uint32_t a[COLS / 32][32], b[COLS / 32][32];
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) { tc::tmem_ld32(t1_addr + c * 32, a[c]); tc::tmem_ld32(t2_addr + c * 32, b[c]); }
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        tc::mbar_arrive(t_empty);
        const uint32_t p_row = pd_row + u * 2 * PD_BYTES, d_row = p_row + PD_BYTES;
        const float neg_l = -lse_r;
        const float2 sc2 = tc::splat2(p.scale_log2), nl2 = tc::splat2(neg_l), nd2 = tc::splat2(-dl_r);
        // values first (in place: a <- packed P, b <- packed dS), then the wait for the buffer, then the stores: the
        // products of the previous block run while this block's exponentials do
#pragma unroll
        for (int cc = 0; cc < COLS / 8; ++cc) {                // 16-byte chunk: columns grp*COLS + 8cc .. +7
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int e = (cc & 3) * 8 + 2 * t;
            // packed pairs (FFMA2 / FADD2 / FMUL2): 3 arithmetic instructions per two elements instead of 6
            const float2 x = tc::ffma2(make_float2(__uint_as_float(a[cc >> 2][e]), __uint_as_float(a[cc >> 2][e + 1])), sc2, nl2);
            const float2 pp = make_float2(ex2a(x.x), ex2a(x.y));
            const float2 ss = tc::fmul2(pp, tc::fadd2(make_float2(__uint_as_float(b[cc >> 2][e]), __uint_as_float(b[cc >> 2][e + 1])), nd2));
            const __half2 hp = __floats2half2_rn(pp.x, pp.y), hs = __floats2half2_rn(ss.x, ss.y);
            a[cc >> 2][(cc & 3) * 8 + t] = *reinterpret_cast<const uint32_t *>(&hp);      // slot 8(cc&3)+t <= e: already consumed
            b[cc >> 2][(cc & 3) * 8 + t] = *reinterpret_cast<const uint32_t *>(&hs);
          }
        }
        tc::mbar_wait(&pd_empty[u], pph ^ 1);                  // the products that read this buffer are done
        if (!(p.dbg & 1))
#pragma unroll
        for (int cc = 0; cc < COLS / 8; ++cc) {
          const int o = (cc & 3) * 8;
          const uint32_t off = (uint32_t)(((chunk0 + cc) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(d_row + off), "r"(b[cc >> 2][o]), "r"(b[cc >> 2][o + 1]),
                       "r"(b[cc >> 2][o + 2]), "r"(b[cc >> 2][o + 3]) : "memory");
          if (it.kind == 0)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + off), "r"(a[cc >> 2][o]), "r"(a[cc >> 2][o + 1]),
                         "r"(a[cc >> 2][o + 2]), "r"(a[cc >> 2][o + 3]) : "memory");
        }
        tc::fence_proxy_async_smem();
        tc::tc_fence_before();
        tc::mbar_arrive(&pd_full[u]);
        if (it.kind == 0) { lse_r = lse_n; dl_r = dl_n; }
      }
      // ---- epilogue: the accumulators of this item (ACOLS of the 64 columns per thread)
      tc::mbar_wait(acc_full, w & 1);
      tc::tc_fence_after();
      const int trow = it.tile * BT + row;
      const bool valid = trow < (it.kind == 0 ? p.n_k : p.n_q);
      auto store_acc = [&](uint32_t taddr, __half *dst, float mul) {
        uint32_t v[ACOLS];
        if constexpr (ACOLS == 32) tc::tmem_ld32(taddr, v); else tc::tmem_ld16(taddr, v);
        tc::tmem_ld_wait();
        if (!valid) return;
#pragma unroll
        for (int i = 0; i < ACOLS; i += 8) {
          __align__(16) __half2 hh[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            hh[t] = __floats2half2_rn(__uint_as_float(v[i + 2 * t]) * mul, __uint_as_float(v[i + 2 * t + 1]) * mul);
          *reinterpret_cast<uint4 *>(dst + i) = *reinterpret_cast<uint4 *>(hh);
        }
      };
      if (it.kind == 0) {
        store_acc(tmem_base + TB_A + grp * ACOLS + lane_off,
                  p.dk + ((long long)it.img * p.n_k + trow) * p.lddk + (long long)it.h * p.hsdk + grp * ACOLS, p.scale);
        store_acc(tmem_base + TB_B + grp * ACOLS + lane_off,
                  p.dv + ((long long)it.img * p.n_k + trow) * p.lddv + (long long)it.h * p.hsdv + grp * ACOLS, 1.f);
      } else {
        store_acc(tmem_base + TB_A + grp * ACOLS + lane_off,
                  p.dq + ((long long)it.img * p.n_q + trow) * p.lddq + (long long)it.h * p.hsdq + grp * ACOLS, p.scale);
      }
      tc::tc_fence_before();
      tc::mbar_arrive(acc_empty);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc<TB_COLS>(tmem_base);
  }
}

}  // namespace

extern "C" int foho_tc_attention_bwd(const foho_attn_bwd_desc *d, void *cuda_stream) {
  if (!d || !d->q || !d->k || !d->v || !d->d_out || !d->lse2 || !d->delta || !d->dk || !d->dv) return FOHO_E_NULL;
  if (d->n_img <= 0 || d->heads <= 0 || d->n_q <= 0 || d->n_k <= 0 || d->n_k % BT) return FOHO_E_SHAPE;
  if (d->ldq % 8 || d->ldk % 8 || d->ldv % 8 || d->lddo % 8 || d->hsq % 8 || d->hsk % 8 || d->hsv % 8 || d->hsdo % 8) return FOHO_E_ARG;
  if (d->lddq % 8 || d->lddk % 8 || d->lddv % 8 || d->hsdq % 8 || d->hsdk % 8 || d->hsdv % 8) return FOHO_E_ARG;
  if ((reinterpret_cast<uintptr_t>(d->dq) | reinterpret_cast<uintptr_t>(d->dk) | reinterpret_cast<uintptr_t>(d->dv)) & 15) return FOHO_E_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  CUtensorMap tmQ, tmK, tmV, tmdO;
  const uint64_t q_rows = (uint64_t)d->n_q * d->n_img, k_rows = (uint64_t)d->n_k * d->n_img;
  int rc = tc::make_tmap_f16(&tmQ, d->q, HD, q_rows, d->heads, d->ldq, d->hsq, BT);
  if (rc) return rc;
  if ((rc = tc::make_tmap_f16(&tmK, d->k, HD, k_rows, d->heads, d->ldk, d->hsk, BT))) return rc;
  if ((rc = tc::make_tmap_f16(&tmV, d->v, HD, k_rows, d->heads, d->ldv, d->hsv, BT))) return rc;
  if ((rc = tc::make_tmap_f16(&tmdO, d->d_out, HD, q_rows, d->heads, d->lddo, d->hsdo, BT))) return rc;
  BwdParams p;
  p.n_img = d->n_img; p.heads = d->heads; p.n_q = d->n_q; p.n_k = d->n_k;
  p.q_tiles = (d->n_q + BT - 1) / BT; p.k_tiles = d->n_k / BT;
  p.scale = d->scale; p.scale_log2 = d->scale * 1.4426950408889634f;
  p.lse2 = d->lse2; p.lse_stride = d->lse2_stride > 0 ? d->lse2_stride : d->n_q;
  p.delta = d->delta; p.delta_stride = d->delta_stride > 0 ? d->delta_stride : d->n_q;
  p.dq = reinterpret_cast<__half *>(d->dq); p.dk = reinterpret_cast<__half *>(d->dk); p.dv = reinterpret_cast<__half *>(d->dv);
  p.lddq = d->lddq; p.hsdq = d->hsdq; p.lddk = d->lddk; p.hsdk = d->hsdk; p.lddv = d->lddv; p.hsdv = d->hsdv;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    FOHO_CUDA_TRY(cudaGetDevice(&dev));
    FOHO_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long items = (long long)p.n_img * p.heads * (p.k_tiles + (p.dq ? p.q_tiles : 0));
  p.dbg = 0;
  int grid = (int)(items < sm_count ? items : sm_count);
  if (d->max_ctas > 0 && grid > d->max_ctas) grid = d->max_ctas;
  static int mode = 0, dbg = 0;
  if (!mode) {
    const char *db = getenv("FOHO_ATTN_BWD_DBG");
    dbg = db ? atoi(db) : 0;
    const char *e = getenv("FOHO_ATTN_BWD_VARIANT");      // measurement switch: <element-wise warpgroups><tile buffers>: 21, 22, 41 (default), 42
    mode = e ? atoi(e) : 41;
    if (mode != 21 && mode != 22 && mode != 42) mode = 41;
  }
  p.dbg = dbg;
#define FOHO_BWD_LAUNCH(NG, PDB)                                                                                              \
  do {                                                                                                                        \
    FOHO_CUDA_TRY(cudaFuncSetAttribute(k_attn_bwd<NG, PDB>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));          \
    k_attn_bwd<NG, PDB><<<grid, 128 + 128 * NG, BWD_SMEM, st>>>(tmQ, tmK, tmV, tmdO, p);                                      \
  } while (0)
  if (mode == 21) FOHO_BWD_LAUNCH(2, 1);
  else if (mode == 22) FOHO_BWD_LAUNCH(2, 2);
  else if (mode == 42) FOHO_BWD_LAUNCH(4, 2);
  else FOHO_BWD_LAUNCH(4, 1);
#undef FOHO_BWD_LAUNCH
  FOHO_LAUNCH_CHECK();
  return 0;
}
