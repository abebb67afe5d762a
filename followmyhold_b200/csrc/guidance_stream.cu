// Dense volume stream of the guidance evaluation (the HBM-bound kernel).
//
// Reads the object field S[B,D,D,D] exactly once and writes the dense gradient
// G = dE/dS exactly once.  Per voxel g (lattice index) with S<0 the occupancy-weighted
// second-moment term (DESIGN.md "a10v", volume form of `obj_verts_loss_3`,
// third_party_patches/hy3dgen/shapegen/pipelines.py:1570) contributes
//     G[g] = -(w_mom/N) * |y(g)|^2,   |y|^2 = kappa^2 |g|^2 + 2 e.g + f,
// and the kernel accumulates the moments  M0 = sum w, M1 = sum w g, M2 = sum w |g|^2,
// w = relu(-S); k_assemble turns them into the energy and the
// object-leaf gradients.  The sparse terms (hand voxels, vertex samples) are added to G
// afterwards by their own small kernels.
//
// Algorithmic traffic: 4 B read + 4 B written per voxel.
//
// Three code paths:
//   * k_stream_tma  : D power of two; 1-D bulk TMA (cp.async.bulk + mbarrier) global->smem,
//                     compute in place, bulk TMA smem->global.  6-stage ring per CTA.
//   * k_stream_ldg  : D power of two; 128-bit LDG/STG with streaming hints, 4 loads in
//                     flight per thread.
//   * k_stream_any  : any D (the reference's 65 and 385); scalar.
#include "foho_common.cuh"

namespace {

struct StreamCoef {
  float k2;        // -cN * kappa^2
  float ex, ey, ez;// -cN * 2 e
  float f;         // -cN * f
};

__device__ __forceinline__ StreamCoef load_coef(const FohoFrame &fr, float cN) {
  StreamCoef c;
  c.k2 = -cN * fr.kappa * fr.kappa;
  c.ex = -cN * 2.f * fr.e[0];
  c.ey = -cN * 2.f * fr.e[1];
  c.ez = -cN * 2.f * fr.e[2];
  c.f = -cN * fr.f;
  return c;
}

// where the per-sample frame comes from: the stream kernels derive their coefficients from the leaves
// themselves (thread 0 of each CTA), so they do not depend on k_prep and can run beside it.
struct StreamSrc {
  const float *theta, *T_h2m, *obj_center;
  float bound;
  unsigned long long *trace;
};
__device__ __forceinline__ StreamCoef stream_coef(const StreamSrc &src, int b, int D, float cN, StreamCoef *sc) {
  if (threadIdx.x == 0) {
    FohoFrame fr;
    foho_object_frame(src.theta + (size_t)b * 16, src.T_h2m + (size_t)b * 16, src.obj_center + (size_t)b * 3, src.bound, D, fr);
    *sc = load_coef(fr, cN);
  }
  __syncthreads();
  return *sc;
}

struct StreamAcc {
  float m0, m1x, m1y, m2xy, cnt;
  float az[4];   // per z-slot sum of w (z is a per-thread constant in the fast paths)
};

__device__ __forceinline__ void acc_init(StreamAcc &a) {
  a.m0 = a.m1x = a.m1y = a.m2xy = a.cnt = 0.f;
  a.az[0] = a.az[1] = a.az[2] = a.az[3] = 0.f;
}

// one float4 of a row (ix,iy) at z = z0..z0+3; cz[k] = k2*z^2 + ez*z precomputed
__device__ __forceinline__ float4 voxel4(float4 s, float fx, float fy, const StreamCoef &c, const float (&cz)[4],
                                         StreamAcc &a) {
  float r2 = fmaf(fy, fy, __fmul_rn(fx, fx));          // explicit: every instantiation rounds the same way
  float crow = fmaf(c.k2, r2, fmaf(c.ex, fx, fmaf(c.ey, fy, c.f)));
  float w0 = foho_relu(-s.x), w1 = foho_relu(-s.y), w2 = foho_relu(-s.z), w3 = foho_relu(-s.w);
  float4 g;
  g.x = s.x < 0.f ? cz[0] + crow : 0.f;
  g.y = s.y < 0.f ? cz[1] + crow : 0.f;
  g.z = s.z < 0.f ? cz[2] + crow : 0.f;
  g.w = s.w < 0.f ? cz[3] + crow : 0.f;
  a.az[0] += w0; a.az[1] += w1; a.az[2] += w2; a.az[3] += w3;
  float rs = (w0 + w1) + (w2 + w3);
  a.m0 += rs;
  a.m1x = fmaf(fx, rs, a.m1x);
  a.m1y = fmaf(fy, rs, a.m1y);
  a.m2xy = fmaf(r2, rs, a.m2xy);
  return g;
}

__device__ __forceinline__ void acc_store(const StreamAcc &a, float z0, float *part, float *red_smem) {
  float v[6];
  float m1z = 0.f, m2z = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float z = z0 + (float)k;
    m1z = fmaf(z, a.az[k], m1z);
    m2z = fmaf(z * z, a.az[k], m2z);
  }
  v[0] = a.m0; v[1] = a.m1x; v[2] = a.m1y; v[3] = m1z; v[4] = a.m2xy + m2z; v[5] = a.cnt;
  block_sum<6>(v, red_smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) part[i] = v[i];
    part[6] = 0.f; part[7] = 0.f;
  }
}

// ------------------------------------------------------------------------- LDG path
constexpr int LDG_THREADS = 256;
constexpr int LDG_UNROLL = 4;

__global__ void __launch_bounds__(LDG_THREADS) k_stream_ldg(const float *__restrict__ sdf, float *__restrict__ grad,
                                                            const StreamSrc src,
                                                            float *__restrict__ partials, int D, int logD, float cN) {
  FohoTrace trace_(src.trace, TR_STREAM);
  __shared__ float red[6 * 32];
  __shared__ StreamCoef sc;
  const int b = blockIdx.y;
  const StreamCoef c = stream_coef(src, b, D, cN, &sc);
  const size_t vol = (size_t)D * D * D;
  const float4 *__restrict__ S4 = reinterpret_cast<const float4 *>(sdf + (size_t)b * vol);
  float4 *__restrict__ G4 = reinterpret_cast<float4 *>(grad + (size_t)b * vol);
  const int log4 = logD - 2;                       // float4 per row = 1 << log4
  const long long total4 = (long long)(vol >> 2);
  const long long tile4 = (long long)LDG_THREADS * LDG_UNROLL;
  const long long ntiles = (total4 + tile4 - 1) / tile4;
  const int row4mask = (1 << log4) - 1;
  // when 256 % row4 == 0 (D <= 1024) the thread's z slot is the same for every load
  const float z0 = (float)((threadIdx.x & row4mask) << 2);
  float cz[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { float z = z0 + (float)k; cz[k] = fmaf(c.k2 * z, z, c.ez * z); }
  const bool fast = logD >= 6 && log4 <= 8;         // see k_stream_tma
  const int rows_per_tile = (int)tile4 >> log4;
  float dy[LDG_UNROLL];
#pragma unroll
  for (int u = 0; u < LDG_UNROLL; ++u) dy[u] = (float)((u * LDG_THREADS) >> log4);
  StreamAcc a; acc_init(a);
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    long long base = t * tile4 + threadIdx.x;
    float4 s[LDG_UNROLL];
    if (fast) {
#pragma unroll
      for (int u = 0; u < LDG_UNROLL; ++u) s[u] = __ldcs(S4 + base + (long long)u * LDG_THREADS);
      const int first_row = (int)t * rows_per_tile;
      const float fx = (float)(first_row >> logD);
      const float fyb = (float)((first_row & (D - 1)) + (threadIdx.x >> log4));
#pragma unroll
      for (int u = 0; u < LDG_UNROLL; ++u) __stcs(G4 + base + (long long)u * LDG_THREADS, voxel4(s[u], fx, fyb + dy[u], c, cz, a));
      continue;
    }
#pragma unroll
    for (int u = 0; u < LDG_UNROLL; ++u) {
      long long i = base + (long long)u * LDG_THREADS;
      s[u] = i < total4 ? __ldcs(S4 + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
#pragma unroll
    for (int u = 0; u < LDG_UNROLL; ++u) {
      long long i = base + (long long)u * LDG_THREADS;
      if (i < total4) {
        int row = (int)(i >> log4);
        float fx = (float)(row >> logD), fy = (float)(row & (D - 1));
        float4 g = voxel4(s[u], fx, fy, c, cz, a);
        __stcs(G4 + i, g);
      }
    }
  }
  acc_store(a, z0, partials + ((size_t)b * FOHO_MAX_STREAM_CTAS + blockIdx.x) * FOHO_STREAM_PARTIALS, red);
}

// ------------------------------------------------------------------------- TMA path
constexpr int TMA_THREADS = 256;
constexpr int TMA_F4_PER_THREAD = 4;                                  // 16 KB tiles
constexpr int TMA_TILE_F4 = TMA_THREADS * TMA_F4_PER_THREAD;          // 1024 float4
constexpr int TMA_TILE_BYTES = TMA_TILE_F4 * 16;
constexpr int TMA_MAX_STAGES = 13;   // ring depth and prefetch distance are launch parameters

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(TMA_THREADS, 2) k_stream_tma(const float *__restrict__ sdf, float *__restrict__ grad,
                                                               const StreamSrc src,
                                                               float *__restrict__ partials, int D, int logD, float cN,
                                                               int nstages, int nprefetch) {
  FohoTrace trace_(src.trace, TR_STREAM);
  __shared__ StreamCoef sc;
  __shared__ uint64_t full[TMA_MAX_STAGES];
  __shared__ float red[6 * 32];
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float4 *stage = reinterpret_cast<float4 *>(smem_raw);                       // [nstages][TILE_F4]

  const int b = blockIdx.y;
  const StreamCoef c = stream_coef(src, b, D, cN, &sc);
  const size_t vol = (size_t)D * D * D;
  const float4 *S4 = reinterpret_cast<const float4 *>(sdf + (size_t)b * vol);
  float4 *G4 = reinterpret_cast<float4 *>(grad + (size_t)b * vol);
  const int log4 = logD - 2;
  const long long total4 = (long long)(vol >> 2);
  const long long ntiles = (total4 + TMA_TILE_F4 - 1) / TMA_TILE_F4;
  const int row4mask = (1 << log4) - 1;
  const float z0 = (float)((threadIdx.x & row4mask) << 2);
  float cz[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { float z = z0 + (float)k; cz[k] = fmaf(c.k2 * z, z, c.ez * z); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // my tiles: t_k = blockIdx.x + k * gridDim.x
  const long long first = blockIdx.x;
  const long long nmine = first < ntiles ? (ntiles - first + gridDim.x - 1) / gridDim.x : 0;
  auto tile_f4 = [&](long long k) -> int {
    long long t = first + k * gridDim.x;
    long long rem = total4 - t * TMA_TILE_F4;
    return (int)(rem < TMA_TILE_F4 ? rem : TMA_TILE_F4);
  };
  if (threadIdx.x == 0) {
    for (int k = 0; k < nprefetch && k < nmine; ++k) {
      int n4 = tile_f4(k);
      mbar_expect_tx(&full[k], (uint32_t)n4 * 16u);
      bulk_load(stage + (size_t)k * TMA_TILE_F4, S4 + (first + (long long)k * gridDim.x) * TMA_TILE_F4, (uint32_t)n4 * 16u,
                &full[k]);
    }
  }
  const bool fast = logD >= 6 && log4 <= 8;
  const int rows_per_tile = TMA_TILE_F4 >> log4;
  float dy[TMA_F4_PER_THREAD];
#pragma unroll
  for (int u = 0; u < TMA_F4_PER_THREAD; ++u) dy[u] = (float)((u * TMA_THREADS) >> log4);
  StreamAcc a; acc_init(a);
  int s = 0, sn = nprefetch % nstages;     // ring slots of tile k and of tile k + nprefetch
  uint32_t ph = 0;
  const int pending = nstages - nprefetch; // store groups that may stay in flight when a slot is refilled
  for (long long k = 0; k < nmine; ++k) {
    const long long t = first + k * gridDim.x;
    const int n4 = tile_f4(k);
    float4 *buf = stage + (size_t)s * TMA_TILE_F4;
    mbar_wait(&full[s], ph);
    if (fast) {
      // D >= 64: a tile (4096 voxels) lies inside one x-slab and is never partial, so x is a per-tile
      // constant and y a per-thread base plus a per-slot constant -- same float values, same arithmetic
      // as the general path below (bit-identical results), without the per-float4 index math
      const int first_row = (int)t * rows_per_tile;
      const float fx = (float)(first_row >> logD);
      const float fyb = (float)((first_row & (D - 1)) + (threadIdx.x >> log4));
#pragma unroll
      for (int u = 0; u < TMA_F4_PER_THREAD; ++u) {
        const int f = threadIdx.x + u * TMA_THREADS;
        buf[f] = voxel4(buf[f], fx, fyb + dy[u], c, cz, a);
      }
    } else {
#pragma unroll
      for (int u = 0; u < TMA_F4_PER_THREAD; ++u) {
        int f = threadIdx.x + u * TMA_THREADS;
        if (f < n4) {
          long long i = t * TMA_TILE_F4 + f;
          int row = (int)(i >> log4);
          float fx = (float)(row >> logD), fy = (float)(row & (D - 1));
          buf[f] = voxel4(buf[f], fx, fy, c, cz, a);
        }
      }
    }
    fence_proxy_async();       // make the generic-proxy smem writes visible to the bulk store
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(G4 + t * TMA_TILE_F4, buf, (uint32_t)n4 * 16u);
      long long kn = k + nprefetch;
      if (kn < nmine) {
        // slot sn was last stored from at iteration kn - nstages = k - pending; the `pending`
        // younger store groups may stay in flight.
        if (pending >= 4) bulk_wait_read<4>();
        else if (pending == 3) bulk_wait_read<3>();
        else if (pending == 2) bulk_wait_read<2>();
        else bulk_wait_read<1>();
        int nn4 = tile_f4(kn);
        mbar_expect_tx(&full[sn], (uint32_t)nn4 * 16u);
        bulk_load(stage + (size_t)sn * TMA_TILE_F4, S4 + (first + kn * gridDim.x) * TMA_TILE_F4, (uint32_t)nn4 * 16u,
                  &full[sn]);
      }
    }
    if (++s == nstages) { s = 0; ph ^= 1u; }
    if (++sn == nstages) sn = 0;
  }
  if (threadIdx.x == 0) bulk_wait_all();
  acc_store(a, z0, partials + ((size_t)b * FOHO_MAX_STREAM_CTAS + blockIdx.x) * FOHO_STREAM_PARTIALS, red);
}

// ------------------------------------------------------------------------- generic path
__global__ void __launch_bounds__(256) k_stream_any(const float *__restrict__ sdf, float *__restrict__ grad,
                                                    const StreamSrc src, float *__restrict__ partials,
                                                    int D, float cN) {
  FohoTrace trace_(src.trace, TR_STREAM);
  __shared__ float red[6 * 32];
  __shared__ StreamCoef sc;
  const int b = blockIdx.y;
  const StreamCoef c = stream_coef(src, b, D, cN, &sc);
  const long long vol = (long long)D * D * D;
  const float *__restrict__ S = sdf + (size_t)b * vol;
  float *__restrict__ G = grad + (size_t)b * vol;
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vol; i += stride) {
    int iz = (int)(i % D);
    long long r = i / D;
    int iy = (int)(r % D), ix = (int)(r / D);
    float fx = (float)ix, fy = (float)iy, fz = (float)iz;
    float s = __ldcs(S + i);
    float r2 = fx * fx + fy * fy;
    float crow = fmaf(c.k2, r2, fmaf(c.ex, fx, fmaf(c.ey, fy, c.f)));
    float czz = fmaf(c.k2 * fz, fz, c.ez * fz);
    float w = foho_relu(-s);
    __stcs(G + i, s < 0.f ? czz + crow : 0.f);
    v[0] += w;
    v[1] = fmaf(fx, w, v[1]);
    v[2] = fmaf(fy, w, v[2]);
    v[3] = fmaf(fz, w, v[3]);
    v[4] = fmaf(r2 + fz * fz, w, v[4]);
  }
  block_sum<6>(v, red);
  if (threadIdx.x == 0) {
    float *part = partials + ((size_t)b * FOHO_MAX_STREAM_CTAS + blockIdx.x) * FOHO_STREAM_PARTIALS;
    for (int i = 0; i < 6; ++i) part[i] = v[i];
    part[6] = part[7] = 0.f;
  }
}

// Generic path, vectorised: the volume of image b is one flat array; a 16-byte aligned body is streamed as float4
// (head / tail of at most 3 voxels scalar, by one thread), lattice coordinates come from one magic-number division
// per float4 and a carry for the other three voxels -- no per-voxel div / mod (the reference's lattices are odd:
// 65^3, 385^3).
__global__ void __launch_bounds__(256) k_stream_any4(const float *__restrict__ sdf, float *__restrict__ grad,
                                                     const StreamSrc src, float *__restrict__ partials,
                                                     int D, float cN, unsigned long long magic) {
  FohoTrace trace_(src.trace, TR_STREAM);
  __shared__ float red[6 * 32];
  __shared__ StreamCoef sc;
  const int b = blockIdx.y;
  const StreamCoef c = stream_coef(src, b, D, cN, &sc);
  const long long vol = (long long)D * D * D;
  const float *__restrict__ S = sdf + (size_t)b * vol;
  float *__restrict__ G = grad + (size_t)b * vol;
  float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto one = [&](long long i, int ix, int iy, int iz, float s) -> float {
    const float fx = (float)ix, fy = (float)iy, fz = (float)iz;
    const float r2 = fx * fx + fy * fy;
    const float crow = fmaf(c.k2, r2, fmaf(c.ex, fx, fmaf(c.ey, fy, c.f)));
    const float czz = fmaf(c.k2 * fz, fz, c.ez * fz);
    const float w = foho_relu(-s);
    v[0] += w;
    v[1] = fmaf(fx, w, v[1]);
    v[2] = fmaf(fy, w, v[2]);
    v[3] = fmaf(fz, w, v[3]);
    v[4] = fmaf(r2 + fz * fz, w, v[4]);
    (void)i;
    return s < 0.f ? czz + crow : 0.f;
  };
  auto coords = [&](long long i, int &ix, int &iy, int &iz) {
    const unsigned long long r = ((unsigned long long)i * magic) >> 40;     // i / D, exact for i < 2^40 / D
    iz = (int)(i - (long long)r * D);
    const unsigned long long q = (r * magic) >> 40;
    ix = (int)q; iy = (int)(r - q * D);
  };
  const int head = (int)((4 - ((reinterpret_cast<uintptr_t>(S) >> 2) & 3)) & 3);          // voxels before the aligned body
  const long long n4 = (vol - head) >> 2;
  const long long tail0 = head + (n4 << 2);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long i = 0; i < head && i < vol; ++i) { int ix, iy, iz; coords(i, ix, iy, iz); G[i] = one(i, ix, iy, iz, S[i]); }
    for (long long i = tail0; i < vol; ++i) { int ix, iy, iz; coords(i, ix, iy, iz); G[i] = one(i, ix, iy, iz, S[i]); }
  }
  const float4 *__restrict__ S4 = reinterpret_cast<const float4 *>(S + head);
  float4 *__restrict__ G4 = reinterpret_cast<float4 *>(G + head);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += stride) {
    const long long i = head + (k << 2);
    int ix, iy, iz;
    coords(i, ix, iy, iz);
    const float4 s = __ldcs(S4 + k);
    const float sv[4] = {s.x, s.y, s.z, s.w};
    float g[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      g[t] = one(i + t, ix, iy, iz, sv[t]);
      if (++iz == D) { iz = 0; if (++iy == D) { iy = 0; ++ix; } }
    }
    __stcs(G4 + k, make_float4(g[0], g[1], g[2], g[3]));
  }
  block_sum<6>(v, red);
  if (threadIdx.x == 0) {
    float *part = partials + ((size_t)b * FOHO_MAX_STREAM_CTAS + blockIdx.x) * FOHO_STREAM_PARTIALS;
    for (int i = 0; i < 6; ++i) part[i] = v[i];
    part[6] = part[7] = 0.f;
  }
}

int ilog2_exact(int D) {
  int l = 0;
  while ((1 << l) < D) ++l;
  return (1 << l) == D ? l : -1;
}

}  // namespace

int foho_launch_stream(const foho_guidance_desc *d, const FohoWorkspace &ws, int *grid_x_out, bool /*shared_sm*/, cudaStream_t st) {
  FohoDeviceState *ds = foho_device_state();
  if (!ds) return (int)cudaGetLastError();
  const int sm_count = ds->sm_count;
  const int D = d->D, B = d->B;
  const double N = (double)D * D * D;
  const float cN = (float)((double)d->w.w_mom / N);
  const int logD = ilog2_exact(D);
  int variant = d->stream_variant;
  StreamSrc src;
  src.theta = d->theta; src.T_h2m = d->T_h2m; src.obj_center = d->obj_center; src.bound = d->bound;
  src.trace = ws.trace;
  const bool pow2 = logD >= 3 && D <= 1024;       // needs >= 2 float4 per row
  if (!pow2) variant = 3;
  else if (variant == 0) variant = 2;
  int gx;
  if (variant == 2) {
    // Launch shape (measured on B200, scripts/stage_times.py): what matters is ~64 KB of bulk loads in
    // flight per SM -- 48 KB or 96 KB are 7-13 % slower, 32 KB 29 % slower -- not how many CTAs issue
    // them.  One persistent CTA per SM with a 6-slot ring and 4 loads in flight reaches that with 96 KB
    // of shared memory and 256 threads, leaving the rest of the SM to the sparse kernels that run
    // beside the stream (d->stream_* override for experiments).
    const int ctas = d->stream_ctas > 0 ? d->stream_ctas : 1;
    int nstages = d->stream_stages > 0 ? d->stream_stages : 6;
    if (nstages < 2) nstages = 2;
    if (nstages > TMA_MAX_STAGES) nstages = TMA_MAX_STAGES;
    int nprefetch = d->stream_prefetch > 0 ? d->stream_prefetch : (d->stream_stages > 0 ? nstages / 2 : 4);
    if (nprefetch >= nstages) nprefetch = nstages - 1;
    const size_t smem = (size_t)nstages * TMA_TILE_BYTES;
    // the largest ring is opted into once; the max-shared carve-out lets kernels running beside the stream
    // find room for their own shared memory without waiting for the SM to drain
    {
      int rc = foho_func_attrs((const void *)k_stream_tma, FA_STREAM_TMA, (size_t)TMA_MAX_STAGES * TMA_TILE_BYTES, true);
      if (rc != FOHO_OK) return rc;
    }
    long long ntiles = ((long long)(N / 4) + TMA_TILE_F4 - 1) / TMA_TILE_F4;
    gx = sm_count * ctas / B;                     // at most `ctas` resident CTAs per SM over the whole batch:
                                                  // one wave, never a straggler CTA waiting for a free SM
    if (gx > ntiles) gx = (int)ntiles;
    if (gx > FOHO_MAX_STREAM_CTAS) gx = FOHO_MAX_STREAM_CTAS;
    if (gx < 1) gx = 1;
    k_stream_tma<<<dim3(gx, B), TMA_THREADS, smem, st>>>(d->sdf, d->grad_sdf, src, ws.stream_part, D, logD, cN, nstages,
                                                             nprefetch);
  } else if (variant == 1) {
    long long ntiles = ((long long)(N / 4) + LDG_THREADS * LDG_UNROLL - 1) / (LDG_THREADS * LDG_UNROLL);
    gx = (sm_count * 8 + B - 1) / B;
    if (gx > ntiles) gx = (int)ntiles;
    if (gx > FOHO_MAX_STREAM_CTAS) gx = FOHO_MAX_STREAM_CTAS;
    if (gx < 1) gx = 1;
    k_stream_ldg<<<dim3(gx, B), LDG_THREADS, 0, st>>>(d->sdf, d->grad_sdf, src, ws.stream_part, D, logD, cN);
  } else {
    long long nblk = ((long long)N + 255) / 256;
    gx = (sm_count * 8 + B - 1) / B;
    if (gx > nblk) gx = (int)nblk;
    if (gx > FOHO_MAX_STREAM_CTAS) gx = FOHO_MAX_STREAM_CTAS;
    if (gx < 1) gx = 1;
    if (((reinterpret_cast<uintptr_t>(d->sdf) | reinterpret_cast<uintptr_t>(d->grad_sdf)) & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(d->sdf) & 15) == (reinterpret_cast<uintptr_t>(d->grad_sdf) & 15) && D >= 4) {
      const unsigned long long magic = ((1ull << 40) + (unsigned long long)D - 1) / (unsigned long long)D;
      k_stream_any4<<<dim3(gx, B), 256, 0, st>>>(d->sdf, d->grad_sdf, src, ws.stream_part, D, cN, magic);
    } else {
      k_stream_any<<<dim3(gx, B), 256, 0, st>>>(d->sdf, d->grad_sdf, src, ws.stream_part, D, cN);
    }
  }
  FOHO_LAUNCH_CHECK();
  *grid_x_out = gx;
  return FOHO_OK;
}
