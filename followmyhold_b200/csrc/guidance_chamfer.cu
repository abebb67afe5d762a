// Exact 1-NN chamfer between the transformed hand vertices and the MoGe cloud (NS row a15; same
// semantics as the reference's pytorch3d knn_points(K=1) call of
// third_party_patches/hy3dgen/shapegen/pipelines.py:1529-1532: squared L2 to the single nearest
// neighbour), with search structures that are built ONCE per image and reused by every one of the
// ~1200 evaluations of a guided diffusion:
//
//   cloud -> hand : the hand moves by a similarity, so "nearest transformed hand vertex of c" is
//                   "nearest REST vertex of c pulled back into the rest frame".  The rest vertices are
//                   Morton-sorted once into leaves of 8 with tight AABBs and super-boxes of 8 leaves;
//                   each cloud point descends that two-level hierarchy with its running best as the
//                   pruning radius (k_chamfer_c2h).  The cloud is stored cell-sorted, so the 32 points
//                   of a warp are neighbours and take the same branches.
//   hand -> cloud : the cloud never moves: it is binned once into a 32^3 Morton-ordered uniform grid
//                   (CSR); each hand vertex scans Chebyshev rings of cells until its best distance is
//                   below the ring radius (k_chamfer_h2c, one warp per vertex).
//
// Both are exact (not approximate) searches; the brute-force k_chamfer in guidance_sparse.cu remains
// as the path used when the caller passes no accel buffer, and as the cross-check in the tests.
#include "foho_common.cuh"

namespace {

constexpr int ACC_THREADS = 1024;

// ----------------------------------------------------------------------------- build: hand
__device__ __forceinline__ unsigned int spread3(unsigned int v) {   // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(ACC_THREADS) k_accel_hand(const float *__restrict__ hand_rest, int Vh, FohoAccel acc) {
  __shared__ unsigned long long key[FOHO_ACCEL_HV];
  __shared__ float red[6 * 32];
  __shared__ float bb[6];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float *rest = hand_rest + (size_t)b * Vh * 3;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < Vh; i += blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a) { float v = rest[3 * i + a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = warp_min(mn[a]); mx[a] = warp_max(mx[a]); }
  if (lane == 0)
#pragma unroll
    for (int a = 0; a < 3; ++a) { red[a * 32 + wid] = mn[a]; red[(3 + a) * 32 + wid] = mx[a]; }
  __syncthreads();
  if (tid < 6) {
    float v = red[tid * 32];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) v = tid < 3 ? fminf(v, red[tid * 32 + w]) : fmaxf(v, red[tid * 32 + w]);
    bb[tid] = v;
  }
  __syncthreads();
  // same bbox-centre arithmetic as k_prep (pipelines.py:111)
  const float ch[3] = {(bb[0] + bb[3]) / 2.0f, (bb[1] + bb[4]) / 2.0f, (bb[2] + bb[5]) / 2.0f};
  const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), fmaxf(bb[5] - bb[2], 1e-30f));
  const float q = 1023.f / ext;
  for (int i = tid; i < FOHO_ACCEL_HV; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < Vh) {
      unsigned int ix = (unsigned int)fminf(fmaxf((rest[3 * i] - bb[0]) * q, 0.f), 1023.f);
      unsigned int iy = (unsigned int)fminf(fmaxf((rest[3 * i + 1] - bb[1]) * q, 0.f), 1023.f);
      unsigned int iz = (unsigned int)fminf(fmaxf((rest[3 * i + 2] - bb[2]) * q, 0.f), 1023.f);
      unsigned int m = (spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz);
      k = ((unsigned long long)m << 32) | (unsigned int)i;      // unique keys: deterministic order
    }
    key[i] = k;
  }
  __syncthreads();
  // bitonic sort of FOHO_ACCEL_HV keys (one element per thread)
  for (int size = 2; size <= FOHO_ACCEL_HV; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int i = tid, j = i ^ stride;
      if (j > i) {
        const bool up = (i & size) == 0;
        unsigned long long a = key[i], c = key[j];
        if ((a > c) == up) { key[i] = c; key[j] = a; }
      }
      __syncthreads();
    }
  FohoAccelHand *H = acc.hand + b;
  const int nleaf = (Vh + 7) >> 3, nsup = (nleaf + 7) >> 3;
  // sorted vertices relative to the bbox centre; pad slots repeat the last real vertex
  for (int i = tid; i < nsup * 64; i += blockDim.x) {
    const int src = (int)(unsigned int)key[i < Vh ? i : Vh - 1];
    H->v[i] = make_float4(rest[3 * src] - ch[0], rest[3 * src + 1] - ch[1], rest[3 * src + 2] - ch[2], __int_as_float(src));
  }
  __syncthreads();
  for (int l = tid; l < nsup * 8; l += blockDim.x) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (l < nleaf)
      for (int k = 0; k < 8; ++k) {
        const float4 p = H->v[l * 8 + k];
        lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
        hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
      }
    H->leaf_lo[l] = make_float4(lo[0], lo[1], lo[2], 0.f);     // empty pad leaves: lo=+inf => never visited
    H->leaf_hi[l] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
  __syncthreads();
  for (int s = tid; s < FOHO_ACCEL_SUPERS; s += blockDim.x) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (s < nsup)
      for (int k = 0; k < 8; ++k) {
        const float4 a = H->leaf_lo[s * 8 + k], c = H->leaf_hi[s * 8 + k];
        lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
        hi[0] = fmaxf(hi[0], c.x); hi[1] = fmaxf(hi[1], c.y); hi[2] = fmaxf(hi[2], c.z);
      }
    H->sup_lo[s] = make_float4(lo[0], lo[1], lo[2], 0.f);
    H->sup_hi[s] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
  if (tid == 0) { H->n_leaves = nleaf; H->n_supers = nsup; H->Vh = Vh; H->pad = 0; }
}

// ----------------------------------------------------------------------------- build: cloud grid
__device__ __forceinline__ unsigned int spread5(unsigned int v) {   // 5 bits -> every third bit
  v &= 0x1fu;
  v = (v | (v << 8)) & 0x0000100Fu;
  v = (v | (v << 4)) & 0x000100C3u;
  v = (v | (v << 2)) & 0x00001249u;
  return v;
}
__device__ __forceinline__ int cell_code(int ix, int iy, int iz) {
  return (int)((spread5((unsigned)ix) << 2) | (spread5((unsigned)iy) << 1) | spread5((unsigned)iz));
}
__device__ __forceinline__ int cell_coord(float x, float origin, float inv_cell, int dim) {
  int c = (int)floorf((x - origin) * inv_cell);
  return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

__global__ void __launch_bounds__(ACC_THREADS) k_accel_cloud_bbox(const float *__restrict__ cloud, int P, FohoAccel acc) {
  __shared__ float red[6 * 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const float *c = cloud + (size_t)b * P * 3;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < P; i += blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; ++a) { float v = c[3 * (size_t)i + a]; mn[a] = fminf(mn[a], v); mx[a] = fmaxf(mx[a], v); }
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = warp_min(mn[a]); mx[a] = warp_max(mx[a]); }
  if (lane == 0)
#pragma unroll
    for (int a = 0; a < 3; ++a) { red[a * 32 + wid] = mn[a]; red[(3 + a) * 32 + wid] = mx[a]; }
  __syncthreads();
  if (tid == 0) {
    float lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
      lo[a] = red[a * 32]; hi[a] = red[(3 + a) * 32];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo[a] = fminf(lo[a], red[a * 32 + w]); hi[a] = fmaxf(hi[a], red[(3 + a) * 32 + w]); }
    }
    FohoAccelGrid *G = acc.grid + b;
    const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), fmaxf(hi[2] - lo[2], 1e-20f));
    const float cell = ext * (1.0f + 1e-5f) / (float)FOHO_ACCEL_G;     // cubic cells
    G->cell = cell; G->inv_cell = 1.0f / cell;
    for (int a = 0; a < 3; ++a) {
      G->origin[a] = lo[a];
      int dmax = (int)floorf((hi[a] - lo[a]) / cell) + 1;
      G->dims[a] = dmax < 1 ? 1 : (dmax > FOHO_ACCEL_G ? FOHO_ACCEL_G : dmax);
    }
    G->P = P; G->pad = 0;
  }
  // zero the per-cell counters of this sample
  int *cnt = acc.cell_fill + (size_t)b * FOHO_ACCEL_CELLS;
  for (int i = tid; i < FOHO_ACCEL_CELLS; i += blockDim.x) cnt[i] = 0;
}

__global__ void __launch_bounds__(256) k_accel_cloud_count(const float *__restrict__ cloud, int P, FohoAccel acc) {
  const int b = blockIdx.y;
  const FohoAccelGrid G = acc.grid[b];
  const float *c = cloud + (size_t)b * P * 3;
  int *cnt = acc.cell_fill + (size_t)b * FOHO_ACCEL_CELLS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const int code = cell_code(cell_coord(c[3 * (size_t)i], G.origin[0], G.inv_cell, G.dims[0]),
                               cell_coord(c[3 * (size_t)i + 1], G.origin[1], G.inv_cell, G.dims[1]),
                               cell_coord(c[3 * (size_t)i + 2], G.origin[2], G.inv_cell, G.dims[2]));
    atomicAdd(cnt + code, 1);
  }
}

// exclusive scan of the 32768 cell counts of one sample: 1024 threads x 32 consecutive cells
__global__ void __launch_bounds__(ACC_THREADS) k_accel_cloud_scan(FohoAccel acc) {
  __shared__ int wsum[32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int *cnt = acc.cell_fill + (size_t)b * FOHO_ACCEL_CELLS;
  int *start = acc.cell_start + (size_t)b * (FOHO_ACCEL_CELLS + 1);
  constexpr int PER = FOHO_ACCEL_CELLS / ACC_THREADS;
  int local[PER], sum = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) { local[k] = cnt[tid * PER + k]; sum += local[k]; }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int v = wsum[lane], inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    wsum[lane] = inc - v;
  }
  __syncthreads();
  int run = wsum[wid] + incl - sum;
#pragma unroll
  for (int k = 0; k < PER; ++k) { start[tid * PER + k] = run; run += local[k]; cnt[tid * PER + k] = 0; }
  if (tid == ACC_THREADS - 1) start[FOHO_ACCEL_CELLS] = run;
}

__global__ void __launch_bounds__(256) k_accel_cloud_scatter(const float *__restrict__ cloud, int P, FohoAccel acc) {
  const int b = blockIdx.y;
  const FohoAccelGrid G = acc.grid[b];
  const float *c = cloud + (size_t)b * P * 3;
  int *fill = acc.cell_fill + (size_t)b * FOHO_ACCEL_CELLS;
  const int *start = acc.cell_start + (size_t)b * (FOHO_ACCEL_CELLS + 1);
  float4 *pts = acc.pts + (size_t)b * P;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
    const float x = c[3 * (size_t)i], y = c[3 * (size_t)i + 1], z = c[3 * (size_t)i + 2];
    const int code = cell_code(cell_coord(x, G.origin[0], G.inv_cell, G.dims[0]), cell_coord(y, G.origin[1], G.inv_cell, G.dims[1]),
                               cell_coord(z, G.origin[2], G.inv_cell, G.dims[2]));
    const int slot = start[code] + atomicAdd(fill + code, 1);
    pts[slot] = make_float4(x, y, z, __int_as_float(i));
  }
}

// ----------------------------------------------------------------------------- cloud -> hand
constexpr int C2H_THREADS = 256;
constexpr int C2H_PER_THREAD = 4;
constexpr int C2H_POINTS = C2H_THREADS * C2H_PER_THREAD;

__device__ __forceinline__ float box_dist2(float4 lo, float4 hi, float qx, float qy, float qz) {
  const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

__global__ void __launch_bounds__(C2H_THREADS) k_chamfer_c2h(foho_guidance_desc d, FohoWorkspace ws, FohoAccel acc) {
  __shared__ float4 sv[FOHO_ACCEL_HV];
  __shared__ float4 s_llo[FOHO_ACCEL_LEAVES], s_lhi[FOHO_ACCEL_LEAVES];
  __shared__ float4 s_slo[FOHO_ACCEL_SUPERS], s_shi[FOHO_ACCEL_SUPERS];
  __shared__ float gacc[FOHO_ACCEL_HV * 3];
  __shared__ float red[32];
  const int b = blockIdx.y, Vh = d.Vh, P = d.P, tid = threadIdx.x;
  const int base = blockIdx.x * C2H_POINTS;
  if (base >= P) return;
  const FohoAccelHand *H = acc.hand + b;
  const int nleaf = H->n_leaves, nsup = H->n_supers;
  for (int i = tid; i < nsup * 64; i += blockDim.x) sv[i] = H->v[i];
  for (int i = tid; i < nsup * 8; i += blockDim.x) { s_llo[i] = H->leaf_lo[i]; s_lhi[i] = H->leaf_hi[i]; }
  if (tid < FOHO_ACCEL_SUPERS) { s_slo[tid] = H->sup_lo[tid]; s_shi[tid] = H->sup_hi[tid]; }
  for (int i = tid; i < Vh * 3; i += blockDim.x) gacc[i] = 0.f;
  __syncthreads();
  const FohoFrame &fr = ws.frames[b];
  // pull-back of a cloud point into the rest frame (relative to the rest bbox centre):
  //   q = R_h^T ((c - c_o) - (ch - c_o + t_h)) / s_h
  const float ox = fr.chc[0] + fr.th[0], oy = fr.chc[1] + fr.th[1], oz = fr.chc[2] + fr.th[2];
  const float is = 1.f / fr.sh;
  float Rt[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rt[3 * r + c] = fr.Rh[3 * c + r] * is;
  const float cx = fr.co[0], cy = fr.co[1], cz = fr.co[2];
  const float4 *pts = acc.pts + (size_t)b * P;
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;
  const float coef = 2.f * d.w.w_ch / (float)P;
  float sum = 0.f;
  int prev_leaf = 0;
#pragma unroll 1
  for (int u = 0; u < C2H_PER_THREAD; ++u) {
    const int i = base + u * C2H_THREADS + tid;
    if (i >= P) break;
    const float4 c4 = pts[i];
    const float px = c4.x - cx, py = c4.y - cy, pz = c4.z - cz;          // centred MoGe
    const float rx = px - ox, ry = py - oy, rz = pz - oz;
    const float qx = Rt[0] * rx + Rt[1] * ry + Rt[2] * rz;
    const float qy = Rt[3] * rx + Rt[4] * ry + Rt[5] * rz;
    const float qz = Rt[6] * rx + Rt[7] * ry + Rt[8] * rz;
    float best = INFINITY;
    int bj = prev_leaf * 8;
    // seed the pruning radius with the leaf that held the previous point's neighbour
    {
      const float4 *v = sv + prev_leaf * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 p = v[k];
        const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2 < best) { best = d2; bj = prev_leaf * 8 + k; }
      }
    }
    for (int s = 0; s < nsup; ++s) {
      if (!(box_dist2(s_slo[s], s_shi[s], qx, qy, qz) < best)) continue;
#pragma unroll 1
      for (int l = s * 8; l < s * 8 + 8; ++l) {
        if (l == prev_leaf || !(box_dist2(s_llo[l], s_lhi[l], qx, qy, qz) < best)) continue;
        const float4 *v = sv + l * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 p = v[k];
          const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
          const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
          if (d2 < best) { best = d2; bj = l * 8 + k; }
        }
      }
    }
    prev_leaf = bj >> 3;
    // exact squared distance in the (centred) MoGe frame, as the oracle evaluates it
    const int j = __float_as_int(sv[bj].w);
    const float hx = hmc[3 * j], hy = hmc[3 * j + 1], hz = hmc[3 * j + 2];
    const float dx = hx - px, dy = hy - py, dz = hz - pz;
    sum += fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    atomicAdd(gacc + 3 * j, coef * dx); atomicAdd(gacc + 3 * j + 1, coef * dy); atomicAdd(gacc + 3 * j + 2, coef * dz);
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < C2H_THREADS / 32; ++w) t += red[w];
    atomicAdd(ws.acc + (size_t)b * ACC_NUM + ACC_CH_CLOUD, t);
  }
  float *Ghm = ws.G_hm + (size_t)b * Vh * 3;
  for (int i = tid; i < Vh * 3; i += blockDim.x) {
    const float g = gacc[i];
    if (g != 0.f) atomicAdd(Ghm + i, g);
  }
  (void)nleaf;
}

// ----------------------------------------------------------------------------- hand -> cloud
constexpr int H2C_THREADS = 256;
constexpr int H2C_MAX_RING = 6;

__global__ void __launch_bounds__(H2C_THREADS) k_chamfer_h2c(foho_guidance_desc d, FohoWorkspace ws, FohoAccel acc) {
  const int b = blockIdx.y, Vh = d.Vh, P = d.P;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * (H2C_THREADS / 32) + wid;
  if (i >= Vh) return;
  const FohoAccelGrid G = acc.grid[b];
  const FohoFrame &fr = ws.frames[b];
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;
  // absolute MoGe position of this hand vertex (the grid lives in absolute coordinates)
  const float hx = hmc[3 * i] + fr.co[0], hy = hmc[3 * i + 1] + fr.co[1], hz = hmc[3 * i + 2] + fr.co[2];
  const int c0x = cell_coord(hx, G.origin[0], G.inv_cell, G.dims[0]);
  const int c0y = cell_coord(hy, G.origin[1], G.inv_cell, G.dims[1]);
  const int c0z = cell_coord(hz, G.origin[2], G.inv_cell, G.dims[2]);
  const int *start = acc.cell_start + (size_t)b * (FOHO_ACCEL_CELLS + 1);
  const float4 *pts = acc.pts + (size_t)b * P;
  float best = INFINITY;
  int bidx = -1;
  bool done = false;
  for (int r = 0; r <= H2C_MAX_RING && !done; ++r) {
    const int w = 2 * r + 1, n = w * w * w;
    for (int t = lane; t < n; t += 32) {
      const int dz = t % w - r, dy = (t / w) % w - r, dx = t / (w * w) - r;
      if (max(abs(dx), max(abs(dy), abs(dz))) != r) continue;                 // shell only
      const int x = c0x + dx, y = c0y + dy, z = c0z + dz;
      if (x < 0 || y < 0 || z < 0 || x >= G.dims[0] || y >= G.dims[1] || z >= G.dims[2]) continue;
      const int code = cell_code(x, y, z);
      const int p0 = start[code], p1 = start[code + 1];
      for (int p = p0; p < p1; ++p) {
        const float4 c = pts[p];
        const float ex = hx - c.x, ey = hy - c.y, ez = hz - c.z;
        const float d2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
        if (d2 < best) { best = d2; bidx = __float_as_int(c.w); }
      }
    }
    // everything outside the cube of half-width r cells around the (clamped) query is at least
    // r cells away from it (projection onto the grid bbox is non-expansive)
    const float wb = warp_min(best);
    const float lim = (float)r * G.cell;
    done = wb <= lim * lim * 0.9999f;
    if (r + 1 > max(G.dims[0], max(G.dims[1], G.dims[2]))) done = wb < INFINITY;   // whole grid covered
  }
  if (!done) {
    // far from the cloud: exhaustive scan by the warp (bounded, exact)
    for (int p = lane; p < P; p += 32) {
      const float4 c = pts[p];
      const float ex = hx - c.x, ey = hy - c.y, ez = hz - c.z;
      const float d2 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
      if (d2 < best) { best = d2; bidx = __float_as_int(c.w); }
    }
  }
  // warp arg-min; ties -> smallest original index
  unsigned long long key = bidx >= 0 ? (((unsigned long long)__float_as_uint(best) << 32) | (unsigned int)bidx) : ~0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned int lo = __shfl_xor_sync(0xffffffffu, (unsigned int)key, o), hi = __shfl_xor_sync(0xffffffffu, (unsigned int)(key >> 32), o);
    unsigned long long other = ((unsigned long long)hi << 32) | lo;
    key = other < key ? other : key;
  }
  if (lane == 0) ws.knn[(size_t)b * Vh + i] = key;
}

}  // namespace

// ----------------------------------------------------------------------------- host side
static inline void foho_accel_layout(FohoAccel &a, char *base, int B, int P) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += foho_align_up(bytes, 256); return p; };
  a.hand = (FohoAccelHand *)take(sizeof(FohoAccelHand) * (size_t)B);
  a.grid = (FohoAccelGrid *)take(sizeof(FohoAccelGrid) * (size_t)B);
  a.cell_start = (int *)take(sizeof(int) * (size_t)B * (FOHO_ACCEL_CELLS + 1));
  a.cell_fill = (int *)take(sizeof(int) * (size_t)B * FOHO_ACCEL_CELLS);
  a.pts = (float4 *)take(sizeof(float4) * (size_t)B * (P > 0 ? P : 1));
  a.total = off;
}

extern "C" size_t foho_guidance_accel_bytes(int32_t B, int32_t Vh, int32_t P) {
  if (B < 1 || Vh < 1 || Vh > FOHO_ACCEL_HV || P < 1) return 0;
  FohoAccel a;
  foho_accel_layout(a, nullptr, B, P);
  return a.total;
}

extern "C" int foho_guidance_prepare_statics(const foho_guidance_desc *dp, void *cuda_stream) {
  if (!dp) return FOHO_E_NULL;
  const foho_guidance_desc &d = *dp;
  if (!d.hand_rest || !d.cloud || !d.accel) return FOHO_E_NULL;
  if (d.B < 1 || d.Vh < 1 || d.Vh > FOHO_ACCEL_HV || d.P < 1) return FOHO_E_SHAPE;
  if (((uintptr_t)d.accel & 255) != 0) return FOHO_E_WORKSPACE;
  FohoAccel a;
  foho_accel_layout(a, (char *)d.accel, d.B, d.P);
  if (a.total > d.accel_bytes) return FOHO_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  k_accel_hand<<<d.B, ACC_THREADS, 0, st>>>(d.hand_rest, d.Vh, a);
  FOHO_LAUNCH_CHECK();
  k_accel_cloud_bbox<<<d.B, ACC_THREADS, 0, st>>>(d.cloud, d.P, a);
  FOHO_LAUNCH_CHECK();
  int gx = (d.P + 255) / 256;
  if (gx > 256) gx = 256;
  k_accel_cloud_count<<<dim3(gx, d.B), 256, 0, st>>>(d.cloud, d.P, a);
  FOHO_LAUNCH_CHECK();
  k_accel_cloud_scan<<<d.B, ACC_THREADS, 0, st>>>(a);
  FOHO_LAUNCH_CHECK();
  k_accel_cloud_scatter<<<dim3(gx, d.B), 256, 0, st>>>(d.cloud, d.P, a);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

int foho_launch_chamfer_accel(const foho_guidance_desc *dp, const FohoWorkspace &ws, cudaStream_t st) {
  const foho_guidance_desc &d = *dp;
  FohoAccel a;
  foho_accel_layout(a, (char *)d.accel, d.B, d.P);
  if (a.total > d.accel_bytes) return FOHO_E_WORKSPACE;
  k_chamfer_h2c<<<dim3((d.Vh + H2C_THREADS / 32 - 1) / (H2C_THREADS / 32), d.B), H2C_THREADS, 0, st>>>(d, ws, a);
  FOHO_LAUNCH_CHECK();
  k_chamfer_c2h<<<dim3((d.P + C2H_POINTS - 1) / C2H_POINTS, d.B), C2H_THREADS, 0, st>>>(d, ws, a);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
