// Thinning of surface samples by minimum distance, on the device.
//
// The reference's `icp` draws its point sets with trimesh.sample.sample_surface_even (src/foho/alignment/mesh_align.py:79,85):
// 3x oversampling, then trimesh.points.remove_close(points, radius) --
//
//   pairs  = cKDTree(points).query_pairs(radius)           every (i < j) with |p_i - p_j| <= radius
//   count  = bincount(pairs.ravel())                       in how many pairs a point appears
//   drop, for EVERY pair, the member with the larger count (the first one, i, when the counts are equal)
//
// -- one k-d tree build and one pair query per point set, four sets per image: most of the alignment stage's wall time
// once the iteration loop runs on the GPU.  Here: points are keyed by the cell of a uniform grid of pitch >= radius,
// sorted (cell, index), and every point walks the 3 x 3 runs of cells around its own twice -- once to count its
// partners, once to decide whether any pair drops it.  Float64, the same operation order as the tree (dx^2 + dy^2 +
// dz^2 without contraction, compared against radius * radius), so the mask equals the host statement bit for bit.
#include "foho_common.cuh"

namespace {

constexpr int RC_CELL_BITS = 14;                     // cells per axis (the pitch grows beyond `radius` if the box needs more)
constexpr int RC_IDX_BITS = 64 - 3 * RC_CELL_BITS;   // 22: up to 4M points
constexpr int RC_MAX_CELL = (1 << RC_CELL_BITS) - 2; // the all-ones cell is left to the padding keys

struct RcWs {
  unsigned long long *keys;   // [P2] (cx, cy, cz, index), sorted
  double *box;                // lo[3], pitch, 1/pitch
  int *count;                 // [N]
  int P2;
  size_t total;
};

inline void rc_layout(RcWs &w, char *base, int N) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += foho_align_up(bytes, 256); return p; };
  w.P2 = 2048;
  while (w.P2 < N) w.P2 <<= 1;
  w.keys = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)w.P2);
  w.box = (double *)take(sizeof(double) * 8);
  w.count = (int *)take(sizeof(int) * (size_t)N);
  w.total = off;
}

__global__ void __launch_bounds__(1024) k_rc_box(const double *__restrict__ p, int N, double radius, RcWs w) {
  __shared__ double smn[3][32], smx[3][32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < N; i += blockDim.x)
    for (int a = 0; a < 3; ++a) { const double v = p[3 * (size_t)i + a]; mn[a] = fmin(mn[a], v); mx[a] = fmax(mx[a], v); }
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  if (lane == 0) for (int a = 0; a < 3; ++a) { smn[a][wid] = mn[a]; smx[a][wid] = mx[a]; }
  __syncthreads();
  if (tid == 0) {
    double ext = 0.0;
    for (int a = 0; a < 3; ++a) {
      double lo = smn[a][0], hi = smx[a][0];
      for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { lo = fmin(lo, smn[a][k]); hi = fmax(hi, smx[a][k]); }
      w.box[a] = lo;
      ext = fmax(ext, hi - lo);
    }
    // pitch >= radius (then the 27 cells around a point hold every partner), and few enough cells for the key
    double pitch = fmax(radius * (1.0 + 1e-9), ext / (double)(RC_MAX_CELL - 1));   // the margin covers the rounding of the cell index
    if (!(pitch > 0.0)) pitch = 1.0;
    w.box[3] = pitch;
    w.box[4] = 1.0 / pitch;
  }
}

__device__ __forceinline__ void rc_cell(const RcWs &w, const double *q, int c[3]) {
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double f = floor((q[a] - w.box[a]) * w.box[4]);
    c[a] = (int)fmin(fmax(f, 0.0), (double)RC_MAX_CELL);
  }
}

__device__ __forceinline__ unsigned long long rc_key(int cx, int cy, int cz) {
  return ((((unsigned long long)cx << RC_CELL_BITS) | (unsigned long long)cy) << RC_CELL_BITS | (unsigned long long)cz) << RC_IDX_BITS;
}

__global__ void __launch_bounds__(256) k_rc_keys(const double *__restrict__ p, int N, RcWs w) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.P2; i += gridDim.x * blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < N) {
      const double q[3] = {p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]};
      int c[3];
      rc_cell(w, q, c);
      k = rc_key(c[0], c[1], c[2]) | (unsigned long long)i;
    }
    w.keys[i] = k;
  }
}

// first position whose key is >= k
__device__ __forceinline__ int rc_lower_bound(const unsigned long long *__restrict__ keys, int n, unsigned long long k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(keys + mid) < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// f(j) for every point j != i with |p_i - p_j|^2 <= r2 (the pairs of cKDTree.query_pairs that contain i)
template <typename F>
__device__ __forceinline__ void rc_for_partners(const double *__restrict__ p, int N, const RcWs &w, int i, double r2, F &&f) {
  const double q[3] = {p[3 * (size_t)i], p[3 * (size_t)i + 1], p[3 * (size_t)i + 2]};
  int c[3];
  rc_cell(w, q, c);
  const int z0 = max(c[2] - 1, 0), z1 = min(c[2] + 1, RC_MAX_CELL);
  for (int dx = -1; dx <= 1; ++dx) {
    const int cx = c[0] + dx;
    if (cx < 0 || cx > RC_MAX_CELL) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int cy = c[1] + dy;
      if (cy < 0 || cy > RC_MAX_CELL) continue;
      // cells (cx, cy, z0..z1) are one run of the sorted keys
      const unsigned long long k0 = rc_key(cx, cy, z0), k1 = rc_key(cx, cy, z1) + (1ull << RC_IDX_BITS);
      for (int s = rc_lower_bound(w.keys, N, k0); s < N; ++s) {
        const unsigned long long k = __ldg(w.keys + s);
        if (k >= k1) break;
        const int j = (int)(k & ((1ull << RC_IDX_BITS) - 1ull));
        if (j == i) continue;
        const double ex = q[0] - __ldg(p + 3 * (size_t)j), ey = q[1] - __ldg(p + 3 * (size_t)j + 1), ez = q[2] - __ldg(p + 3 * (size_t)j + 2);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
        if (d2 <= r2) f(j);
      }
    }
  }
}

__global__ void __launch_bounds__(128) k_rc_count(const double *__restrict__ p, int N, double r2, RcWs w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int n = 0;
  rc_for_partners(p, N, w, i, r2, [&](int) { ++n; });
  w.count[i] = n;
}

__global__ void __launch_bounds__(128) k_rc_mask(const double *__restrict__ p, int N, double r2, RcWs w, uint8_t *__restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int ci = w.count[i];
  bool dropped = false;
  // pair (a < b): a goes when count[a] >= count[b] (argmax takes the first of equal counts), else b
  rc_for_partners(p, N, w, i, r2, [&](int j) {
    const int cj = __ldg(w.count + j);
    if (i < j ? ci >= cj : ci > cj) dropped = true;
  });
  keep[i] = dropped ? 0 : 1;
}

}  // namespace

extern "C" size_t foho_remove_close_workspace_bytes(int32_t N) {
  if (N < 1) return 0;
  RcWs w;
  rc_layout(w, nullptr, N);
  return w.total;
}

extern "C" int foho_remove_close(const double *points, int32_t N, double radius, uint8_t *keep_mask, void *workspace,
                                 size_t workspace_bytes, void *cuda_stream) {
  if (!points || !keep_mask || !workspace) return FOHO_E_NULL;
  if (N < 1 || N >= (1 << RC_IDX_BITS)) return FOHO_E_SHAPE;
  if (!(radius >= 0.0) || !(radius < 1e300)) return FOHO_E_ARG;
  if (((uintptr_t)workspace & 255) != 0) return FOHO_E_WORKSPACE;
  RcWs w;
  rc_layout(w, (char *)workspace, N);
  if (w.total > workspace_bytes) return FOHO_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  k_rc_box<<<1, 1024, 0, st>>>(points, N, radius, w);
  int gx = (w.P2 + 255) / 256;
  if (gx > 512) gx = 512;
  k_rc_keys<<<gx, 256, 0, st>>>(points, N, w);
  FOHO_LAUNCH_CHECK();
  int rc = foho_sort_u64(w.keys, w.P2, 1, st);
  if (rc != FOHO_OK) return rc;
  const double r2 = radius * radius;
  k_rc_count<<<(N + 127) / 128, 128, 0, st>>>(points, N, r2, w);
  k_rc_mask<<<(N + 127) / 128, 128, 0, st>>>(points, N, r2, w, keep_mask);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
