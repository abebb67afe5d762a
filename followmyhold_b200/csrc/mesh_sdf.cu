// Exact mesh -> signed distance on a rectilinear lattice, and the REF penetration count.
//
// Replaces `mesh2sdf` / `get_sdf_of_meshes` of the reference's
// third_party/utilz/kaolin_sdf_ops.py:88-109,131-160 (kaolin point_to_mesh_distance +
// check_sign on the (res+1)^3 union-bbox grid) and `honerf_intersection_loss`
// (third_party_patches/hy3dgen/shapegen/pipelines.py:231-239).
//
//   k_m2s_raster : +z ray-parity voxelisation (same bit-exact rule as the guidance kernels,
//                  evaluated at the lattice's actual float32 coordinates)
//   k_m2s_dist   : one thread per lattice point, face bounding spheres staged through shared
//                  memory, exact closest-point test only for faces that can beat the current best
#include "foho_common.cuh"

namespace {

struct M2sWorkspace {
  uint32_t *parity;   // [nx*ny*W]
  float4 *sph;        // [F]
  int W;
  size_t total;
};

inline void m2s_layout(M2sWorkspace &w, char *base, int F, int nx, int ny, int nz) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += foho_align_up(bytes, 256); return p; };
  w.W = (nz + 31) / 32;
  w.parity = (uint32_t *)take(sizeof(uint32_t) * (size_t)nx * ny * w.W);
  w.sph = (float4 *)take(sizeof(float4) * (size_t)F);
  w.total = off;
}

__device__ __forceinline__ int lower_index(const float *c, int n, float v) {
  // first index i with c[i] >= v (c ascending)
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (c[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(128) k_m2s_raster(const float *__restrict__ verts, const int *__restrict__ faces, int F,
                                                    const float *__restrict__ xs, const float *__restrict__ ys,
                                                    const float *__restrict__ zs, int nx, int ny, int nz,
                                                    M2sWorkspace w) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
  const foho_f3 a = f3(verts[3 * ia], verts[3 * ia + 1], verts[3 * ia + 2]);
  const foho_f3 b = f3(verts[3 * ib], verts[3 * ib + 1], verts[3 * ib + 2]);
  const foho_f3 c = f3(verts[3 * ic], verts[3 * ic + 1], verts[3 * ic + 2]);
  const float xmin = fminf(a.x, fminf(b.x, c.x)), xmax = fmaxf(a.x, fmaxf(b.x, c.x));
  const float ymin = fminf(a.y, fminf(b.y, c.y)), ymax = fmaxf(a.y, fmaxf(b.y, c.y));
  const int i0 = lower_index(xs, nx, xmin), j0 = lower_index(ys, ny, ymin);
  // sphere for the distance pass
  {
    foho_f3 m = (1.f / 3.f) * (a + b + c);
    foho_f3 da = a - m, db = b - m, dc = c - m;
    float r2 = fmaxf(dot3(da, da), fmaxf(dot3(db, db), dot3(dc, dc)));
    w.sph[f] = make_float4(m.x, m.y, m.z, sqrtf(r2) * 1.00001f + 1e-12f);
  }
  for (int i = i0; i < nx && xs[i] <= xmax; ++i)
    for (int j = j0; j < ny && ys[j] <= ymax; ++j) {
      float zc;
      if (!column_hits_triangle(ia, ib, ic, a, b, c, xs[i], ys[j], &zc)) continue;
      if (!(zc == zc)) continue;
      const int nzb = lower_index(zs, nz, zc);         // number of k with zs[k] < zc
      if (nzb <= 0) continue;
      uint32_t *col = w.parity + ((size_t)i * ny + j) * w.W;
      const int full = nzb >> 5, rem = nzb & 31;
      for (int q = 0; q < full; ++q) atomicXor(col + q, 0xFFFFFFFFu);
      if (rem) atomicXor(col + full, (1u << rem) - 1u);
    }
}

constexpr int M2S_THREADS = 256;
constexpr int M2S_TILE = 1024;

__global__ void __launch_bounds__(M2S_THREADS) k_m2s_dist(const float *__restrict__ verts, const int *__restrict__ faces,
                                                          int F, const float *__restrict__ xs, const float *__restrict__ ys,
                                                          const float *__restrict__ zs, int nx, int ny, int nz,
                                                          float *__restrict__ out, M2sWorkspace w) {
  __shared__ float4 ssph[M2S_TILE];
  const long long n = (long long)nx * ny * nz;
  const long long idx = (long long)blockIdx.x * M2S_THREADS + threadIdx.x;
  const bool live = idx < n;
  int iz = 0, iy = 0, ix = 0;
  foho_f3 p = f3(0.f, 0.f, 0.f);
  if (live) {
    iz = (int)(idx % nz);
    long long r = idx / nz;
    iy = (int)(r % ny);
    ix = (int)(r / ny);
    p = f3(xs[ix], ys[iy], zs[iz]);
  }
  float best2 = INFINITY;
  for (int t0 = 0; t0 < F; t0 += M2S_TILE) {
    const int m = min(M2S_TILE, F - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += M2S_THREADS) ssph[k] = w.sph[t0 + k];
    __syncthreads();
    if (!live) continue;
    for (int k = 0; k < m; ++k) {
      const float4 s = ssph[k];
      const foho_f3 q = f3(s.x, s.y, s.z) - p;
      const float dc = sqrtf(dot3(q, q));
      const float lb = dc - s.w;
      if (lb > 0.f && lb * lb > best2) continue;
      const int f = t0 + k;
      const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
      float wa, wb, wc;
      const float d2 = closest_point_triangle(f3(0.f, 0.f, 0.f), f3(verts[3 * ia], verts[3 * ia + 1], verts[3 * ia + 2]) - p,
                                              f3(verts[3 * ib], verts[3 * ib + 1], verts[3 * ib + 2]) - p,
                                              f3(verts[3 * ic], verts[3 * ic + 1], verts[3 * ic + 2]) - p, wa, wb, wc);
      best2 = fminf(best2, d2);
    }
  }
  if (live) {
    const uint32_t bits = w.parity[((size_t)ix * ny + iy) * w.W + (iz >> 5)];
    const bool inside = (bits >> (iz & 31)) & 1u;
    const float dist = sqrtf(best2);
    out[idx] = inside ? -dist : dist;
  }
}

__global__ void __launch_bounds__(256) k_count(const float *__restrict__ sh, const float *__restrict__ so, long long n,
                                               unsigned long long *out) {
  unsigned int c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += (so[i] < 0.f && sh[i] < 0.f) ? 1u : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

}  // namespace

extern "C" size_t foho_mesh2sdf_workspace_bytes(int32_t V, int32_t F, int32_t nx, int32_t ny, int32_t nz) {
  if (V < 1 || F < 1 || nx < 1 || ny < 1 || nz < 1) return 0;
  M2sWorkspace w;
  m2s_layout(w, nullptr, F, nx, ny, nz);
  return w.total;
}

extern "C" int foho_mesh2sdf_lattice(const float *verts, int32_t V, const int32_t *faces, int32_t F, const float *xs,
                                     const float *ys, const float *zs, int32_t nx, int32_t ny, int32_t nz, float *sdf_out,
                                     void *workspace, size_t workspace_bytes, void *cuda_stream) {
  if (!verts || !faces || !xs || !ys || !zs || !sdf_out || !workspace) return FOHO_E_NULL;
  if (V < 1 || F < 1 || nx < 1 || ny < 1 || nz < 1) return FOHO_E_SHAPE;
  if (((uintptr_t)workspace & 255) != 0) return FOHO_E_WORKSPACE;
  M2sWorkspace w;
  m2s_layout(w, (char *)workspace, F, nx, ny, nz);
  if (w.total > workspace_bytes) return FOHO_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  FOHO_CUDA_TRY(cudaMemsetAsync(w.parity, 0, sizeof(uint32_t) * (size_t)nx * ny * w.W, st));
  k_m2s_raster<<<(F + 127) / 128, 128, 0, st>>>(verts, faces, F, xs, ys, zs, nx, ny, nz, w);
  FOHO_LAUNCH_CHECK();
  const long long n = (long long)nx * ny * nz;
  k_m2s_dist<<<(unsigned)((n + M2S_THREADS - 1) / M2S_THREADS), M2S_THREADS, 0, st>>>(verts, faces, F, xs, ys, zs, nx, ny,
                                                                                     nz, sdf_out, w);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

extern "C" int foho_intersection_count(const float *sdf_hand, const float *sdf_obj, int64_t n, long long *count_out,
                                       void *cuda_stream) {
  if (!sdf_hand || !sdf_obj || !count_out) return FOHO_E_NULL;
  if (n < 1) return FOHO_E_SHAPE;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  FOHO_CUDA_TRY(cudaMemsetAsync(count_out, 0, sizeof(long long), st));
  long long blocks = (n + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  k_count<<<(int)blocks, 256, 0, st>>>(sdf_hand, sdf_obj, n, reinterpret_cast<unsigned long long *>(count_out));
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
