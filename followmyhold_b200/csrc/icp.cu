// Trimmed similarity ICP, float64, the whole iteration loop on the device.
//
// Restates the per-iteration arithmetic of `icp` in the reference's
// src/foho/alignment/mesh_align.py:104-142 for the configuration both callers use
// (h2m.py:35-54, mano.py:24-43: on_surface=False, one "cube" = identity start):
//
//   p      = transform . source                                   (:106)
//   dist,q = Euclidean 1-NN of p in the target                    (:111-112, scipy cKDTree.query)
//   drop the n_outliers = int(outliers*count_source) largest dist (:114-120)
//   cost   = mean(inlier dist)                                    (:118)
//   next   = trimesh.registration.procrustes(p_in, q_in, reflection=False, scale=not fixed_scale) (:127)
//   transform = next @ transform; renormalise + clip scale        (:129-135)
//   keep `transform` if this iteration's (pre-update) cost is the best so far (:140-142)
//
// Two paths (no host round trip in either; state lives in the workspace):
//   k_icp_loop   (targets of >= 1024 points, the reference's sizes) ONE persistent cooperative launch for the
//                whole run: every CTA searches its source points in the box hierarchy, then -- redundantly, so
//                that no CTA has to wait for another -- radix-selects the trim threshold, sums its own inliers,
//                and after the second grid barrier reduces the per-CTA sums in a fixed order and does the 3x3
//                SVD + transform update itself.  Two grid barriers per iteration, nothing else between them.
//   k_icp_nn + k_icp_step   (small targets, n_iter <= 1, FOHO_ICP_LEGACY=1) a launch pair per iteration: tiled
//                brute-force 1-NN, then a single CTA for select / covariances / SVD / update.
#include "foho_common.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

constexpr int NN_THREADS = 256;
constexpr int NN_TILE = 256;           // target points staged per shared-memory tile
constexpr int NN_MAX_CHUNKS = 128;
constexpr int STEP_THREADS = 1024;
constexpr int LOOP_MAX_CTAS = 1024;
constexpr int LOOP_NSUM = 19;          // sum a (3), sum b (3), sum dist, sum b a^T (9), sum |a|^2, sum |b|^2, pad
constexpr int LOOP_MAX_BATCH = 64;
constexpr size_t LOOP_DESC_BYTES = 512;

struct IcpState {
  double T[16];        // current transform (row-major 4x4)
  double best_T[16];
  double best_cost;
  double cost;
  int iter;
  int pad;
};

struct IcpWorkspace {
  IcpState *state;
  double *pd2;         // [nchunks, Ns] partial min squared distance
  int *pidx;           // [nchunks, Ns]
  double *dist;        // [Ns]
  int *qi;             // [Ns]
  double *p;           // [Ns,3] transformed source of the current iteration
  // static search structure over the target (built once per run): Morton-ordered points, AABBs of every
  // 32 consecutive points and of every 32 such groups
  unsigned long long *keys;   // [P2] build scratch
  double *tpts;        // [Nt,3] target in Morton order
  int *tidx;           // [Nt]   original index of each sorted point
  double *glo, *ghi;   // [NG,3]
  double *slo, *shi;   // [NS,3]
  double *tbox;        // [6] bbox of the target
  int *seed;           // [Ns] sorted position of last iteration's neighbour (-1: none yet)
  double *ssrc;        // [Ns,3] source in Morton order (persistent loop: neighbouring warps search neighbouring boxes)
  int *sperm;          // [Ns]   original index of each sorted source point
  int P2s;             // power of two >= Ns for the source sort
  unsigned int *bar;   // persistent loop: [0] arrivals, [32] generation of CTA 0's published results (both monotonic)
  double *pub;         // persistent loop: what CTA 0 publishes: T rows 0-2 (12), trim threshold bits [16], last_eq / dup flag
  double *partials;    // [ceil(Ns/16), LOOP_NSUM] inlier sums per block of consecutive source points (persistent loop)
  void *descs;         // device copy of the problem descriptors of a batched run (first problem's workspace)
  int P2, NG, NS;
  int nchunks, chunk;
  size_t total;
};

inline void icp_ws_layout(IcpWorkspace &w, char *base, int Ns, int Nt) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += foho_align_up(bytes, 256); return p; };
  // enough chunks to fill the machine, each a multiple of the smem tile
  int blocks_s = (Ns + NN_THREADS - 1) / NN_THREADS;
  // FP64 issue is the limit (64 lanes per SM): spread the Ns x Nt distance evaluations over ~2 CTAs per SM
  int want = (2 * 148 + blocks_s - 1) / blocks_s;
  int max_chunks = (Nt + NN_TILE - 1) / NN_TILE;
  int nchunks = want < max_chunks ? want : max_chunks;
  if (nchunks < 1) nchunks = 1;
  if (nchunks > NN_MAX_CHUNKS) nchunks = NN_MAX_CHUNKS;
  int chunk = (Nt + nchunks - 1) / nchunks;
  chunk = (chunk + NN_TILE - 1) / NN_TILE * NN_TILE;
  nchunks = (Nt + chunk - 1) / chunk;
  w.nchunks = nchunks; w.chunk = chunk;
  w.state = (IcpState *)take(sizeof(IcpState));
  w.pd2 = (double *)take(sizeof(double) * (size_t)nchunks * Ns);
  w.pidx = (int *)take(sizeof(int) * (size_t)nchunks * Ns);
  w.dist = (double *)take(sizeof(double) * (size_t)Ns);
  w.qi = (int *)take(sizeof(int) * (size_t)Ns);
  w.p = (double *)take(sizeof(double) * 3 * (size_t)Ns);
  w.P2 = 2048;
  while (w.P2 < Nt) w.P2 <<= 1;
  w.NG = (Nt + 31) / 32;
  w.NS = (w.NG + 31) / 32;
  w.P2s = 2048;
  while (w.P2s < Ns) w.P2s <<= 1;
  w.keys = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)(w.P2 > w.P2s ? w.P2 : w.P2s));
  w.tpts = (double *)take(sizeof(double) * 3 * (size_t)Nt);
  w.tidx = (int *)take(sizeof(int) * (size_t)Nt);
  w.glo = (double *)take(sizeof(double) * 3 * (size_t)w.NG);
  w.ghi = (double *)take(sizeof(double) * 3 * (size_t)w.NG);
  w.slo = (double *)take(sizeof(double) * 3 * (size_t)w.NS);
  w.shi = (double *)take(sizeof(double) * 3 * (size_t)w.NS);
  w.tbox = (double *)take(sizeof(double) * 6);
  w.seed = (int *)take(sizeof(int) * (size_t)Ns);
  w.ssrc = (double *)take(sizeof(double) * 3 * (size_t)Ns);
  w.sperm = (int *)take(sizeof(int) * (size_t)Ns);
  w.bar = (unsigned int *)take(256);
  w.pub = (double *)take(256);
  w.partials = (double *)take(sizeof(double) * LOOP_NSUM * ((size_t)Ns / 16 + 1));
  w.descs = (void *)take(LOOP_DESC_BYTES * LOOP_MAX_BATCH);
  w.total = off;
}

__global__ void k_icp_init(IcpWorkspace w) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { w.state->T[i] = (i % 5 == 0) ? 1.0 : 0.0; w.state->best_T[i] = w.state->T[i]; }
    w.state->best_cost = INFINITY;
    w.state->cost = INFINITY;
    w.state->iter = 0;
    w.bar[0] = 0u; w.bar[32] = 0u;
  }
}

__global__ void __launch_bounds__(NN_THREADS) k_icp_nn(const double *__restrict__ src, int Ns,
                                                       const double *__restrict__ tgt, int Nt, IcpWorkspace w) {
  __shared__ double st[NN_TILE * 3];
  const int i = blockIdx.x * NN_THREADS + threadIdx.x;
  const int c0 = blockIdx.y * w.chunk;
  const int c1 = min(c0 + w.chunk, Nt);
  const double *T = w.state->T;
  double px = 0, py = 0, pz = 0;
  if (i < Ns) {
    const double x = src[3 * i], y = src[3 * i + 1], z = src[3 * i + 2];
    // trimesh.transform_points: dot(T[:3,:3], p) + T[:3,3]
    px = T[0] * x + T[1] * y + T[2] * z + T[3];
    py = T[4] * x + T[5] * y + T[6] * z + T[7];
    pz = T[8] * x + T[9] * y + T[10] * z + T[11];
    if (blockIdx.y == 0) { w.p[3 * i] = px; w.p[3 * i + 1] = py; w.p[3 * i + 2] = pz; }
  }
  double best = INFINITY;
  int bi = -1;
  for (int t0 = c0; t0 < c1; t0 += NN_TILE) {
    const int n = min(NN_TILE, c1 - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < 3 * n; k += NN_THREADS) st[k] = tgt[(size_t)3 * t0 + k];
    __syncthreads();
    if (i < Ns) {
#pragma unroll 4
      for (int k = 0; k < n; ++k) {
        const double dx = px - st[3 * k], dy = py - st[3 * k + 1], dz = pz - st[3 * k + 2];
        const double d2 = dx * dx + dy * dy + dz * dz;
        if (d2 < best) { best = d2; bi = t0 + k; }
      }
    }
  }
  if (i < Ns) {
    w.pd2[(size_t)blockIdx.y * Ns + i] = best;
    w.pidx[(size_t)blockIdx.y * Ns + i] = bi;
  }
}

// ---------------------------------------------------------------------------- target search structure
// The target never moves during a run (150 iterations): it is put into Morton order once and boxed in
// groups of 32 points and super-groups of 32 groups; every iteration then costs a few dozen box / point
// evaluations per source point instead of Nt (k_icp_nn above stays as the structure-free path).
__global__ void __launch_bounds__(1024) k_icp_tbox(const double *__restrict__ tgt, int Nt, IcpWorkspace w) {
  __shared__ double smn[3][32], smx[3][32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < Nt; i += blockDim.x)
    for (int a = 0; a < 3; ++a) { const double v = tgt[3 * (size_t)i + a]; mn[a] = fmin(mn[a], v); mx[a] = fmax(mx[a], v); }
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  if (lane == 0) for (int a = 0; a < 3; ++a) { smn[a][wid] = mn[a]; smx[a][wid] = mx[a]; }
  __syncthreads();
  if (tid < 3) {
    double lo = smn[tid][0], hi = smx[tid][0];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { lo = fmin(lo, smn[tid][k]); hi = fmax(hi, smx[tid][k]); }
    w.tbox[tid] = lo; w.tbox[3 + tid] = hi;
  }
}

__device__ __forceinline__ unsigned int icp_spread3(unsigned int v) {   // 10 bits -> every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) k_icp_tkeys(const double *__restrict__ tgt, int Nt, IcpWorkspace w) {
  const double ext = fmax(fmax(w.tbox[3] - w.tbox[0], w.tbox[4] - w.tbox[1]), fmax(w.tbox[5] - w.tbox[2], 1e-300));
  const double q = 1023.0 / ext;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.P2; i += gridDim.x * blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < Nt) {
      unsigned int c[3];
      for (int a = 0; a < 3; ++a) c[a] = (unsigned int)fmin(fmax((tgt[3 * (size_t)i + a] - w.tbox[a]) * q, 0.0), 1023.0);
      k = ((unsigned long long)((icp_spread3(c[0]) << 2) | (icp_spread3(c[1]) << 1) | icp_spread3(c[2])) << 32) | (unsigned int)i;
    }
    w.keys[i] = k;
  }
}

__global__ void __launch_bounds__(256) k_icp_tgather(const double *__restrict__ tgt, int Nt, int Ns, IcpWorkspace w) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < Nt; s += gridDim.x * blockDim.x) {
    const unsigned int i = (unsigned int)w.keys[s];
    w.tpts[3 * (size_t)s] = tgt[3 * (size_t)i]; w.tpts[3 * (size_t)s + 1] = tgt[3 * (size_t)i + 1];
    w.tpts[3 * (size_t)s + 2] = tgt[3 * (size_t)i + 2];
    w.tidx[s] = (int)i;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Ns; i += gridDim.x * blockDim.x) w.seed[i] = -1;
}

// the source in Morton order too (keys on the target's box; the start transform is the identity and every update a
// small similarity, so neighbours stay neighbours): the warps of a CTA then walk the same few boxes of the target
__global__ void __launch_bounds__(256) k_icp_skeys(const double *__restrict__ src, int Ns, IcpWorkspace w) {
  const double ext = fmax(fmax(w.tbox[3] - w.tbox[0], w.tbox[4] - w.tbox[1]), fmax(w.tbox[5] - w.tbox[2], 1e-300));
  const double q = 1023.0 / ext;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w.P2s; i += gridDim.x * blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < Ns) {
      unsigned int c[3];
      for (int a = 0; a < 3; ++a) c[a] = (unsigned int)fmin(fmax((src[3 * (size_t)i + a] - w.tbox[a]) * q, 0.0), 1023.0);
      k = ((unsigned long long)((icp_spread3(c[0]) << 2) | (icp_spread3(c[1]) << 1) | icp_spread3(c[2])) << 32) | (unsigned int)i;
    }
    w.keys[i] = k;
  }
}

__global__ void __launch_bounds__(256) k_icp_sgather(const double *__restrict__ src, int Ns, IcpWorkspace w) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < Ns; s += gridDim.x * blockDim.x) {
    const unsigned int i = (unsigned int)w.keys[s];
    w.ssrc[3 * (size_t)s] = src[3 * (size_t)i]; w.ssrc[3 * (size_t)s + 1] = src[3 * (size_t)i + 1];
    w.ssrc[3 * (size_t)s + 2] = src[3 * (size_t)i + 2];
    w.sperm[s] = (int)i;
  }
}

// level 0: one warp per group of 32 sorted points; level 1: one warp per 32 groups
__global__ void __launch_bounds__(256) k_icp_tboxes(int Nt, int level, IcpWorkspace w) {
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_out = level == 0 ? w.NG : w.NS, n_in = level == 0 ? Nt : w.NG;
  if (g >= n_out) return;
  const int i = g * 32 + lane, ii = i < n_in ? i : g * 32;
  double lo[3], hi[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = level == 0 ? w.tpts[3 * (size_t)ii + a] : w.glo[3 * (size_t)ii + a];
    hi[a] = level == 0 ? lo[a] : w.ghi[3 * (size_t)ii + a];
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  }
  if (lane == 0)
    for (int a = 0; a < 3; ++a) {
      (level == 0 ? w.glo : w.slo)[3 * (size_t)g + a] = lo[a];
      (level == 0 ? w.ghi : w.shi)[3 * (size_t)g + a] = hi[a];
    }
}

__device__ __forceinline__ double icp_box_d2(const double *lo, const double *hi, double x, double y, double z) {
  const double dx = fmax(fmax(lo[0] - x, x - hi[0]), 0.0);
  const double dy = fmax(fmax(lo[1] - y, y - hi[1]), 0.0);
  const double dz = fmax(fmax(lo[2] - z, z - hi[2]), 0.0);
  return dx * dx + dy * dy + dz * dz;
}

struct IcpBest { double d2; int idx, pos; };
// fold the 32 lane candidates into the warp-uniform best: smallest d2, ties -> smallest ORIGINAL index
// (what the brute-force scan in index order returns)
__device__ __forceinline__ void icp_fold(IcpBest &b, double d2, int idx, int pos) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d2, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o), op = __shfl_xor_sync(0xffffffffu, pos, o);
    if (od < d2 || (od == d2 && oi < idx)) { d2 = od; idx = oi; pos = op; }
  }
  if (d2 < b.d2 || (d2 == b.d2 && idx < b.idx)) { b.d2 = d2; b.idx = idx; b.pos = pos; }
}

// one warp per source point: transform, then exact 1-NN over the box hierarchy, warm-started from the
// group that held last iteration's neighbour.  Writes dist (Euclidean, like cKDTree.query) and the index.
__global__ void __launch_bounds__(256) k_icp_nn_tree(const double *__restrict__ src, int Ns, int Nt, IcpWorkspace w,
                                                     int *__restrict__ nn_out) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= Ns) return;
  const double *T = w.state->T;
  const double x = src[3 * (size_t)i], y = src[3 * (size_t)i + 1], z = src[3 * (size_t)i + 2];
  const double px = T[0] * x + T[1] * y + T[2] * z + T[3];
  const double py = T[4] * x + T[5] * y + T[6] * z + T[7];
  const double pz = T[8] * x + T[9] * y + T[10] * z + T[11];
  if (lane == 0) { w.p[3 * (size_t)i] = px; w.p[3 * (size_t)i + 1] = py; w.p[3 * (size_t)i + 2] = pz; }
  IcpBest b; b.d2 = INFINITY; b.idx = 0x7fffffff; b.pos = -1;
  auto scan_group = [&](int g) {
    const int s = g * 32 + lane;
    double d2 = INFINITY; int idx = 0x7fffffff;
    if (s < Nt) {
      const double dx = px - w.tpts[3 * (size_t)s], dy = py - w.tpts[3 * (size_t)s + 1], dz = pz - w.tpts[3 * (size_t)s + 2];
      d2 = dx * dx + dy * dy + dz * dz;
      idx = w.tidx[s];
    }
    icp_fold(b, d2, idx, s);
  };
  auto scan_super = [&](int sg, int skip) {
    const int g = sg * 32 + lane;
    double lb = INFINITY;
    if (g < w.NG && g != skip) lb = icp_box_d2(w.glo + 3 * (size_t)g, w.ghi + 3 * (size_t)g, px, py, pz);
    unsigned mask = __ballot_sync(0xffffffffu, lb <= b.d2);
    while (mask) {
      scan_group(sg * 32 + __ffs(mask) - 1);
      mask &= mask - 1;
      mask &= __ballot_sync(0xffffffffu, lb <= b.d2);
    }
  };
  int s0 = -1, g0 = -1;
  const int seed = w.seed[i];
  if (seed >= 0 && seed < Nt) {
    g0 = seed >> 5;
    scan_group(g0);
  } else {
    // greedy descent: nearest super-group, its nearest group
    double bl = INFINITY; int bs = 0;
    for (int s = lane; s < w.NS; s += 32) {
      const double v = icp_box_d2(w.slo + 3 * (size_t)s, w.shi + 3 * (size_t)s, px, py, pz);
      if (v < bl) { bl = v; bs = s; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bl, o);
      const int os = __shfl_xor_sync(0xffffffffu, bs, o);
      if (ov < bl || (ov == bl && os < bs)) { bl = ov; bs = os; }
    }
    s0 = bs;
    const int g = s0 * 32 + lane;
    double lb = g < w.NG ? icp_box_d2(w.glo + 3 * (size_t)g, w.ghi + 3 * (size_t)g, px, py, pz) : INFINITY;
    int bg = g;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, lb, o);
      const int og = __shfl_xor_sync(0xffffffffu, bg, o);
      if (ov < lb || (ov == lb && og < bg)) { lb = ov; bg = og; }
    }
    g0 = bg;
    scan_group(g0);
    scan_super(s0, g0);
  }
  for (int sb = 0; sb < w.NS; sb += 32) {
    const int s = sb + lane;
    double lb = INFINITY;
    if (s < w.NS && s != s0) lb = icp_box_d2(w.slo + 3 * (size_t)s, w.shi + 3 * (size_t)s, px, py, pz);
    unsigned mask = __ballot_sync(0xffffffffu, lb <= b.d2);
    while (mask) {
      const int sg = sb + __ffs(mask) - 1;
      scan_super(sg, sg == (g0 >> 5) ? g0 : -1);
      mask &= mask - 1;
      mask &= __ballot_sync(0xffffffffu, lb <= b.d2);
    }
  }
  if (lane == 0) {
    w.dist[i] = sqrt(b.d2);
    w.qi[i] = b.idx;
    w.seed[i] = b.pos;
    if (nn_out) nn_out[i] = b.idx;
  }
}

// ---- block-wide helpers for k_icp_step (1024 threads)
// N sums at once: one pair of barriers for all of them; fixed order => deterministic, same value everywhere
template <int N>
__device__ void block_sum_dn(double (&v)[N], double *sm /* [N*32] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < N; ++k) sm[k * 32 + wid] = v[k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double r = 0;
    for (int w = 0; w < nw; ++w) r += sm[k * 32 + w];
    v[k] = r;
  }
}

__global__ void __launch_bounds__(STEP_THREADS) k_icp_step(const double *__restrict__ tgt, int Ns, int n_outliers,
                                                           int fixed_scale, double min_scale, double max_scale,
                                                           IcpWorkspace w, double *cost_history, int *nn_out, int reduce_partials) {
  __shared__ double smd[11 * 32];
  __shared__ unsigned int hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_remaining;
  __shared__ int s_take_eq;
  const int tid = threadIdx.x;

  // 1. reduce the per-chunk partial minima; dist = sqrt(d2) (cKDTree returns Euclidean distance)
  for (int i = tid; reduce_partials && i < Ns; i += STEP_THREADS) {
    double best = w.pd2[i];
    int bi = w.pidx[i];
    for (int c = 1; c < w.nchunks; ++c) {
      double v = w.pd2[(size_t)c * Ns + i];
      if (v < best) { best = v; bi = w.pidx[(size_t)c * Ns + i]; }
    }
    w.dist[i] = sqrt(best);
    w.qi[i] = bi;
    if (nn_out) nn_out[i] = bi;
  }
  __syncthreads();

  // 2. trim threshold = the n_in-th smallest distance (n_in = Ns - n_outliers), MSB radix select
  //    over the IEEE bits (non-negative doubles order like unsigned integers).
  const int n_in = Ns - (n_outliers > 0 ? n_outliers : 0);
  unsigned long long thr = ~0ull;
  int take_eq = 0;                       // how many elements equal to thr are inliers (lowest indices first)
  // the distances stay in registers over the eight radix passes (up to 8 per thread: Ns <= 8192); larger
  // sets re-read them from memory
  constexpr int REG_PER = 8;
  const bool in_regs = Ns <= REG_PER * STEP_THREADS;
  unsigned long long rbits[REG_PER];
  if (in_regs) {
#pragma unroll
    for (int k = 0; k < REG_PER; ++k) {
      const int i = tid + k * STEP_THREADS;
      rbits[k] = i < Ns ? (unsigned long long)__double_as_longlong(w.dist[i]) : ~0ull;
    }
  }
  if (n_outliers > 0) {
    if (tid == 0) { s_prefix = 0ull; s_remaining = n_in; }
    __syncthreads();
    for (int shift = 56; shift >= 0; shift -= 8) {
      if (tid < 256) hist[tid] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      const unsigned long long mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
      // the distances share their leading bytes, so most lanes of a warp hit the same bucket: one
      // shared-memory atomic per distinct digit per warp instead of one per element
#pragma unroll
      for (int k = 0; k < REG_PER; ++k) {
        const int base = k * STEP_THREADS;
        if (base >= Ns) break;                                    // uniform
        const int i = base + tid;
        unsigned int key = 256u;                                  // 256 = not a candidate
        if (i < Ns) {
          const unsigned long long bits = in_regs ? rbits[k] : (unsigned long long)__double_as_longlong(w.dist[i]);
          if ((bits & mask) == prefix) key = (unsigned int)((bits >> shift) & 255ull);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key < 256u && (tid & 31) == __ffs(peers) - 1) atomicAdd(&hist[key], (unsigned int)__popc(peers));
      }
      for (int base = REG_PER * STEP_THREADS; base < Ns; base += STEP_THREADS) {     // Ns > 8192 only
        const int i = base + tid;
        unsigned int key = 256u;
        if (i < Ns) {
          const unsigned long long bits = (unsigned long long)__double_as_longlong(w.dist[i]);
          if ((bits & mask) == prefix) key = (unsigned int)((bits >> shift) & 255ull);
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key < 256u && (tid & 31) == __ffs(peers) - 1) atomicAdd(&hist[key], (unsigned int)__popc(peers));
      }
      __syncthreads();
      if (tid < 32) {
        // warp 0: lane l owns buckets 8l..8l+7; find the bucket holding the s_remaining-th element
        const int rem0 = s_remaining;
        unsigned int c[8], tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = hist[tid * 8 + k]; tot += c[k]; }
        unsigned int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
        const unsigned int excl = incl - tot;
        const bool mine = (int)excl < rem0 && (int)incl >= rem0;          // exactly one lane (rem0 <= total)
        const unsigned who = __ballot_sync(0xffffffffu, mine);
        if (who == 0u) {
          if (tid == 0) { s_prefix = prefix | (255ull << shift); s_remaining = 0; }
        } else if (mine) {
          int rem = rem0 - (int)excl, bkt = 0;
          for (; bkt < 8; ++bkt) {
            if ((int)c[bkt] >= rem) break;
            rem -= (int)c[bkt];
          }
          if (bkt > 7) bkt = 7;
          s_prefix = prefix | ((unsigned long long)(tid * 8 + bkt) << shift);
          s_remaining = rem;
        }
      }
      __syncthreads();
    }
    thr = s_prefix;
    if (tid == 0) s_take_eq = s_remaining;   // of the elements == thr, this many are inliers
    __syncthreads();
    take_eq = s_take_eq;
  }
  // rank elements equal to thr by index (stable) so exactly n_in inliers are kept
  // (ties at the threshold are measure-zero for real data; the rule only has to be deterministic)
  auto is_inlier = [&](int i, int &eq_seen) -> bool {
    if (n_outliers <= 0) return true;
    unsigned long long bits = (unsigned long long)__double_as_longlong(w.dist[i]);
    if (bits < thr) return true;
    if (bits > thr) return false;
    (void)eq_seen;
    return true;   // provisional; corrected below when duplicates exist
  };
  // count elements == thr; if more than take_eq exist, only thread 0 resolves (rare path)
  __shared__ int s_eq_total;
  if (tid == 0) s_eq_total = 0;
  __syncthreads();
  if (n_outliers > 0) {
    int local = 0;
    for (int i = tid; i < Ns; i += STEP_THREADS)
      if ((unsigned long long)__double_as_longlong(w.dist[i]) == thr) ++local;
    if (local) atomicAdd(&s_eq_total, local);
  }
  __syncthreads();
  const bool dup_ties = n_outliers > 0 && s_eq_total > take_eq;
  __shared__ int s_last_eq_index;          // inliers among ties: index <= s_last_eq_index
  if (tid == 0) {
    s_last_eq_index = 0x7fffffff;
    if (dup_ties) {
      int seen = 0;
      for (int i = 0; i < Ns; ++i)
        if ((unsigned long long)__double_as_longlong(w.dist[i]) == thr) {
          if (++seen == take_eq) { s_last_eq_index = i; break; }
        }
      if (take_eq == 0) s_last_eq_index = -1;
    }
  }
  __syncthreads();
  const int last_eq = s_last_eq_index;

  // 3. inlier statistics (two passes: means, then centred second moments)
  double sa[3] = {0, 0, 0}, sb[3] = {0, 0, 0}, sd = 0;
  int dummy = 0;
  for (int i = tid; i < Ns; i += STEP_THREADS) {
    bool in = is_inlier(i, dummy);
    if (in && dup_ties && (unsigned long long)__double_as_longlong(w.dist[i]) == thr && i > last_eq) in = false;
    if (!in) continue;
    const int q = w.qi[i];
    sa[0] += w.p[3 * i]; sa[1] += w.p[3 * i + 1]; sa[2] += w.p[3 * i + 2];
    sb[0] += tgt[3 * (size_t)q]; sb[1] += tgt[3 * (size_t)q + 1]; sb[2] += tgt[3 * (size_t)q + 2];
    sd += w.dist[i];
  }
  const double n = (double)n_in;
  double am[3], bm[3];
  double r7[7] = {sa[0], sa[1], sa[2], sb[0], sb[1], sb[2], sd};
  block_sum_dn<7>(r7, smd);
  for (int a = 0; a < 3; ++a) { am[a] = r7[a] / n; bm[a] = r7[3 + a] / n; }
  const double cost = r7[6] / n;
  double h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, va = 0, vb = 0;
  for (int i = tid; i < Ns; i += STEP_THREADS) {
    bool in = is_inlier(i, dummy);
    if (in && dup_ties && (unsigned long long)__double_as_longlong(w.dist[i]) == thr && i > last_eq) in = false;
    if (!in) continue;
    const int q = w.qi[i];
    const double a0 = w.p[3 * i] - am[0], a1 = w.p[3 * i + 1] - am[1], a2 = w.p[3 * i + 2] - am[2];
    const double b0 = tgt[3 * (size_t)q] - bm[0], b1 = tgt[3 * (size_t)q + 1] - bm[1], b2 = tgt[3 * (size_t)q + 2] - bm[2];
    va += a0 * a0 + a1 * a1 + a2 * a2;
    vb += b0 * b0 + b1 * b1 + b2 * b2;
    h[0] += b0 * a0; h[1] += b0 * a1; h[2] += b0 * a2;
    h[3] += b1 * a0; h[4] += b1 * a1; h[5] += b1 * a2;
    h[6] += b2 * a0; h[7] += b2 * a1; h[8] += b2 * a2;
  }
  double H[3][3];
  double r11[11] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], va, vb};
  block_sum_dn<11>(r11, smd);
  for (int k = 0; k < 9; ++k) H[k / 3][k % 3] = r11[k];
  va = r11[9];
  vb = r11[10];

  // 4. similarity fit + transform update (thread 0)
  if (tid == 0) {
    IcpState *S = w.state;
    double ascale = 1.0, bscale = 1.0;
    if (!fixed_scale) { ascale = sqrt(va / n); bscale = sqrt(vb / n); }
    // trimesh divides both centred sets by their scale before the SVD; a positive scalar on H
    // does not change its singular vectors, so H is used as accumulated.
    double R[3][3];
    kabsch_rotation(H, R);
    const double sc = bscale / ascale;
    double M[16];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) M[4 * i + j] = sc * R[i][j];
      M[4 * i + 3] = bm[i] - sc * (R[i][0] * am[0] + R[i][1] * am[1] + R[i][2] * am[2]);
    }
    M[12] = M[13] = M[14] = 0.0; M[15] = 1.0;
    double Tn[16];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double acc = 0;
        for (int k = 0; k < 4; ++k) acc += M[4 * i + k] * S->T[4 * k + j];
        Tn[4 * i + j] = acc;
      }
    if (!fixed_scale) {
      double s0 = sqrt(Tn[0] * Tn[0] + Tn[4] * Tn[4] + Tn[8] * Tn[8]);     // norm of the first column (:132)
      double s1 = fmin(fmax(s0, min_scale), max_scale);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Tn[4 * i + j] = Tn[4 * i + j] / s0 * s1;
    }
    for (int k = 0; k < 16; ++k) S->T[k] = Tn[k];
    S->cost = cost;
    if (cost < S->best_cost) {          // updated transform paired with the pre-update cost (:140-142)
      S->best_cost = cost;
      for (int k = 0; k < 16; ++k) S->best_T[k] = Tn[k];
    }
    if (cost_history) cost_history[S->iter] = cost;
    S->iter += 1;
  }
}

// ---------------------------------------------------------------------------- persistent loop
struct IcpProblem {
  const double *src, *tgt;
  int Ns, Nt, n_iter, n_outliers, fixed_scale, pad;
  double min_scale, max_scale;
  double *T_out, *cost_out, *cost_history;
  int *nn_out;
  IcpWorkspace w;
};
static_assert(sizeof(IcpProblem) <= LOOP_DESC_BYTES, "descriptor slot too small");

// exact minimum over the warp of a non-negative double (its IEEE bits order like an unsigned integer): two
// redux.sync instead of a five-step shuffle tree on three values
__device__ __forceinline__ double icp_warp_min_nonneg(double v) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned int hi = (unsigned int)(b >> 32), lo = (unsigned int)b;
  const unsigned int mhi = __reduce_min_sync(0xffffffffu, hi);
  const unsigned int mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
  return __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
}

// Exact 1-NN of one transformed source point in the box hierarchy, one warp.  Same result as k_icp_nn_tree (smallest
// squared distance, ties -> smallest ORIGINAL index); every lane keeps the best of the points IT looked at and only
// the pruning bound is warp-uniform, so a group costs two redux.sync instead of a three-value shuffle tree.
#ifdef FOHO_ICP_PROFILE
__device__ long long g_nn_prof[8];      // groups scanned, supers scanned, points, cycles: seed, supers, total
#define NN_COUNT(k, v) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_nn_prof[k] += (long long)(v); } while (0)
#else
#define NN_COUNT(k, v) do { } while (0)
#endif
__device__ __forceinline__ void icp_nn_point(const IcpWorkspace &w, int Nt, int lane, double px, double py, double pz,
                                             int seed, double &d2_out, int &idx_out, int &pos_out) {
#ifdef FOHO_ICP_PROFILE
  const long long tp0 = clock64();
#endif
  double bd = INFINITY, bound = INFINITY;
  int bidx = 0x7fffffff, bpos = -1;
  auto take = [&](double d2, int idx, int s) {
    if (d2 < bd || (d2 == bd && idx < bidx)) { bd = d2; bidx = idx; bpos = s; }
  };
  auto scan_group = [&](int g) {
    const int s = g * 32 + lane;
    if (s < Nt) {
      const double dx = px - __ldg(w.tpts + 3 * (size_t)s), dy = py - __ldg(w.tpts + 3 * (size_t)s + 1),
                   dz = pz - __ldg(w.tpts + 3 * (size_t)s + 2);
      take(dx * dx + dy * dy + dz * dz, __ldg(w.tidx + s), s);
    }
    bound = icp_warp_min_nonneg(bd);
    NN_COUNT(0, 1);
  };
  // two groups per step: their loads are in flight together and the bound is folded once
  auto scan_group2 = [&](int ga, int gb) {
    const int sa = ga * 32 + lane, sb = gb * 32 + lane;
    const bool va = sa < Nt, vb = sb < Nt;
    double ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
    int ia = 0, ib = 0;
    if (va) { ax = __ldg(w.tpts + 3 * (size_t)sa); ay = __ldg(w.tpts + 3 * (size_t)sa + 1); az = __ldg(w.tpts + 3 * (size_t)sa + 2); ia = __ldg(w.tidx + sa); }
    if (vb) { bx = __ldg(w.tpts + 3 * (size_t)sb); by = __ldg(w.tpts + 3 * (size_t)sb + 1); bz = __ldg(w.tpts + 3 * (size_t)sb + 2); ib = __ldg(w.tidx + sb); }
    if (va) { const double dx = px - ax, dy = py - ay, dz = pz - az; take(dx * dx + dy * dy + dz * dz, ia, sa); }
    if (vb) { const double dx = px - bx, dy = py - by, dz = pz - bz; take(dx * dx + dy * dy + dz * dz, ib, sb); }
    bound = icp_warp_min_nonneg(bd);
    NN_COUNT(0, 2);
  };
  auto scan_super = [&](int sg, int skip) {
    NN_COUNT(1, 1);
    const int g = sg * 32 + lane;
    double lb = INFINITY;
    if (g < w.NG && g != skip) lb = icp_box_d2(w.glo + 3 * (size_t)g, w.ghi + 3 * (size_t)g, px, py, pz);
    unsigned mask = __ballot_sync(0xffffffffu, lb <= bound);
    while (mask) {
      const int ga = sg * 32 + __ffs(mask) - 1;
      mask &= mask - 1;
      if (mask) {
        const int gb = sg * 32 + __ffs(mask) - 1;
        mask &= mask - 1;
        scan_group2(ga, gb);
      } else {
        scan_group(ga);
      }
      mask &= __ballot_sync(0xffffffffu, lb <= bound);
    }
  };
  int s0 = -1, g0 = -1;
  if (seed >= 0 && seed < Nt) {
    g0 = seed >> 5;
    scan_group(g0);
#ifdef FOHO_ICP_PROFILE
    NN_COUNT(3, clock64() - tp0);
#endif
  } else {
    double bl = INFINITY; int bs = 0;
    for (int s = lane; s < w.NS; s += 32) {
      const double v = icp_box_d2(w.slo + 3 * (size_t)s, w.shi + 3 * (size_t)s, px, py, pz);
      if (v < bl) { bl = v; bs = s; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bl, o);
      const int os = __shfl_xor_sync(0xffffffffu, bs, o);
      if (ov < bl || (ov == bl && os < bs)) { bl = ov; bs = os; }
    }
    s0 = bs;
    const int g = s0 * 32 + lane;
    double lb = g < w.NG ? icp_box_d2(w.glo + 3 * (size_t)g, w.ghi + 3 * (size_t)g, px, py, pz) : INFINITY;
    int bg = g;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, lb, o);
      const int og = __shfl_xor_sync(0xffffffffu, bg, o);
      if (ov < lb || (ov == lb && og < bg)) { lb = ov; bg = og; }
    }
    g0 = bg;
    scan_group(g0);
    scan_super(s0, g0);
  }
  for (int sb = 0; sb < w.NS; sb += 32) {
    const int s = sb + lane;
    double lb = INFINITY;
    if (s < w.NS && s != s0) lb = icp_box_d2(w.slo + 3 * (size_t)s, w.shi + 3 * (size_t)s, px, py, pz);
    unsigned mask = __ballot_sync(0xffffffffu, lb <= bound);
    while (mask) {
      const int sg = sb + __ffs(mask) - 1;
      scan_super(sg, sg == (g0 >> 5) ? g0 : -1);
      mask &= mask - 1;
      mask &= __ballot_sync(0xffffffffu, lb <= bound);
    }
  }
#ifdef FOHO_ICP_PROFILE
  NN_COUNT(2, 1); NN_COUNT(5, clock64() - tp0);
#endif
  const unsigned int cand = bd == bound ? (unsigned int)bidx : 0xffffffffu;
  unsigned int midx = __reduce_min_sync(0xffffffffu, cand);
  const unsigned int who = __ballot_sync(0xffffffffu, bd == bound && (unsigned int)bidx == midx);
  pos_out = __shfl_sync(0xffffffffu, bpos, who ? __ffs(who) - 1 : 0);
  if (midx >= (unsigned int)Nt) { midx = 0u; pos_out = -1; }     // NaN coordinates: nothing compared less than infinity
  d2_out = bound;
  idx_out = (int)midx;
}

// similarity fit (trimesh.registration.procrustes, reflection=False) from the summed moments about `shift`, transform
// update with the scale renormalised and clipped (:127-135), best-by-pre-update-cost (:140-142).  One thread.
__device__ __noinline__ double icp_fit_update(const double *tot, const double *shift, double n, int fixed_scale,
                                              double min_scale, double max_scale, double *T, double *bestT, double *best_cost) {
  // (divisions are the long poles of this one-thread step: one reciprocal of n, one square root for the scale)
  const double rn = 1.0 / n;
  double am[3], bm[3], amc[3], bmc[3];
  for (int a = 0; a < 3; ++a) { amc[a] = tot[a] * rn; bmc[a] = tot[3 + a] * rn; am[a] = amc[a] + shift[a]; bm[a] = bmc[a] + shift[a]; }
  const double cost = tot[6] * rn;
  double H[3][3];
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) H[j][k] = tot[7 + 3 * j + k] - n * bmc[j] * amc[k];
  const double va = tot[16] - n * (amc[0] * amc[0] + amc[1] * amc[1] + amc[2] * amc[2]);
  const double vb = tot[17] - n * (bmc[0] * bmc[0] + bmc[1] * bmc[1] + bmc[2] * bmc[2]);
  // scale = sqrt(vb / n) / sqrt(va / n), the ratio of the RMS radii of the two centred sets
  const double sc = fixed_scale ? 1.0 : sqrt(vb / va);
  // trimesh divides both centred sets by their scale before the SVD; a positive scalar on H does not change its
  // singular vectors, so H is used as accumulated.
  double Rm[3][3];
  kabsch_rotation(H, Rm);
  double M[16];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) M[4 * i + j] = sc * Rm[i][j];
    M[4 * i + 3] = bm[i] - sc * (Rm[i][0] * am[0] + Rm[i][1] * am[1] + Rm[i][2] * am[2]);
  }
  M[12] = M[13] = M[14] = 0.0; M[15] = 1.0;
  double Tn[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s2 = 0;
      for (int k = 0; k < 4; ++k) s2 += M[4 * i + k] * T[4 * k + j];
      Tn[4 * i + j] = s2;
    }
  if (!fixed_scale) {
    const double s0 = sqrt(Tn[0] * Tn[0] + Tn[4] * Tn[4] + Tn[8] * Tn[8]);     // norm of the first column (:132)
    const double f = fmin(fmax(s0, min_scale), max_scale) / s0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Tn[4 * i + j] *= f;
  }
  for (int k = 0; k < 16; ++k) T[k] = Tn[k];
  if (cost < *best_cost) {
    *best_cost = cost;
    for (int k = 0; k < 16; ++k) bestT[k] = Tn[k];
  }
  return cost;
}

__device__ __forceinline__ unsigned long long icp_low_mask(int bits) { return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull); }

template <int LOOP_THREADS>
__global__ void __launch_bounds__(LOOP_THREADS, 1) k_icp_loop(const IcpProblem single, const IcpProblem *__restrict__ many) {
  __shared__ IcpProblem P;
  __shared__ double sT[16], sBestT[16], sTot[LOOP_NSUM], sShift[3];
  __shared__ double sBestCost;
  constexpr int SEL_BITS = 11, SEL_PER = (1 << SEL_BITS) / LOOP_THREADS;     // radix-select digit; bins per thread
  constexpr int STASH_R = 12;                           // rounds of step A whose pairs stay in shared memory for step C
  __shared__ double sPair[STASH_R * (LOOP_THREADS / 32)][7];
  __shared__ unsigned int hist[1 << SEL_BITS];
  __shared__ unsigned int s_andor[4], s_wtot[32];
  __shared__ unsigned long long s_prefix, s_cand[32];
  __shared__ int s_remaining, s_cnt, s_last_eq, s_ncand;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  {
    const unsigned int *from = many ? (const unsigned int *)(many + blockIdx.y) : (const unsigned int *)&single;
    unsigned int *to = (unsigned int *)&P;
    for (int k = tid; k < (int)(sizeof(IcpProblem) / 4); k += LOOP_THREADS) to[k] = from[k];
  }
  if (tid < 16) { sT[tid] = (tid % 5 == 0) ? 1.0 : 0.0; sBestT[tid] = sT[tid]; }
  if (tid == 0) sBestCost = INFINITY;
  __syncthreads();
  const IcpWorkspace &w = P.w;
  const int Ns = P.Ns, Nt = P.Nt, n_outliers = P.n_outliers;
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  constexpr int W = LOOP_THREADS / 32;
  const int R = (Ns + G * W - 1) / (G * W);           // source points per warp
  if (tid < 3) sShift[tid] = 0.5 * (w.tbox[tid] + w.tbox[3 + tid]);
  __syncthreads();
  unsigned int bar_target = 0u, gen_target = 0u;
  const int n_in = Ns - (n_outliers > 0 ? n_outliers : 0);
  const double n = (double)n_in;
  constexpr int REG_PER = 8;
  const bool in_regs = Ns <= REG_PER * LOOP_THREADS;
  // Synchronisation between the CTAs of this problem: everybody ARRIVES (w.bar counts), CTA 0 waits for all
  // arrivals, does the serial step alone -- one reader of the distances / block sums instead of gridDim.x readers of
  // the same cache lines -- writes the result to w.pub and bumps the generation counter the others spin on.
  unsigned int *const gen = w.bar + 32;                             // its own 128-byte line
  auto arrive = [&]() {
    __syncthreads();
    if (tid == 0 && G > 1) { __threadfence(); atomicAdd(w.bar, 1u); }
  };
  auto leader_wait = [&]() {                                        // CTA 0
    if (tid == 0 && G > 1) {
      bar_target += (unsigned int)G;
      unsigned int v;
      do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(w.bar) : "memory"); } while (v < bar_target);
    }
    __syncthreads();
  };
  auto publish = [&]() {                                            // CTA 0, after its threads wrote w.pub
    __syncthreads();
    if (tid == 0 && G > 1) { __threadfence(); atomicAdd(gen, 1u); }
  };
  auto follower_wait = [&]() {                                      // the others
    if (tid == 0) {
      gen_target += 1u;
      unsigned int v;
      do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(gen) : "memory"); } while (v < gen_target);
    }
    __syncthreads();
  };

#ifdef FOHO_ICP_PROFILE
  long long prof[6] = {0, 0, 0, 0, 0, 0}, prof2[3] = {0, 0, 0}, tprev = clock64();
  int npass = 0;
  long long ts0 = 0;
#define ICP_PROF(k) do { const long long tn = clock64(); prof[k] += tn - tprev; tprev = tn; } while (0)
#else
#define ICP_PROF(k) do { } while (0)
#endif
  for (int it = 0; it < P.n_iter; ++it) {
    // ---- A. transform + 1-NN of this CTA's source points (point i always goes to the same warp: its seed stays local)
    for (int r = 0; r < R; ++r) {
      const int sl = (r * G + c) * W + wid;                         // slot in Morton order
      if (sl >= Ns) break;                                          // warp-uniform
      const double x = __ldg(w.ssrc + 3 * (size_t)sl), y = __ldg(w.ssrc + 3 * (size_t)sl + 1), z = __ldg(w.ssrc + 3 * (size_t)sl + 2);
      const double px = sT[0] * x + sT[1] * y + sT[2] * z + sT[3];
      const double py = sT[4] * x + sT[5] * y + sT[6] * z + sT[7];
      const double pz = sT[8] * x + sT[9] * y + sT[10] * z + sT[11];
      const int seed = it == 0 ? -1 : w.seed[sl];
      const int i = __ldg(w.sperm + sl);                            // original index: distances and ties are kept by it
      double d2; int idx, pos;
      icp_nn_point(w, Nt, lane, px, py, pz, seed, d2, idx, pos);
      if (lane == 0) {
        const double dist = sqrt(d2);
        w.dist[i] = dist;
        w.seed[sl] = pos;
        if (P.nn_out && it == P.n_iter - 1) P.nn_out[i] = idx;
        if (r < STASH_R) {
          // this CTA sums these points itself in step C: the pair stays in shared memory (about the fixed point)
          double *sp = sPair[r * W + wid];
          const double *bq = pos >= 0 ? w.tpts + 3 * (size_t)pos : P.tgt + 3 * (size_t)idx;
          sp[0] = px - sShift[0]; sp[1] = py - sShift[1]; sp[2] = pz - sShift[2];
          sp[3] = __ldg(bq) - sShift[0]; sp[4] = __ldg(bq + 1) - sShift[1]; sp[5] = __ldg(bq + 2) - sShift[2];
          sp[6] = dist;
        } else {
          w.p[3 * (size_t)sl] = px; w.p[3 * (size_t)sl + 1] = py; w.p[3 * (size_t)sl + 2] = pz;
          w.qi[sl] = idx;
        }
      }
    }
    ICP_PROF(0);
    if (n_outliers > 0) arrive();
    else __syncthreads();                                           // step C reads what other warps left in sPair

    // ---- B. (CTA 0) trim threshold = the n_in-th smallest distance: MSB radix select over the IEEE bits (non-negative
    //         doubles order like unsigned integers), 11 bits per pass, starting below the bits all distances share;
    //         once at most 32 candidates are left one warp ranks them
    unsigned long long thr = ~0ull;
    int last_eq = 0x7fffffff;
    bool dup_ties = false;
    if (n_outliers > 0) {
      if (c == 0) {
        leader_wait();
        ICP_PROF(1);
        int take_eq = 0, eq_total = 0;
        unsigned long long rbits[REG_PER];
        {
          if (tid < 4) s_andor[tid] = tid < 2 ? 0xffffffffu : 0u;
          __syncthreads();
          unsigned long long a = ~0ull, o = 0ull;
          if (in_regs) {
    #pragma unroll
            for (int k = 0; k < REG_PER; ++k) {
              const int i = tid + k * LOOP_THREADS;
              rbits[k] = i < Ns ? (unsigned long long)__double_as_longlong(__ldcg(w.dist + i)) : ~0ull;
              if (i < Ns) { a &= rbits[k]; o |= rbits[k]; }
            }
          } else {
            for (int i = tid; i < Ns; i += LOOP_THREADS) {
              const unsigned long long b = (unsigned long long)__double_as_longlong(__ldcg(w.dist + i));
              a &= b; o |= b;
            }
          }
          {
            const unsigned int ah = __reduce_and_sync(0xffffffffu, (unsigned int)(a >> 32)), al = __reduce_and_sync(0xffffffffu, (unsigned int)a);
            const unsigned int oh = __reduce_or_sync(0xffffffffu, (unsigned int)(o >> 32)), ol = __reduce_or_sync(0xffffffffu, (unsigned int)o);
            if (lane == 0) { atomicAnd(&s_andor[0], ah); atomicAnd(&s_andor[1], al); atomicOr(&s_andor[2], oh); atomicOr(&s_andor[3], ol); }
          }
          __syncthreads();
    #ifdef FOHO_ICP_PROFILE
      ts0 = clock64();
#endif
      const unsigned long long all_and = ((unsigned long long)s_andor[0] << 32) | s_andor[1];
          const unsigned long long all_or = ((unsigned long long)s_andor[2] << 32) | s_andor[3];
          const unsigned long long diff = all_and ^ all_or;
          int hi = diff ? 64 - __clzll((long long)diff) : 0;            // bits [hi-1 .. 0] are not settled yet
          unsigned long long prefix = all_or & ~icp_low_mask(hi);
          int remaining = n_in, cnt = Ns;
          // every candidate's bits, one after the other, to f(bits)
          auto for_candidates = [&](unsigned long long keep, unsigned long long want, auto &&f) {
            if (in_regs) {
    #pragma unroll
              for (int k = 0; k < REG_PER; ++k)
                if (tid + k * LOOP_THREADS < Ns && (rbits[k] & keep) == want) f(rbits[k]);
            } else {
              for (int i = tid; i < Ns; i += LOOP_THREADS) {
                const unsigned long long b = (unsigned long long)__double_as_longlong(__ldcg(w.dist + i));
                if ((b & keep) == want) f(b);
              }
            }
          };
          while (true) {
            if (hi == 0) { thr = prefix; take_eq = remaining; eq_total = cnt; break; }    // the candidates are all equal
            if (cnt <= 32) {
              // gather the candidates (any order) and let warp 0 find the remaining-th smallest among them
              if (tid == 0) s_ncand = 0;
              __syncthreads();
              for_candidates(~icp_low_mask(hi), prefix, [&](unsigned long long b) { s_cand[atomicAdd(&s_ncand, 1) & 31] = b; });
              __syncthreads();
              if (wid == 0) {
                const int nc = s_ncand;
                const unsigned long long v = lane < nc ? s_cand[lane] : ~0ull;
                int less = 0, eq = 0;
                for (int j = 0; j < nc; ++j) {
                  const unsigned long long u = s_cand[j];
                  less += u < v; eq += u == v;
                }
                if (lane < nc && less < remaining && remaining <= less + eq) { s_prefix = v; s_remaining = remaining - less; s_cnt = eq; }
              }
              __syncthreads();
              thr = s_prefix; take_eq = s_remaining; eq_total = s_cnt;
              break;
            }
            const int wd = hi < SEL_BITS ? hi : SEL_BITS, shift = hi - wd;
#ifdef FOHO_ICP_PROFILE
        ++npass;
#endif
    #pragma unroll
            for (int k = 0; k < SEL_PER; ++k) hist[tid * SEL_PER + k] = 0u;
            __syncthreads();
            for_candidates(~icp_low_mask(hi), prefix, [&](unsigned long long b) {
              atomicAdd(&hist[(unsigned int)((b >> shift) & ((1ull << wd) - 1ull))], 1u);
            });
            __syncthreads();
            // block-wide exclusive scan of the bins (thread t owns bins t*SEL_PER ..), then the owner of the bin that holds
            // the remaining-th candidate publishes it
            unsigned int cc[SEL_PER], tot = 0;
    #pragma unroll
            for (int k = 0; k < SEL_PER; ++k) { cc[k] = hist[tid * SEL_PER + k]; tot += cc[k]; }
            unsigned int incl = tot;
    #pragma unroll
            for (int o2 = 1; o2 < 32; o2 <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, incl, o2); if (lane >= o2) incl += t; }
            if (lane == 31) s_wtot[wid] = incl;
            __syncthreads();
            if (wid == 0) {
              const unsigned int mine = lane < W ? s_wtot[lane] : 0u;
              unsigned int inc2 = mine;
    #pragma unroll
              for (int o2 = 1; o2 < 32; o2 <<= 1) { const unsigned int t = __shfl_up_sync(0xffffffffu, inc2, o2); if (lane >= o2) inc2 += t; }
              s_wtot[lane] = inc2 - mine;                               // exclusive prefix of each warp (W <= 32)
            }
            __syncthreads();
            {
              const unsigned int excl = s_wtot[wid] + incl - tot;
              if ((int)excl < remaining && (int)(excl + tot) >= remaining) {      // exactly one thread (remaining <= cnt)
                int rem = remaining - (int)excl, bkt = 0;
    #pragma unroll
                for (int k = 0; k < SEL_PER - 1; ++k)
                  if (bkt == k && (int)cc[k] < rem) { rem -= (int)cc[k]; bkt = k + 1; }
                s_prefix = prefix | ((unsigned long long)(tid * SEL_PER + bkt) << shift);
                s_remaining = rem;
                s_cnt = (int)cc[bkt];
              }
            }
            __syncthreads();
            prefix = s_prefix; remaining = s_remaining; cnt = s_cnt; hi = shift;
            __syncthreads();                                             // s_prefix and s_wtot are rewritten next round
          }
        }
        // duplicates at the threshold (measure zero for real data; the rule only has to be deterministic): the
        // take_eq lowest indices among them are inliers
#ifdef FOHO_ICP_PROFILE
        prof2[0] += ts0 - tprev; prof2[1] += clock64() - ts0; prof2[2] += npass; npass = 0;
#endif
        dup_ties = eq_total > take_eq;
        if (dup_ties) {
          if (tid == 0) {
            int seen = 0, last = -1;
            for (int i = 0; i < Ns && seen < take_eq; ++i)
              if ((unsigned long long)__double_as_longlong(__ldcg(w.dist + i)) == thr) { ++seen; last = i; }
            s_last_eq = last;
          }
          __syncthreads();
        }

        last_eq = dup_ties ? s_last_eq : 0x7fffffff;
        if (tid == 0) {
          ((unsigned long long *)w.pub)[16] = thr;
          ((int *)w.pub)[34] = last_eq;
          ((int *)w.pub)[35] = dup_ties ? 1 : 0;
        }
        publish();
      } else {
        follower_wait();
        ICP_PROF(1);
        thr = __ldcg((const unsigned long long *)w.pub + 16);
        last_eq = __ldcg((const int *)w.pub + 34);
        dup_ties = __ldcg((const int *)w.pub + 35) != 0;
      }
    }
    ICP_PROF(2);

    // ---- C. sums over this CTA's inliers, taken about a fixed point near the data (the centre of the target's box):
    //         one pass gives the means and the centred second moments.  Sums are formed per block of W consecutive
    //         slots (a shuffle tree) and the blocks added in index order in step D, so the result does not depend on
    //         how many CTAs share the problem.  Warp k forms sum k straight from the pairs step A left in shared memory.
    constexpr int CH = W < 32 ? W : 32;
    const int own = R * W;
    const int own_stash = own < STASH_R * W ? own : STASH_R * W;
    for (int k = wid; k < LOOP_NSUM - 1; k += W) {
      for (int t0 = 0; t0 < own_stash; t0 += 32) {
        const int t = t0 + lane;
        const int blk = (t / W) * G + c;
        const int sl = blk * W + (t % W);
        double v = 0.0;
        if (t < own_stash && sl < Ns) {
          const double *sp = sPair[t];
          const unsigned long long bits = (unsigned long long)__double_as_longlong(sp[6]);
          bool in = n_outliers <= 0 || bits <= thr;
          if (dup_ties && bits == thr) in = __ldg(w.sperm + sl) <= last_eq;
          if (in) {
            if (k < 7) v = sp[k];
            else if (k < 16) v = sp[3 + (k - 7) / 3] * sp[(k - 7) % 3];
            else if (k == 16) v = sp[0] * sp[0] + sp[1] * sp[1] + sp[2] * sp[2];
            else v = sp[3] * sp[3] + sp[4] * sp[4] + sp[5] * sp[5];
          }
        }
#pragma unroll
        for (int o = CH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((lane & (CH - 1)) == 0 && t < own_stash && blk * W < Ns) w.partials[(size_t)blk * LOOP_NSUM + k] = v;
      }
    }
    for (int t0 = STASH_R * W; t0 < own; t0 += LOOP_THREADS) {        // more rounds than the stash holds: from memory
      const int t = t0 + tid;
      const int blk = (t / W) * G + c;
      const int sl = blk * W + (t % W);
      double acc[LOOP_NSUM - 1];
#pragma unroll
      for (int k = 0; k < LOOP_NSUM - 1; ++k) acc[k] = 0.0;
      if (t < own && sl < Ns) {
        const int i = __ldg(w.sperm + sl);
        const int q = w.qi[sl];
        const double dist = __ldcg(w.dist + i);
        const double a0 = w.p[3 * (size_t)sl] - sShift[0], a1 = w.p[3 * (size_t)sl + 1] - sShift[1], a2 = w.p[3 * (size_t)sl + 2] - sShift[2];
        const double b0 = __ldg(P.tgt + 3 * (size_t)q) - sShift[0], b1 = __ldg(P.tgt + 3 * (size_t)q + 1) - sShift[1],
                     b2 = __ldg(P.tgt + 3 * (size_t)q + 2) - sShift[2];
        const unsigned long long bits = (unsigned long long)__double_as_longlong(dist);
        const bool in = n_outliers <= 0 || bits < thr || (bits == thr && (!dup_ties || i <= last_eq));
        if (in) {
          acc[0] = a0; acc[1] = a1; acc[2] = a2; acc[3] = b0; acc[4] = b1; acc[5] = b2; acc[6] = dist;
          acc[7] = b0 * a0; acc[8] = b0 * a1; acc[9] = b0 * a2;
          acc[10] = b1 * a0; acc[11] = b1 * a1; acc[12] = b1 * a2;
          acc[13] = b2 * a0; acc[14] = b2 * a1; acc[15] = b2 * a2;
          acc[16] = a0 * a0 + a1 * a1 + a2 * a2;
          acc[17] = b0 * b0 + b1 * b1 + b2 * b2;
        }
      }
      if (t0 + (wid << 5) < own) {                                   // warp-uniform: this warp holds a block
#pragma unroll
        for (int k = 0; k < LOOP_NSUM - 1; ++k) {
          double v = acc[k];
#pragma unroll
          for (int o = CH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if ((lane & (CH - 1)) == 0 && t < own && blk * W < Ns) w.partials[(size_t)blk * LOOP_NSUM + k] = v;
        }
      }
    }
    ICP_PROF(3);
    arrive();

    // ---- D. (CTA 0) block sums in index order, the similarity fit and the transform update; the others pick T up
    if (c == 0) {
      leader_wait();
      ICP_PROF(4);
      for (int k = wid; k < LOOP_NSUM - 1; k += W) {
        const int nblk = (Ns + W - 1) / W;
        double v = 0.0;
        for (int j = lane; j < nblk; j += 32) v += __ldcg(w.partials + (size_t)j * LOOP_NSUM + k);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sTot[k] = v;
      }
      __syncthreads();
      if (tid == 0) {
        const double cost = icp_fit_update(sTot, sShift, n, P.fixed_scale, P.min_scale, P.max_scale, sT, sBestT, &sBestCost);
        if (P.cost_history) P.cost_history[it] = cost;
        if (G > 1)
          for (int k = 0; k < 12; ++k) w.pub[k] = sT[k];
      }
      publish();
    } else {
      follower_wait();
      ICP_PROF(4);
      if (tid < 12) sT[tid] = __ldcg(w.pub + tid);
      __syncthreads();
    }
    ICP_PROF(5);
  }
#ifdef FOHO_ICP_PROFILE
  if (tid == 0 && (c == 0 || c == G - 1) && blockIdx.y == 0)
    if (c == 0) printf("icp nn per point: groups %.2f supers %.2f cycles seed-scan %.0f total %.0f (points %lld)\n", (double)g_nn_prof[0] / g_nn_prof[2],
                       (double)g_nn_prof[1] / g_nn_prof[2], (double)g_nn_prof[3] / g_nn_prof[2], (double)g_nn_prof[5] / g_nn_prof[2], g_nn_prof[2]);
    printf("icp profile cta %d/%d: cycles per iteration  nn %lld  bar1 %lld  select %lld  sums %lld  bar2 %lld  fit %lld | select: load %lld rest %lld passes x100 %lld\n", c, G,
           prof[0] / P.n_iter, prof[1] / P.n_iter, prof[2] / P.n_iter, prof[3] / P.n_iter, prof[4] / P.n_iter, prof[5] / P.n_iter,
           prof2[0] / P.n_iter, prof2[1] / P.n_iter, 100 * prof2[2] / P.n_iter);
#endif
  if (c == 0) {
    if (tid < 16) P.T_out[tid] = sBestT[tid];
    if (tid == 0 && P.cost_out) P.cost_out[0] = sBestCost;
  }
}

__global__ void k_icp_finish(IcpWorkspace w, double *T_out, double *cost_out) {
  if (threadIdx.x < 16) T_out[threadIdx.x] = w.state->best_T[threadIdx.x];
  if (threadIdx.x == 0 && cost_out) cost_out[0] = w.state->best_cost;
}

}  // namespace

extern "C" size_t foho_icp_workspace_bytes(int32_t Ns, int32_t Nt) {
  if (Ns < 1 || Nt < 1) return 0;
  IcpWorkspace w;
  icp_ws_layout(w, nullptr, Ns, Nt);
  return w.total;
}

namespace {

struct IcpDevice { int sms; int loop_ok; int threads; };

// SM count and whether the persistent kernel fits (one CTA of LOOP_THREADS per SM), per device, looked up once
int icp_device_info(IcpDevice &out) {
  static IcpDevice cache[64];
  static bool have[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= 64) return FOHO_E_ARG;
  if (!have[dev]) {
    int sms = 0, coop = 0, occ = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
    if ((e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev)) != cudaSuccess) return (int)e;
    const char *thr = getenv("FOHO_ICP_LOOP_THREADS");
    const int threads = thr && atoi(thr) == 512 ? 512 : 1024;
    if (threads == 512) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_icp_loop<512>, 512, 0);
    else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_icp_loop<1024>, 1024, 0);
    if (e != cudaSuccess) return (int)e;
    const char *legacy = getenv("FOHO_ICP_LEGACY");
    cache[dev].sms = sms;
    cache[dev].threads = threads;
    cache[dev].loop_ok = coop && occ >= 1 && !(legacy && legacy[0] == '1');
    have[dev] = true;
  }
  out = cache[dev];
  return FOHO_OK;
}

int icp_validate(const IcpProblem &p, void *workspace, size_t workspace_bytes) {
  if (!p.src || !p.tgt || !p.T_out || !workspace) return FOHO_E_NULL;
  if (p.Ns < 1 || p.Nt < 1 || p.n_iter < 0) return FOHO_E_SHAPE;
  if (p.n_outliers < 0 || p.n_outliers >= p.Ns) return FOHO_E_ARG;
  if (!(p.min_scale > 0.0) || !(p.max_scale >= p.min_scale)) return FOHO_E_ARG;
  if (((uintptr_t)workspace & 255) != 0) return FOHO_E_WORKSPACE;
  if (p.w.total > workspace_bytes) return FOHO_E_WORKSPACE;
  return FOHO_OK;
}

inline bool icp_uses_tree(const IcpProblem &p) { return p.Nt >= 1024 && p.n_iter > 1; }

// state reset + (for the tree paths) the search structure over the target
int icp_prepare(const IcpProblem &p, bool sort_source, cudaStream_t st) {
  const IcpWorkspace &w = p.w;
  k_icp_init<<<1, 32, 0, st>>>(w);
  FOHO_LAUNCH_CHECK();
  if (icp_uses_tree(p)) {
    k_icp_tbox<<<1, 1024, 0, st>>>(p.tgt, p.Nt, w);
    int gx = (w.P2 + 255) / 256;
    if (gx > 512) gx = 512;
    k_icp_tkeys<<<gx, 256, 0, st>>>(p.tgt, p.Nt, w);
    FOHO_LAUNCH_CHECK();
    int rc = foho_sort_u64(w.keys, w.P2, 1, st);
    if (rc != FOHO_OK) return rc;
    k_icp_tgather<<<gx, 256, 0, st>>>(p.tgt, p.Nt, p.Ns, w);
    k_icp_tboxes<<<(w.NG + 7) / 8, 256, 0, st>>>(p.Nt, 0, w);
    k_icp_tboxes<<<(w.NS + 7) / 8, 256, 0, st>>>(p.Nt, 1, w);
    FOHO_LAUNCH_CHECK();
    if (sort_source) {
      int gs = (w.P2s + 255) / 256;
      if (gs > 512) gs = 512;
      k_icp_skeys<<<gs, 256, 0, st>>>(p.src, p.Ns, w);
      FOHO_LAUNCH_CHECK();
      rc = foho_sort_u64(w.keys, w.P2s, 1, st);
      if (rc != FOHO_OK) return rc;
      k_icp_sgather<<<gs, 256, 0, st>>>(p.src, p.Ns, w);
      FOHO_LAUNCH_CHECK();
    }
  }
  return FOHO_OK;
}

// a launch pair per iteration (small targets, single iterations, FOHO_ICP_LEGACY=1)
int icp_run_launch_pairs(const IcpProblem &p, cudaStream_t st) {
  const IcpWorkspace &w = p.w;
  const bool tree = icp_uses_tree(p);
  const dim3 nn_grid((p.Ns + NN_THREADS - 1) / NN_THREADS, w.nchunks);
  for (int it = 0; it < p.n_iter; ++it) {
    int *nn_last = it == p.n_iter - 1 ? p.nn_out : nullptr;
    if (tree) k_icp_nn_tree<<<(p.Ns + 7) / 8, 256, 0, st>>>(p.src, p.Ns, p.Nt, w, nn_last);
    else k_icp_nn<<<nn_grid, NN_THREADS, 0, st>>>(p.src, p.Ns, p.tgt, p.Nt, w);
    k_icp_step<<<1, STEP_THREADS, 0, st>>>(p.tgt, p.Ns, p.n_outliers, p.fixed_scale, p.min_scale, p.max_scale, w,
                                           p.cost_history, tree ? nullptr : nn_last, tree ? 0 : 1);
  }
  FOHO_LAUNCH_CHECK();
  k_icp_finish<<<1, 32, 0, st>>>(w, p.T_out, p.cost_out);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

// one cooperative launch: `n` problems side by side, ctas CTAs each (many == nullptr: the single problem by value)
int icp_run_persistent(const IcpProblem &single, const IcpProblem *many_dev, int n, int ctas, int threads, cudaStream_t st) {
  void *args[2] = {(void *)&single, (void *)&many_dev};
  const void *fn = threads == 512 ? (const void *)k_icp_loop<512> : (const void *)k_icp_loop<1024>;
  cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(ctas, n), dim3(threads), args, 0, st);
  return e == cudaSuccess ? FOHO_OK : (int)e;
}

inline int icp_ctas_for(int Ns, int avail, int threads) {
  const int want = (Ns + threads / 32 - 1) / (threads / 32);       // no more CTAs than there are warps' worth of points
  int g = avail < want ? avail : want;
  if (g < 1) g = 1;
  if (g > LOOP_MAX_CTAS) g = LOOP_MAX_CTAS;
  return g;
}

IcpProblem icp_problem_of(const double *source, int32_t Ns, const double *target, int32_t Nt, int32_t n_iter,
                          int32_t n_outliers, int32_t fixed_scale, double min_scale, double max_scale,
                          double *transform_out, double *cost_out, double *cost_history, int32_t *nn_index_last,
                          void *workspace) {
  IcpProblem p;
  memset(&p, 0, sizeof(p));
  p.src = source; p.tgt = target; p.Ns = Ns; p.Nt = Nt; p.n_iter = n_iter; p.n_outliers = n_outliers;
  p.fixed_scale = fixed_scale; p.min_scale = min_scale; p.max_scale = max_scale;
  p.T_out = transform_out; p.cost_out = cost_out; p.cost_history = cost_history; p.nn_out = nn_index_last;
  if (Ns >= 1 && Nt >= 1) icp_ws_layout(p.w, (char *)workspace, Ns, Nt);
  return p;
}

}  // namespace

extern "C" int foho_icp_run(const double *source, int32_t Ns, const double *target, int32_t Nt, int32_t n_iter,
                            int32_t n_outliers, int32_t fixed_scale, double min_scale, double max_scale,
                            double *transform_out, double *cost_out, double *cost_history, int32_t *nn_index_last,
                            void *workspace, size_t workspace_bytes, void *cuda_stream) {
  if (Ns < 1 || Nt < 1) return (!source || !target || !transform_out || !workspace) ? FOHO_E_NULL : FOHO_E_SHAPE;
  const IcpProblem p = icp_problem_of(source, Ns, target, Nt, n_iter, n_outliers, fixed_scale, min_scale, max_scale,
                                      transform_out, cost_out, cost_history, nn_index_last, workspace);
  int rc = icp_validate(p, workspace, workspace_bytes);
  if (rc != FOHO_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  IcpDevice dev;
  if ((rc = icp_device_info(dev)) != FOHO_OK) return rc;
  if ((rc = icp_prepare(p, dev.loop_ok, st)) != FOHO_OK) return rc;
  if (icp_uses_tree(p) && dev.loop_ok) return icp_run_persistent(p, nullptr, 1, icp_ctas_for(Ns, dev.sms, dev.threads), dev.threads, st);
  return icp_run_launch_pairs(p, st);
}

extern "C" int foho_icp_run_batch(const foho_icp_problem *problems, int32_t n_problems, void *cuda_stream) {
  if (!problems) return FOHO_E_NULL;
  if (n_problems < 0) return FOHO_E_SHAPE;
  if (n_problems == 0) return FOHO_OK;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  IcpDevice dev;
  int rc = icp_device_info(dev);
  if (rc != FOHO_OK) return rc;
  std::vector<IcpProblem> loop;            // problems that go into the shared persistent launch
  for (int k = 0; k < n_problems; ++k) {
    const foho_icp_problem &q = problems[k];
    if (q.Ns < 1 || q.Nt < 1) return (!q.source || !q.target || !q.transform_out || !q.workspace) ? FOHO_E_NULL : FOHO_E_SHAPE;
    const IcpProblem p = icp_problem_of(q.source, q.Ns, q.target, q.Nt, q.n_iter, q.n_outliers, q.fixed_scale, q.min_scale,
                                        q.max_scale, q.transform_out, q.cost_out, q.cost_history, q.nn_index_last, q.workspace);
    if ((rc = icp_validate(p, q.workspace, q.workspace_bytes)) != FOHO_OK) return rc;
    if ((rc = icp_prepare(p, dev.loop_ok, st)) != FOHO_OK) return rc;
    if (icp_uses_tree(p) && dev.loop_ok) loop.push_back(p);
    else if ((rc = icp_run_launch_pairs(p, st)) != FOHO_OK) return rc;
  }
  const int per_launch = dev.sms < LOOP_MAX_BATCH ? dev.sms : LOOP_MAX_BATCH;
  for (size_t k0 = 0; k0 < loop.size(); k0 += per_launch) {
    const int nb = (int)std::min(loop.size() - k0, (size_t)per_launch);
    if (nb == 1) {
      if ((rc = icp_run_persistent(loop[k0], nullptr, 1, icp_ctas_for(loop[k0].Ns, dev.sms, dev.threads), dev.threads, st)) != FOHO_OK) return rc;
      continue;
    }
    int ns_max = 0;
    for (int k = 0; k < nb; ++k) ns_max = std::max(ns_max, loop[k0 + k].Ns);
    // the descriptors travel through the first problem's workspace (pageable source: staged before the call returns)
    void *dst = loop[k0].w.descs;
    cudaError_t e = cudaMemcpyAsync(dst, &loop[k0], (size_t)nb * sizeof(IcpProblem), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    if ((rc = icp_run_persistent(loop[k0], (const IcpProblem *)dst, nb, icp_ctas_for(ns_max, dev.sms / nb, dev.threads), dev.threads, st)) != FOHO_OK) return rc;
  }
  return FOHO_OK;
}
