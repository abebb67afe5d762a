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
// Kernels per iteration (no host round trip; state lives in the workspace):
//   k_icp_nn     tiled brute-force 1-NN, target chunk in shared memory, partial minima per chunk
//   k_icp_step   single CTA: reduce the partial minima, radix-select the trim threshold,
//                centred covariances, 3x3 SVD (Jacobi), transform/scale update, best tracking
#include "foho_common.cuh"

namespace {

constexpr int NN_THREADS = 256;
constexpr int NN_TILE = 256;           // target points staged per shared-memory tile
constexpr int NN_MAX_CHUNKS = 128;
constexpr int STEP_THREADS = 1024;

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
  w.keys = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)w.P2);
  w.tpts = (double *)take(sizeof(double) * 3 * (size_t)Nt);
  w.tidx = (int *)take(sizeof(int) * (size_t)Nt);
  w.glo = (double *)take(sizeof(double) * 3 * (size_t)w.NG);
  w.ghi = (double *)take(sizeof(double) * 3 * (size_t)w.NG);
  w.slo = (double *)take(sizeof(double) * 3 * (size_t)w.NS);
  w.shi = (double *)take(sizeof(double) * 3 * (size_t)w.NS);
  w.tbox = (double *)take(sizeof(double) * 6);
  w.seed = (int *)take(sizeof(int) * (size_t)Ns);
  w.total = off;
}

__global__ void k_icp_init(IcpWorkspace w) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) { w.state->T[i] = (i % 5 == 0) ? 1.0 : 0.0; w.state->best_T[i] = w.state->T[i]; }
    w.state->best_cost = INFINITY;
    w.state->cost = INFINITY;
    w.state->iter = 0;
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
__device__ double block_sum_d(double v, double *sm) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[wid] = v;
  __syncthreads();
  double r = 0;
  const int nw = blockDim.x >> 5;
  for (int k = 0; k < nw; ++k) r += sm[k];      // fixed order: deterministic, same value in every thread
  return r;
}

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

extern "C" int foho_icp_run(const double *source, int32_t Ns, const double *target, int32_t Nt, int32_t n_iter,
                            int32_t n_outliers, int32_t fixed_scale, double min_scale, double max_scale,
                            double *transform_out, double *cost_out, double *cost_history, int32_t *nn_index_last,
                            void *workspace, size_t workspace_bytes, void *cuda_stream) {
  if (!source || !target || !transform_out || !workspace) return FOHO_E_NULL;
  if (Ns < 1 || Nt < 1 || n_iter < 0) return FOHO_E_SHAPE;
  if (n_outliers < 0 || n_outliers >= Ns) return FOHO_E_ARG;
  if (!(min_scale > 0.0) || !(max_scale >= min_scale)) return FOHO_E_ARG;
  if (((uintptr_t)workspace & 255) != 0) return FOHO_E_WORKSPACE;
  IcpWorkspace w;
  icp_ws_layout(w, (char *)workspace, Ns, Nt);
  if (w.total > workspace_bytes) return FOHO_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  k_icp_init<<<1, 32, 0, st>>>(w);
  FOHO_LAUNCH_CHECK();
  // small targets: the tiled brute-force scan; otherwise the box hierarchy over the (static) target
  const bool tree = Nt >= 1024 && n_iter > 1;
  if (tree) {
    k_icp_tbox<<<1, 1024, 0, st>>>(target, Nt, w);
    int gx = (w.P2 + 255) / 256;
    if (gx > 512) gx = 512;
    k_icp_tkeys<<<gx, 256, 0, st>>>(target, Nt, w);
    FOHO_LAUNCH_CHECK();
    int rc = foho_sort_u64(w.keys, w.P2, 1, st);
    if (rc != FOHO_OK) return rc;
    k_icp_tgather<<<gx, 256, 0, st>>>(target, Nt, Ns, w);
    k_icp_tboxes<<<(w.NG + 7) / 8, 256, 0, st>>>(Nt, 0, w);
    k_icp_tboxes<<<(w.NS + 7) / 8, 256, 0, st>>>(Nt, 1, w);
    FOHO_LAUNCH_CHECK();
  }
  const dim3 nn_grid((Ns + NN_THREADS - 1) / NN_THREADS, w.nchunks);
  for (int it = 0; it < n_iter; ++it) {
    int *nn_last = it == n_iter - 1 ? nn_index_last : nullptr;
    if (tree) k_icp_nn_tree<<<(Ns + 7) / 8, 256, 0, st>>>(source, Ns, Nt, w, nn_last);
    else k_icp_nn<<<nn_grid, NN_THREADS, 0, st>>>(source, Ns, target, Nt, w);
    k_icp_step<<<1, STEP_THREADS, 0, st>>>(target, Ns, n_outliers, fixed_scale, min_scale, max_scale, w, cost_history,
                                           tree ? nullptr : nn_last, tree ? 0 : 1);
  }
  FOHO_LAUNCH_CHECK();
  k_icp_finish<<<1, 32, 0, st>>>(w, transform_out, cost_out);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
