// Exact 1-NN chamfer between the transformed hand vertices and the MoGe cloud (NS row a15; same
// semantics as the reference's pytorch3d knn_points(K=1) call of
// third_party_patches/hy3dgen/shapegen/pipelines.py:1529-1532: squared L2 to the single nearest
// neighbour), with search structures that are built ONCE per image and reused by every one of the
// ~1200 evaluations of a guided diffusion:
//
//   cloud -> hand : the hand moves by a similarity, so "nearest transformed hand vertex of c" is
//                   "nearest REST vertex of c pulled back into the rest frame".  The rest vertices are
//                   Morton-sorted once into leaves of 8 with tight AABBs.  The cloud is stored
//                   Morton-sorted, so a warp's 32 points are neighbours: the warp prunes the leaves
//                   against the AABB of its 32 pulled-back points (dual-tree bound) and walks the
//                   survivors in lock step (k_chamfer_c2h).
//   hand -> cloud : the cloud never moves: its Morton-sorted points are boxed once in groups of 32 and
//                   super-groups of 32 groups; one warp per hand vertex descends that hierarchy,
//                   testing 32 boxes or 32 points per step (k_chamfer_h2c).
//
// Both are exact (not approximate) searches, stateless between evaluations; the brute-force k_chamfer
// in guidance_sparse.cu remains as the path used when the caller passes no accel buffer, and as the
// cross-check in the tests.
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

// ----------------------------------------------------------------------------- build: face order
// Faces in Morton order of their REST centroid (one CTA per sample, bitonic sort in shared memory):
// any 32 consecutive faces are then a compact patch of the hand, in every pose (the pose is a
// similarity), and face structures keyed by this order stay valid for every evaluation (k_raster writes the posed
// bounding spheres in this order, k_voxdist_staged reads the vertex ids in it).
__global__ void __launch_bounds__(ACC_THREADS) k_accel_faces(const float *__restrict__ hand_rest, const int *__restrict__ faces,
                                                             int Vh, int Fh, FohoAccel acc) {
  __shared__ unsigned long long key[FOHO_ACCEL_FACES];
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
  const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), fmaxf(bb[5] - bb[2], 1e-30f));
  const float q = 1023.f / ext;
  for (int f = tid; f < FOHO_ACCEL_FACES; f += blockDim.x) {
    unsigned long long k = ~0ull;
    if (f < Fh) {
      const int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
      const float cx = (rest[3 * ia] + rest[3 * ib] + rest[3 * ic]) * (1.f / 3.f);
      const float cy = (rest[3 * ia + 1] + rest[3 * ib + 1] + rest[3 * ic + 1]) * (1.f / 3.f);
      const float cz = (rest[3 * ia + 2] + rest[3 * ib + 2] + rest[3 * ic + 2]) * (1.f / 3.f);
      const unsigned ix = (unsigned)fminf(fmaxf((cx - bb[0]) * q, 0.f), 1023.f);
      const unsigned iy = (unsigned)fminf(fmaxf((cy - bb[1]) * q, 0.f), 1023.f);
      const unsigned iz = (unsigned)fminf(fmaxf((cz - bb[2]) * q, 0.f), 1023.f);
      k = ((unsigned long long)((spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz)) << 32) | (unsigned)f;
    }
    key[f] = k;
  }
  __syncthreads();
  for (int size = 2; size <= FOHO_ACCEL_FACES; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const int i = 2 * stride * (tid / stride) + (tid % stride), j = i + stride;
      const bool up = (i & size) == 0;
      const unsigned long long a = key[i], c = key[j];
      if ((a > c) == up) { key[i] = c; key[j] = a; }
      __syncthreads();
    }
  int4 *sv = acc.face_sv + (size_t)b * FOHO_ACCEL_FACES;
  int *rank = acc.face_rank + (size_t)b * FOHO_ACCEL_FACES;
  for (int s = tid; s < Fh; s += blockDim.x) {
    const int f = (int)(unsigned)key[s];
    sv[s] = make_int4(faces[3 * f], faces[3 * f + 1], faces[3 * f + 2], f);
    rank[f] = s;
  }
}

// ----------------------------------------------------------------------------- build: cloud order
// The cloud is put into 30-bit Morton order (10 bits per axis over its cubic bbox) by a bitonic sort
// of (code << 32 | index) keys, so that ANY run of consecutive points is spatially compact.  Runs
// once per image; not on the per-evaluation path.
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
    G->quant = 1023.f / ext;
    for (int a = 0; a < 3; ++a) G->origin[a] = lo[a];
    G->P = P; G->pad[0] = G->pad[1] = G->pad[2] = 0;
  }
}

__global__ void __launch_bounds__(256) k_accel_cloud_keys(const float *__restrict__ cloud, int P, int P2, FohoAccel acc) {
  const int b = blockIdx.y;
  const FohoAccelGrid G = acc.grid[b];
  const float *c = cloud + (size_t)b * P * 3;
  unsigned long long *keys = acc.keys + (size_t)b * P2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P2; i += gridDim.x * blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < P) {
      const unsigned ix = (unsigned)fminf(fmaxf((c[3 * (size_t)i] - G.origin[0]) * G.quant, 0.f), 1023.f);
      const unsigned iy = (unsigned)fminf(fmaxf((c[3 * (size_t)i + 1] - G.origin[1]) * G.quant, 0.f), 1023.f);
      const unsigned iz = (unsigned)fminf(fmaxf((c[3 * (size_t)i + 2] - G.origin[2]) * G.quant, 0.f), 1023.f);
      const unsigned m = (spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz);
      k = ((unsigned long long)m << 32) | (unsigned)i;          // unique keys: deterministic order
    }
    keys[i] = k;
  }
}

constexpr int SORT_CHUNK = 2048;          // keys per CTA in the shared-memory phases (1024 threads)

__device__ __forceinline__ void sort_cmpx(unsigned long long *sk, int i, int l, bool up) {
  const unsigned long long a = sk[i], c = sk[l];
  if ((a > c) == up) { sk[i] = c; sk[l] = a; }
}
// kmin == 2: full bitonic sort of each chunk; kmin == kmax > SORT_CHUNK: the j < SORT_CHUNK tail of merge step k
__global__ void __launch_bounds__(ACC_THREADS) k_sort_local(unsigned long long *__restrict__ keys_all, int P2, int kmin, int kmax) {
  __shared__ unsigned long long sk[SORT_CHUNK];
  unsigned long long *keys = keys_all + (size_t)blockIdx.y * P2 + (size_t)blockIdx.x * SORT_CHUNK;
  const int tid = threadIdx.x, gbase = blockIdx.x * SORT_CHUNK;
  sk[tid] = keys[tid]; sk[tid + ACC_THREADS] = keys[tid + ACC_THREADS];
  __syncthreads();
  for (int k = kmin; k <= kmax; k <<= 1) {
    for (int j = (k < SORT_CHUNK ? k : SORT_CHUNK) >> 1; j > 0; j >>= 1) {
      const int i = 2 * j * (tid / j) + (tid % j);
      sort_cmpx(sk, i, i + j, ((gbase + i) & k) == 0);
      __syncthreads();
    }
  }
  keys[tid] = sk[tid]; keys[tid + ACC_THREADS] = sk[tid + ACC_THREADS];
}
__global__ void __launch_bounds__(256) k_sort_global(unsigned long long *__restrict__ keys_all, int P2, int j, int k) {
  unsigned long long *keys = keys_all + (size_t)blockIdx.y * P2;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (P2 >> 1); t += gridDim.x * blockDim.x) {
    const int i = 2 * j * (t / j) + (t % j);
    sort_cmpx(keys, i, i + j, (i & k) == 0);
  }
}

__global__ void __launch_bounds__(256) k_accel_cloud_gather(const float *__restrict__ cloud, int P, int P2, FohoAccel acc) {
  const int b = blockIdx.y;
  const float *c = cloud + (size_t)b * P * 3;
  const unsigned long long *keys = acc.keys + (size_t)b * P2;
  float4 *pts = acc.pts + (size_t)b * P;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < P; s += gridDim.x * blockDim.x) {
    const unsigned i = (unsigned)keys[s];
    pts[s] = make_float4(c[3 * (size_t)i], c[3 * (size_t)i + 1], c[3 * (size_t)i + 2], __int_as_float((int)i));
  }
}

// ----------------------------------------------------------------------------- build: cloud boxes
// AABB of every group of 32 Morton-sorted cloud points, and of every super-group of 32 groups.
__global__ void __launch_bounds__(256) k_accel_cloud_boxes(int P, FohoAccel acc) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int NG = (P + 31) >> 5;
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= NG) return;
  const float4 *pts = acc.pts + (size_t)b * P;
  const int p = g * 32 + lane;
  const float4 c = pts[p < P ? p : g * 32];
  const float lx = warp_min(c.x), ly = warp_min(c.y), lz = warp_min(c.z);
  const float hx = warp_max(c.x), hy = warp_max(c.y), hz = warp_max(c.z);
  if (lane == 0) {
    acc.g_lo[(size_t)b * acc.NGcap + g] = make_float4(lx, ly, lz, 0.f);
    acc.g_hi[(size_t)b * acc.NGcap + g] = make_float4(hx, hy, hz, 0.f);
  }
}
__global__ void __launch_bounds__(256) k_accel_cloud_supers(int P, FohoAccel acc) {
  const int b = blockIdx.y, lane = threadIdx.x & 31;
  const int NG = (P + 31) >> 5, NS = (NG + 31) >> 5;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= NS) return;
  const int g = s * 32 + lane;
  const int gg = g < NG ? g : s * 32;
  const float4 lo = acc.g_lo[(size_t)b * acc.NGcap + gg], hi = acc.g_hi[(size_t)b * acc.NGcap + gg];
  const float lx = warp_min(lo.x), ly = warp_min(lo.y), lz = warp_min(lo.z);
  const float hx = warp_max(hi.x), hy = warp_max(hi.y), hz = warp_max(hi.z);
  if (lane == 0) {
    acc.s_lo[(size_t)b * acc.NScap + s] = make_float4(lx, ly, lz, 0.f);
    acc.s_hi[(size_t)b * acc.NScap + s] = make_float4(hx, hy, hz, 0.f);
  }
}

// ----------------------------------------------------------------------------- cloud -> hand
// One warp per group of 32 Morton-sorted (= spatially adjacent) cloud points; a dual-tree prune:
//   1. pull the 32 points back into the hand's rest frame, take their AABB G (warp min/max);
//   2. lanes evaluate, for the <=128 hand leaves, mindist^2(G, leaf box) and the upper bound
//      maxdist^2(G, first vertex of the leaf); U = min of the upper bounds bounds every lane's answer;
//   3. the leaf attaining U is searched first; U is then tightened to the largest per-lane best;
//   4. the warp walks the surviving leaves (mindist^2 <= U) in lock step; a lane skips a leaf whose
//      box is farther than its own best; the 8 vertices of a leaf are smem broadcasts.
// All bounds use the same monotone fp32 expression as the point distance, so pruning is exact in
// floating point, ties included (ties -> smallest original vertex index).
constexpr int C2H_THREADS = 256;
constexpr int C2H_GROUPS_PER_WARP = 4;
constexpr int C2H_GROUPS_PER_CTA = (C2H_THREADS / 32) * C2H_GROUPS_PER_WARP;

__device__ __forceinline__ float box_dist2(float4 lo, float4 hi, float qx, float qy, float qz) {
  const float dx = fmaxf(fmaxf(lo.x - qx, qx - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - qy, qy - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - qz, qz - hi.z), 0.f);
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
// min squared distance between two boxes
__device__ __forceinline__ float boxbox_min2(float4 lo, float4 hi, const float (&glo)[3], const float (&ghi)[3]) {
  const float dx = fmaxf(fmaxf(lo.x - ghi[0], glo[0] - hi.x), 0.f);
  const float dy = fmaxf(fmaxf(lo.y - ghi[1], glo[1] - hi.y), 0.f);
  const float dz = fmaxf(fmaxf(lo.z - ghi[2], glo[2] - hi.z), 0.f);
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}
// max squared distance from a point to any point of a box
__device__ __forceinline__ float box_max2(float4 v, const float (&glo)[3], const float (&ghi)[3]) {
  const float dx = fmaxf(fabsf(v.x - glo[0]), fabsf(ghi[0] - v.x));
  const float dy = fmaxf(fabsf(v.y - glo[1]), fabsf(ghi[1] - v.y));
  const float dz = fmaxf(fabsf(v.z - glo[2]), fabsf(ghi[2] - v.z));
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// Add (gx,gy,gz) to gacc[3*j..] for every valid lane, one shared-memory atomic triple per DISTINCT j in
// the warp: neighbouring cloud points mostly share their nearest vertex, and float atomics on one
// shared address serialise (CAS loop).  Must be called by all 32 lanes.
__device__ __forceinline__ void warp_scatter_add3(float *gacc, bool valid, int j, float gx, float gy, float gz) {
  const int lane = threadIdx.x & 31;
  unsigned todo = __ballot_sync(0xffffffffu, valid);
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int key = __shfl_sync(0xffffffffu, j, leader);
    const bool mine = valid && j == key;
    const unsigned grp = __ballot_sync(0xffffffffu, mine);
    const float sx = warp_sum(mine ? gx : 0.f), sy = warp_sum(mine ? gy : 0.f), sz = warp_sum(mine ? gz : 0.f);
    if (lane == leader) { atomicAdd(gacc + 3 * key, sx); atomicAdd(gacc + 3 * key + 1, sy); atomicAdd(gacc + 3 * key + 2, sz); }
    todo &= ~grp;
  }
}

struct C2HBest { float d2; int slot; };

// 8 vertices of one leaf: all distances, a min tree, and (rarely) the arg-min
__device__ __forceinline__ void c2h_scan_leaf(const float4 *__restrict__ v, int l, float qx, float qy, float qz, C2HBest &bst) {
  float dd[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float4 p = v[l * 8 + k];
    const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
    dd[k] = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  }
  const float m = fminf(fminf(fminf(dd[0], dd[1]), fminf(dd[2], dd[3])), fminf(fminf(dd[4], dd[5]), fminf(dd[6], dd[7])));
  if (m < bst.d2) {
    int k = 7;
#pragma unroll
    for (int t = 6; t >= 0; --t) k = dd[t] == m ? t : k;
    bst.d2 = m; bst.slot = l * 8 + k;
  }
}

__global__ void __launch_bounds__(C2H_THREADS) k_chamfer_c2h(foho_guidance_desc d, FohoWorkspace ws, FohoAccel acc) {
  FohoTrace trace_(ws.trace, TR_C2H);
  __shared__ float4 sv[FOHO_ACCEL_HV];
  __shared__ float4 s_llo[FOHO_ACCEL_LEAVES], s_lhi[FOHO_ACCEL_LEAVES];
  __shared__ float gacc[FOHO_ACCEL_HV * 3];
  __shared__ float red[32];
  const int b = blockIdx.y, Vh = d.Vh, P = d.P, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int NG = (P + 31) >> 5;
  const int g0 = blockIdx.x * C2H_GROUPS_PER_CTA;
  if (g0 >= NG) return;
  const FohoAccelHand *H = acc.hand + b;
  const int nleaf = H->n_leaves, nsup = H->n_supers;
  const int nleafpad = nsup * 8;
  for (int i = tid; i < nsup * 64; i += blockDim.x) sv[i] = H->v[i];
  for (int i = tid; i < FOHO_ACCEL_LEAVES; i += blockDim.x) {
    if (i < nleafpad) { s_llo[i] = H->leaf_lo[i]; s_lhi[i] = H->leaf_hi[i]; }
    else { s_llo[i] = make_float4(INFINITY, INFINITY, INFINITY, 0.f); s_lhi[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f); }
  }
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
#pragma unroll 1
  for (int u = 0; u < C2H_GROUPS_PER_WARP; ++u) {
    const int g = g0 + wid * C2H_GROUPS_PER_WARP + u;
    if (g >= NG) break;
    const int pi = g * 32 + lane;
    const bool valid = pi < P;
    const float4 c4 = pts[valid ? pi : g * 32];
    const float px = c4.x - cx, py = c4.y - cy, pz = c4.z - cz;          // centred MoGe
    const float rx = px - ox, ry = py - oy, rz = pz - oz;
    const float qx = Rt[0] * rx + Rt[1] * ry + Rt[2] * rz;
    const float qy = Rt[3] * rx + Rt[4] * ry + Rt[5] * rz;
    const float qz = Rt[6] * rx + Rt[7] * ry + Rt[8] * rz;
    float glo[3] = {warp_min(qx), warp_min(qy), warp_min(qz)};
    float ghi[3] = {warp_max(qx), warp_max(qy), warp_max(qz)};
    // warm start: the slot found by the previous evaluation (-1 right after prepare_statics).  It only
    // tightens the pruning radius -- the search below is exact whatever the seed is.
    int *seedp = acc.seed_c2h + (size_t)b * P + (valid ? pi : g * 32);
    const int seed = *seedp;
    const bool seeded = __all_sync(0xffffffffu, seed >= 0);
    C2HBest bst; bst.d2 = INFINITY; bst.slot = 0;
    float lb[4];
    int l0 = -1;
    if (seeded) {
      const float4 p = sv[seed];
      const float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
      bst.d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx)); bst.slot = seed;
#pragma unroll
      for (int r = 0; r < 4; ++r) lb[r] = boxbox_min2(s_llo[r * 32 + lane], s_lhi[r * 32 + lane], glo, ghi);
    } else {
      // bounds of every hand leaf against the group box; the leaf with the smallest upper bound first
      float ub = INFINITY;
      int ubl = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int l = r * 32 + lane;
        lb[r] = boxbox_min2(s_llo[l], s_lhi[l], glo, ghi);        // +inf for pad leaves
        if (l < nleaf) {
          const float m = box_max2(sv[l * 8], glo, ghi);
          if (m < ub) { ub = m; ubl = l; }
        }
      }
      const unsigned ubits = __float_as_uint(ub);
      const unsigned umin = __reduce_min_sync(0xffffffffu, ubits);
      const unsigned who = __ballot_sync(0xffffffffu, ubits == umin);
      l0 = __shfl_sync(0xffffffffu, ubl, __ffs(who) - 1);
      c2h_scan_leaf(sv, l0, qx, qy, qz, bst);
    }
    const float U = warp_max(bst.d2);                             // every lane's answer is <= its best <= U
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      unsigned mask = __ballot_sync(0xffffffffu, lb[r] <= U);
      while (mask) {
        const int l = r * 32 + __ffs(mask) - 1;
        mask &= mask - 1;
        if (l == l0) continue;
        if (box_dist2(s_llo[l], s_lhi[l], qx, qy, qz) < bst.d2) c2h_scan_leaf(sv, l, qx, qy, qz, bst);
      }
    }
    if (valid && bst.slot != seed) *seedp = bst.slot;
    {
      // exact squared distance in the (centred) MoGe frame, as the oracle evaluates it
      const int j = __float_as_int(sv[bst.slot].w);
      const float hx = hmc[3 * j], hy = hmc[3 * j + 1], hz = hmc[3 * j + 2];
      const float dx = hx - px, dy = hy - py, dz = hz - pz;
      if (valid) sum += fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      warp_scatter_add3(gacc, valid, j, coef * dx, coef * dy, coef * dz);
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
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
}

// ----------------------------------------------------------------------------- cloud -> hand, graph walk
// When the caller supplies the Delaunay neighbour graph of the REST hand vertices (desc->hand_nbr*,
// built once per image on the host), the nearest rest vertex of a pulled-back cloud point is found by
// greedy descent: from the vertex found last time, move to the neighbour closest to the query until no
// neighbour is closer.  On a Delaunay graph that stops only at the true nearest vertex (every
// non-nearest vertex has a Delaunay neighbour closer to the query), so the result is exact; with the
// warm start it takes ~0.3 moves and ~27 distance evaluations per point instead of the ~400 of the box
// search, which matters because most cloud points are far from the hand, where boxes prune poorly.
constexpr int WALK_THREADS = 256;
// Per-vertex gradient sums are accumulated per CTA in fixed point (2^-33 units, |h - p| < 64) split over
// two 32-bit shared-memory words -- hi = v >> 20 (signed), lo = v & 0xFFFFF -- because only 32-bit
// INTEGER shared atomics are native (float and 64-bit ones are CAS spin loops, and neighbouring cloud
// points mostly hit the same vertex).  A CTA adds at most 4096 points, so neither word can overflow;
// integer addition is associative, so the per-CTA sums do not depend on the arrival order.
constexpr float WALK_FIX = 8589934592.f;             // 2^33
constexpr float WALK_UNFIX = 1.f / 8589934592.f;
constexpr int WALK_MAX_CHUNK = 4096;
__device__ __forceinline__ void walk_fix_add(int *hi, unsigned *lo, float x, int *flags) {
  if (!(fabsf(x) <= 63.f)) atomicOr(flags, 2);        // out of the fixed-point range: reported in terms[FLAGS]
  const long long v = __float2ll_rn(fminf(fmaxf(x, -63.f), 63.f) * WALK_FIX);
  atomicAdd(hi, (int)(v >> 20));
  atomicAdd(lo, (unsigned)(v & 0xFFFFF));
}

__global__ void __launch_bounds__(WALK_THREADS) k_chamfer_c2h_walk(foho_guidance_desc d, FohoWorkspace ws, FohoAccel acc,
                                                                     int chunk) {
  FohoTrace trace_(ws.trace, TR_C2H);
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.y, Vh = d.Vh, P = d.P, tid = threadIdx.x;
  const int p0 = blockIdx.x * chunk;
  if (p0 >= P) return;
  const int p1 = min(P, p0 + chunk);
  const int *goff = d.hand_nbr_off + (size_t)b * (Vh + 1);
  const uint16_t *gadj = d.hand_nbr + (size_t)b * d.nbr_stride;
  const int E = goff[Vh];
  float4 *sv = reinterpret_cast<float4 *>(sm_raw);                                   // [Vh] rest verts - bbox centre
  int *ghi = reinterpret_cast<int *>(sv + Vh);                                       // [Vh*3] fixed point, high part
  unsigned *glo = reinterpret_cast<unsigned *>(ghi + 3 * Vh);                        // [Vh*3] low 20 bits
  int *soff = reinterpret_cast<int *>(glo + 3 * Vh);                                 // [Vh+1]
  uint16_t *sadj = reinterpret_cast<uint16_t *>(sm_raw + foho_align_up((size_t)Vh * 40 + (size_t)(Vh + 1) * 4, 16));   // [E]
  __shared__ float red[32];
  const FohoFrame &fr = ws.frames[b];
  const float *rest = d.hand_rest + (size_t)b * Vh * 3;
  for (int i = tid; i < Vh; i += blockDim.x) {
    sv[i] = make_float4(rest[3 * i] - fr.ch[0], rest[3 * i + 1] - fr.ch[1], rest[3 * i + 2] - fr.ch[2], 0.f);
    ghi[3 * i] = 0; ghi[3 * i + 1] = 0; ghi[3 * i + 2] = 0;
    glo[3 * i] = 0u; glo[3 * i + 1] = 0u; glo[3 * i + 2] = 0u;
  }
  for (int i = tid; i <= Vh; i += blockDim.x) soff[i] = goff[i];
  {
    // 16-byte copies: nbr_stride is a multiple of 8 and both bases are 16-byte aligned (checked by the launcher)
    const uint4 *g4 = reinterpret_cast<const uint4 *>(gadj);
    uint4 *s4 = reinterpret_cast<uint4 *>(sadj);
    const int n4 = (E + 7) >> 3;
    for (int i = tid; i < n4; i += blockDim.x) s4[i] = g4[i];
  }
  __syncthreads();
  const float ox = fr.chc[0] + fr.th[0], oy = fr.chc[1] + fr.th[1], oz = fr.chc[2] + fr.th[2];
  const float is = 1.f / fr.sh;
  float Rt[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rt[3 * r + c] = fr.Rh[3 * c + r] * is;
  const float cx = fr.co[0], cy = fr.co[1], cz = fr.co[2];
  const float4 *pts = acc.pts + (size_t)b * P;
  int *seeds = acc.seed_c2h + (size_t)b * P;
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;
  float sum = 0.f;
  int prev = 0;
  for (int pi = p0 + tid; pi < p1; pi += blockDim.x) {
    const float4 c4 = pts[pi];
    const int seed = seeds[pi];
    const float px = c4.x - cx, py = c4.y - cy, pz = c4.z - cz;          // centred MoGe
    const float rx = px - ox, ry = py - oy, rz = pz - oz;
    const float qx = Rt[0] * rx + Rt[1] * ry + Rt[2] * rz;
    const float qy = Rt[3] * rx + Rt[4] * ry + Rt[5] * rz;
    const float qz = Rt[6] * rx + Rt[7] * ry + Rt[8] * rz;
    int cur = (seed >= 0 && seed < Vh) ? seed : prev;                    // cold start: this thread's previous answer
    float dc;
    {
      const float4 v = sv[cur];
      const float dx = qx - v.x, dy = qy - v.y, dz = qz - v.z;
      dc = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    }
    while (true) {
      const int e0 = soff[cur], e1 = soff[cur + 1];
      int best = cur;
      float db = dc;
#pragma unroll 4
      for (int e = e0; e < e1; ++e) {
        const int n = sadj[e];
        const float4 v = sv[n];
        const float dx = qx - v.x, dy = qy - v.y, dz = qz - v.z;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2 < db) { db = d2; best = n; }
      }
      if (best == cur) break;
      cur = best; dc = db;
    }
    prev = cur;
    if (cur != seed) seeds[pi] = cur;
    // exact squared distance in the (centred) MoGe frame, as the oracle evaluates it
    const float hx = hmc[3 * cur], hy = hmc[3 * cur + 1], hz = hmc[3 * cur + 2];
    const float dx = hx - px, dy = hy - py, dz = hz - pz;
    sum += fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    int *flags = ws.cnt + (size_t)b * CNT_NUM + CNT_FLAGS;
    walk_fix_add(ghi + 3 * cur, glo + 3 * cur, dx, flags);
    walk_fix_add(ghi + 3 * cur + 1, glo + 3 * cur + 1, dy, flags);
    walk_fix_add(ghi + 3 * cur + 2, glo + 3 * cur + 2, dz, flags);
  }
  sum = warp_sum(sum);
  if ((tid & 31) == 0) red[tid >> 5] = sum;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < WALK_THREADS / 32; ++w) t += red[w];
    atomicAdd(ws.acc + (size_t)b * ACC_NUM + ACC_CH_CLOUD, t);
  }
  const float coef = 2.f * d.w.w_ch / (float)P * WALK_UNFIX;
  float *Ghm = ws.G_hm + (size_t)b * Vh * 3;
  for (int i = tid; i < Vh * 3; i += blockDim.x) {
    const long long g = ((long long)ghi[i] << 20) + (long long)glo[i];
    if (g != 0) atomicAdd(Ghm + i, coef * (float)g);
  }
}

// ----------------------------------------------------------------------------- hand -> cloud
// One warp per hand vertex over the static two-level box hierarchy of the cloud (groups of 32
// Morton-sorted points, super-groups of 32 groups): a greedy descent seeds the radius, then every
// super-group / group whose box is within the current best is scanned, 32 boxes or points per step.
constexpr int H2C_THREADS = 256;

struct H2CBest { unsigned d2; unsigned idx; int pos; };
// fold the 32 lane candidates (d2 bits, original index) into the warp-uniform best (ties -> smallest index)
__device__ __forceinline__ void h2c_fold(H2CBest &bst, unsigned d2bits, unsigned idx, int pos) {
  const unsigned wm = __reduce_min_sync(0xffffffffu, d2bits);
  if (wm > bst.d2) return;                                        // warp-uniform
  const unsigned im = __reduce_min_sync(0xffffffffu, d2bits == wm ? idx : 0xffffffffu);
  if (wm < bst.d2 || im < bst.idx) {
    bst.d2 = wm; bst.idx = im;
    const unsigned who = __ballot_sync(0xffffffffu, d2bits == wm && idx == im);
    bst.pos = __shfl_sync(0xffffffffu, pos, __ffs(who) - 1);
  }
}

__global__ void __launch_bounds__(H2C_THREADS) k_chamfer_h2c(foho_guidance_desc d, FohoWorkspace ws, FohoAccel acc) {
  FohoTrace trace_(ws.trace, TR_H2C);
  const int b = blockIdx.y, Vh = d.Vh, P = d.P;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = blockIdx.x * (H2C_THREADS / 32) + wid;
  if (i >= Vh) return;
  const int NG = (P + 31) >> 5, NS = (NG + 31) >> 5;
  const FohoFrame &fr = ws.frames[b];
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;
  // absolute MoGe position of this hand vertex (the cloud boxes live in absolute coordinates)
  const float hx = hmc[3 * i] + fr.co[0], hy = hmc[3 * i + 1] + fr.co[1], hz = hmc[3 * i + 2] + fr.co[2];
  const float4 *pts = acc.pts + (size_t)b * P;
  const float4 *glo = acc.g_lo + (size_t)b * acc.NGcap, *ghi = acc.g_hi + (size_t)b * acc.NGcap;
  const float4 *slo = acc.s_lo + (size_t)b * acc.NScap, *shi = acc.s_hi + (size_t)b * acc.NScap;
  const unsigned INF_BITS = 0x7f800000u;
  H2CBest bst; bst.d2 = 0xffffffffu; bst.idx = 0xffffffffu; bst.pos = -1;

  auto scan_group = [&](int g) {
    const int p = g * 32 + lane;
    unsigned db = 0xffffffffu, ix = 0xffffffffu;
    if (p < P) {
      const float4 c = pts[p];
      const float ex = hx - c.x, ey = hy - c.y, ez = hz - c.z;
      db = __float_as_uint(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
      ix = (unsigned)__float_as_int(c.w);
    }
    h2c_fold(bst, db, ix, p);
  };
  auto scan_super = [&](int s, int skip_g) {
    const int g = s * 32 + lane;
    unsigned lb = 0xffffffffu;
    if (g < NG && g != skip_g) lb = __float_as_uint(box_dist2(glo[g], ghi[g], hx, hy, hz));
    unsigned mask = __ballot_sync(0xffffffffu, lb <= bst.d2 && lb <= INF_BITS);
    while (mask) {
      const int gl = __ffs(mask) - 1;
      scan_group(s * 32 + gl);
      mask &= mask - 1;
      mask &= __ballot_sync(0xffffffffu, lb <= bst.d2);
    }
  };

  // warm start: the sorted position found by the previous evaluation (-1 right after prepare_statics);
  // otherwise a greedy descent: nearest super-group -> its nearest group -> scan
  int s0 = -1, gseed = -1;
  int *seedp = acc.seed_h2c + (size_t)b * Vh + i;
  const int seed = *seedp;
  if (seed >= 0 && seed < P) {
    scan_group(seed >> 5);
  } else {
    unsigned bl = 0xffffffffu; int bs = 0;
    for (int s = lane; s < NS; s += 32) {
      const unsigned v = __float_as_uint(box_dist2(slo[s], shi[s], hx, hy, hz));
      if (v < bl) { bl = v; bs = s; }
    }
    const unsigned m = __reduce_min_sync(0xffffffffu, bl);
    s0 = __shfl_sync(0xffffffffu, bs, __ffs(__ballot_sync(0xffffffffu, bl == m)) - 1);
    const int g = s0 * 32 + lane;
    const unsigned lb = g < NG ? __float_as_uint(box_dist2(glo[g], ghi[g], hx, hy, hz)) : 0xffffffffu;
    const unsigned m2 = __reduce_min_sync(0xffffffffu, lb);
    gseed = s0 * 32 + __ffs(__ballot_sync(0xffffffffu, lb == m2)) - 1;
    scan_group(gseed);
    scan_super(s0, gseed);
  }
  // exact pass over the remaining super-groups
  for (int sb = 0; sb < NS; sb += 32) {
    const int s = sb + lane;
    unsigned lb = 0xffffffffu;
    if (s < NS && s != s0) lb = __float_as_uint(box_dist2(slo[s], shi[s], hx, hy, hz));
    unsigned mask = __ballot_sync(0xffffffffu, lb <= bst.d2 && lb <= INF_BITS);
    while (mask) {
      const int sl = __ffs(mask) - 1;
      scan_super(sb + sl, -1);
      mask &= mask - 1;
      mask &= __ballot_sync(0xffffffffu, lb <= bst.d2);
    }
  }
  if (lane == 0) {
    ws.knn[(size_t)b * Vh + i] = bst.idx != 0xffffffffu ? (((unsigned long long)bst.d2 << 32) | bst.idx) : ~0ull;
    if (bst.pos != seed) *seedp = bst.pos;
  }
}

}  // namespace

// ----------------------------------------------------------------------------- host side
void foho_accel_layout(FohoAccel &a, char *base, int B, int P) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += foho_align_up(bytes, 256); return p; };
  a.hand = (FohoAccelHand *)take(sizeof(FohoAccelHand) * (size_t)B);
  a.grid = (FohoAccelGrid *)take(sizeof(FohoAccelGrid) * (size_t)B);
  a.P2 = SORT_CHUNK;
  while (a.P2 < P) a.P2 <<= 1;
  a.keys = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * a.P2);
  a.pts = (float4 *)take(sizeof(float4) * (size_t)B * (P > 0 ? P : 1));
  a.NGcap = ((P > 0 ? P : 1) + 31) / 32;
  a.NScap = (a.NGcap + 31) / 32;
  a.g_lo = (float4 *)take(sizeof(float4) * (size_t)B * a.NGcap);
  a.g_hi = (float4 *)take(sizeof(float4) * (size_t)B * a.NGcap);
  a.s_lo = (float4 *)take(sizeof(float4) * (size_t)B * a.NScap);
  a.s_hi = (float4 *)take(sizeof(float4) * (size_t)B * a.NScap);
  a.face_sv = (int4 *)take(sizeof(int4) * (size_t)B * FOHO_ACCEL_FACES);
  a.face_rank = (int *)take(sizeof(int) * (size_t)B * FOHO_ACCEL_FACES);
  a.seed_c2h = (int *)take(sizeof(int) * (size_t)B * (P > 0 ? P : 1));
  a.seed_h2c = (int *)take(sizeof(int) * (size_t)B * FOHO_ACCEL_HV);
  a.total = off;
}

// bitonic sort of `batch` arrays of P2 (a power of two >= 2048) 64-bit keys, ascending
int foho_sort_u64(unsigned long long *keys, int P2, int batch, cudaStream_t st) {
  if (P2 < SORT_CHUNK || (P2 & (P2 - 1)) != 0) return FOHO_E_SHAPE;
  int gx = (P2 + 255) / 256;
  if (gx > 512) gx = 512;
  const int nchunk = P2 / SORT_CHUNK;
  k_sort_local<<<dim3(nchunk, batch), ACC_THREADS, 0, st>>>(keys, P2, 2, SORT_CHUNK);
  FOHO_LAUNCH_CHECK();
  for (int k = SORT_CHUNK * 2; k <= P2; k <<= 1) {
    for (int j = k >> 1; j >= SORT_CHUNK; j >>= 1) {
      k_sort_global<<<dim3(gx, batch), 256, 0, st>>>(keys, P2, j, k);
      FOHO_LAUNCH_CHECK();
    }
    k_sort_local<<<dim3(nchunk, batch), ACC_THREADS, 0, st>>>(keys, P2, k, k);
    FOHO_LAUNCH_CHECK();
  }
  return FOHO_OK;
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
  if (d.hand_faces && d.Fh >= 1 && d.Fh <= FOHO_ACCEL_FACES) {
    k_accel_faces<<<d.B, ACC_THREADS, 0, st>>>(d.hand_rest, d.hand_faces, d.Vh, d.Fh, a);
    FOHO_LAUNCH_CHECK();
  } else if (d.Fh >= 1 && d.Fh <= FOHO_ACCEL_FACES) {
    return FOHO_E_NULL;          // evaluations with this accel would expect the face order
  }
  k_accel_cloud_bbox<<<d.B, ACC_THREADS, 0, st>>>(d.cloud, d.P, a);
  FOHO_LAUNCH_CHECK();
  int gx = (a.P2 + 255) / 256;
  if (gx > 512) gx = 512;
  k_accel_cloud_keys<<<dim3(gx, d.B), 256, 0, st>>>(d.cloud, d.P, a.P2, a);
  FOHO_LAUNCH_CHECK();
  {
    int rc = foho_sort_u64(a.keys, a.P2, d.B, st);
    if (rc != FOHO_OK) return rc;
  }
  k_accel_cloud_gather<<<dim3(gx, d.B), 256, 0, st>>>(d.cloud, d.P, a.P2, a);
  FOHO_LAUNCH_CHECK();
  k_accel_cloud_boxes<<<dim3((a.NGcap + 7) / 8, d.B), 256, 0, st>>>(d.P, a);
  FOHO_LAUNCH_CHECK();
  k_accel_cloud_supers<<<dim3((a.NScap + 7) / 8, d.B), 256, 0, st>>>(d.P, a);
  FOHO_LAUNCH_CHECK();
  // no warm-start hints yet (0xff.. = -1)
  FOHO_CUDA_TRY(cudaMemsetAsync(a.seed_c2h, 0xff, sizeof(int) * (size_t)d.B * d.P, st));
  FOHO_CUDA_TRY(cudaMemsetAsync(a.seed_h2c, 0xff, sizeof(int) * (size_t)d.B * FOHO_ACCEL_HV, st));
  return FOHO_OK;
}

int foho_launch_chamfer_h2c(const foho_guidance_desc *dp, const FohoWorkspace &ws, cudaStream_t st) {
  const foho_guidance_desc &d = *dp;
  FohoAccel a;
  foho_accel_layout(a, (char *)d.accel, d.B, d.P);
  if (a.total > d.accel_bytes) return FOHO_E_WORKSPACE;
  k_chamfer_h2c<<<dim3((d.Vh + H2C_THREADS / 32 - 1) / (H2C_THREADS / 32), d.B), H2C_THREADS, 0, st>>>(d, ws, a);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

int foho_launch_chamfer_c2h(const foho_guidance_desc *dp, const FohoWorkspace &ws, cudaStream_t st) {
  const foho_guidance_desc &d = *dp;
  FohoAccel a;
  foho_accel_layout(a, (char *)d.accel, d.B, d.P);
  if (a.total > d.accel_bytes) return FOHO_E_WORKSPACE;
  if (d.hand_nbr_off && d.hand_nbr) {
    if (d.nbr_stride < 8 || (d.nbr_stride & 7) != 0 || d.nbr_stride > 65535 * 64) return FOHO_E_ARG;
    if (((uintptr_t)d.hand_nbr & 15) != 0) return FOHO_E_ARG;
    // whole neighbour lists in shared memory: size for the worst case the stride allows
    const size_t smem = foho_align_up((size_t)d.Vh * 40 + (size_t)(d.Vh + 1) * 4, 16) + (size_t)d.nbr_stride * 2;
    if (smem > 200 * 1024) return FOHO_E_SHAPE;
    {
      int rc = foho_func_attrs((const void *)k_chamfer_c2h_walk, FA_C2H_WALK, smem, true);
      if (rc != FOHO_OK) return rc;
    }
    // At most one CTA per SM over the whole batch: beside the dense stream's CTA an SM has room
    // for exactly one of these, and a grid that does not fit at once would hold back every kernel
    // queued behind it.  Each CTA walks a contiguous chunk of the Morton-sorted cloud.
    FohoDeviceState *ds = foho_device_state();
    if (!ds) return (int)cudaGetLastError();
    const int sm_count = ds->sm_count;
    int gx = sm_count / d.B;
    if (gx < 1) gx = 1;
    int chunk = (d.P + gx - 1) / gx;
    chunk = (chunk + WALK_THREADS - 1) / WALK_THREADS * WALK_THREADS;
    if (chunk > WALK_MAX_CHUNK) chunk = WALK_MAX_CHUNK;               // overflow bound of the fixed-point sums
    gx = (d.P + chunk - 1) / chunk;
    k_chamfer_c2h_walk<<<dim3(gx, d.B), WALK_THREADS, smem, st>>>(d, ws, a, chunk);
    FOHO_LAUNCH_CHECK();
    return FOHO_OK;
  }
  const int NG = (d.P + 31) / 32;
  {
    int rc = foho_func_attrs((const void *)k_chamfer_c2h, FA_C2H, 0, true);
    if (rc != FOHO_OK) return rc;
  }
  k_chamfer_c2h<<<dim3((NG + C2H_GROUPS_PER_CTA - 1) / C2H_GROUPS_PER_CTA, d.B), C2H_THREADS, 0, st>>>(d, ws, a);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
