// Internal declarations shared by the .cu translation units (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/foho_b200.h"
#include "foho_math.cuh"

#define FOHO_CUDA_TRY(expr)                      \
  do {                                           \
    cudaError_t _e = (expr);                     \
    if (_e != cudaSuccess) return (int)_e;       \
  } while (0)

#define FOHO_LAUNCH_CHECK()                      \
  do {                                           \
    cudaError_t _e = cudaGetLastError();         \
    if (_e != cudaSuccess) return (int)_e;       \
  } while (0)

// Per-sample geometry of one evaluation, written by k_prep, read by everything else.
struct FohoFrame {
  float Rh[9], sh, th[3], ch[3];   // hand similarity about the rest bbox centre ch
  float chc[3];                    // ch - c_o
  float Ro[9], so, to[3], co[3];   // object similarity about c_o
  float Ah[9];                     // 3x3 of T_h2m (scale * rotation)
  float u0[3];                     // A_h(-bound*1) + t_h2m - c_o
  float s_h2m, step;               // |A_h[:,0]|, lattice spacing in Hunyuan units
  float A[9];                      // lattice -> MoGe linear part  s_o R_o A_h step
  float bc[3];                     // lattice origin in centred MoGe coords: s_o R_o u0 + t_o
  float Ainv[9];                   // inverse of A
  float Ahs_inv[9];                // inverse of (A_h * step)
  float kappa;                     // MoGe length of one lattice step
  float e[3], f;                   // |y|^2 = kappa^2 |g|^2 + 2 e.g + f   (absolute MoGe frame)
  int lo[3], hi[3];                // hand bbox on the lattice (inclusive, clipped; lo>hi = empty)
  int pad[2];
};

// per-sample accumulators (floats) zeroed by k_prep
enum {
  ACC_INT = 0,      // sum over candidates of (-S) * d_grid
  ACC_GKAPPA = 1,   // dE/dkappa
  ACC_CH_CLOUD = 2, // sum over cloud points of d2 to nearest hand vertex
  ACC_DIST = 3,     // sum over hand verts of clamp(d2 - margin, 0) to the object mesh (REF a7)
  ACC_MEAN_D2 = 4,  // sum over hand verts of d2 to the object mesh
  ACC_VREG = 5,     // sum over object verts of |ot|^2 (REF a10)
  ACC_EDGE = 6,     // sum over object edges of |ot0 - ot1|^2 (REF a10)
  ACC_NUM = 8
};
enum { CNT_NCAND = 0, CNT_COUNT = 1, CNT_FLAGS = 2, CNT_NUM = 4 };

#define FOHO_STREAM_PARTIALS 8      // m0, m1x, m1y, m1z, m2, count_obj, pad, pad
#define FOHO_MAX_STREAM_CTAS 2048   // per sample
#define FOHO_VE_MAX_CTAS 16         // k_vertex_early CTAs per sample (Vh <= 4096)
#define FOHO_FIN_NRED 32            // per-sample sums produced by k_finalize_verts

// Per-sample state of the explicit object mesh (REF a5/a6 on the FlexiCubes vertices).
struct FohoObjInfo {
  float c[3];                 // bbox centre of T_h2m(obj verts)  (pipelines.py:111, current verts)
  int amin[3], amax[3];       // packed indices of the arg-min / arg-max vertex per axis
  int v0, v1, e0, e1;         // packed vertex / edge ranges of this sample
  int pad;
};

struct FohoWorkspace {
  FohoFrame *frames;          // [B]
  float *hmc;                 // [B,Vh,3] transformed hand verts, centred on c_o
  float *hg;                  // [B,Vh,3] same in lattice units
  float *G_hm;                // [B,Vh,3] dE/d(hm) (MoGe)
  float *G_hg;                // [B,Vh,3] dE/d(hg) (lattice)
  float *acc;                 // [B,ACC_NUM]
  int *cnt;                   // [B,CNT_NUM]
  unsigned long long *knn;    // [B,Vh] packed (d2 bits << 32 | cloud index)
  float *stream_part;         // [B,FOHO_MAX_STREAM_CTAS,FOHO_STREAM_PARTIALS]
  uint32_t *parity;           // [B,D*D*W]
  int *cand;                  // [B,cap]
  float *ot;                  // [Vo,3] transformed object verts, centred on c_o
  float *g_ot;                // [Vo,3] dE/d(ot)
  unsigned long long *knn_obj;// [B,Vh] packed (d2 bits << 32 | packed object vertex index)
  FohoObjInfo *oinfo;         // [B]
  float4 *sph;                // [B,Fh] bounding sphere (centroid, radius) of each hand face, lattice units
  float *E_hm, *E_hg;         // [B,Vh,3] k_vertex_early: key-point/external part of dE/d(hm), field part of dE/d(hg)
  float *pen_part;            // [B,FOHO_VE_MAX_CTAS,2] per-CTA sums of the a13 penalties
  float *kpbuf;               // [B,64] k_keypoints -> k_finalize_verts: dE/d(21 key-points) | loss
  float *fin_acc;             // [B,FOHO_FIN_NRED] per-sample sums of k_finalize_verts
  float *cand_val;            // [B,cap] dE/dS contribution of each candidate voxel (applied by k_assemble)
  int *tri_idx;               // [B,Vh,8] voxel index of the trilinear corners of each vertex sample
  float *tri_val;             // [B,Vh,8] dE/dS contribution at those corners (0 = none)
  unsigned long long *trace;  // optional (desc->trace): per kernel [first CTA start, last CTA end], globaltimer ns
  int cap;
  int W;                      // words per column
  size_t total;
};

// ---- per-image search structures for the chamfer term (guidance_chamfer.cu), built once by
//      foho_guidance_prepare_statics into the caller's `accel` buffer
#define FOHO_ACCEL_HV 1024        // max hand vertices the structured search handles
#define FOHO_ACCEL_LEAVES 128     // leaves of 8 Morton-consecutive rest vertices
#define FOHO_ACCEL_SUPERS 16      // super-boxes of 8 leaves
#define FOHO_ACCEL_FACES 2048     // max hand faces the grouped point->mesh search handles

struct FohoAccelHand {
  float4 v[FOHO_ACCEL_HV];                 // sorted rest verts relative to the rest bbox centre; w = original index
  float4 leaf_lo[FOHO_ACCEL_LEAVES], leaf_hi[FOHO_ACCEL_LEAVES];
  float4 sup_lo[FOHO_ACCEL_SUPERS], sup_hi[FOHO_ACCEL_SUPERS];
  int n_leaves, n_supers, Vh, pad;
};
struct FohoAccelGrid {
  float origin[3];                         // min corner of the cloud bbox (absolute MoGe)
  float quant;                             // 1023 / largest bbox extent: Morton quantisation
  int P, pad[3];
};
struct FohoAccel {
  FohoAccelHand *hand;        // [B]
  FohoAccelGrid *grid;        // [B]
  unsigned long long *keys;   // [B,P2] build scratch: (Morton code << 32 | index), sorted
  int P2;                     // P padded to a power of two >= 2048
  float4 *pts;                // [B,P] cloud in Morton order; w = original index
  float4 *g_lo, *g_hi;        // [B,NGcap] AABB of each group of 32 sorted points (absolute MoGe)
  float4 *s_lo, *s_hi;        // [B,NScap] AABB of each super-group of 32 groups
  int NGcap, NScap;
  int4 *face_sv;              // [B,FOHO_ACCEL_FACES] faces in Morton order of their rest centroid: (ia, ib, ic, face id)
  int *face_rank;             // [B,FOHO_ACCEL_FACES] face id -> position in that order
  int *seed_c2h;              // [B,P]  warm start: hand slot found for each sorted cloud point last time
  int *seed_h2c;              // [B,FOHO_ACCEL_HV] warm start: sorted cloud position found for each hand vertex
  size_t total;
};

FOHO_HD size_t foho_align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static inline int foho_cand_capacity(int D) {
  long long n = (long long)D * D * D;
  return (int)(n < (1ll << 18) ? n : (1ll << 18));
}

static inline void foho_ws_layout(FohoWorkspace &w, char *base, int B, int D, int Vh, int Fh, int /*P*/, int Vo) {
  size_t off = 0;
  auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += foho_align_up(bytes, 256); return p; };
  w.W = (D + 31) / 32;
  w.cap = foho_cand_capacity(D);
  w.frames = (FohoFrame *)take(sizeof(FohoFrame) * (size_t)B);
  w.hmc = (float *)take(sizeof(float) * 3 * (size_t)B * Vh);
  w.hg = (float *)take(sizeof(float) * 3 * (size_t)B * Vh);
  w.G_hm = (float *)take(sizeof(float) * 3 * (size_t)B * Vh);
  w.G_hg = (float *)take(sizeof(float) * 3 * (size_t)B * Vh);
  w.acc = (float *)take(sizeof(float) * ACC_NUM * (size_t)B);
  w.cnt = (int *)take(sizeof(int) * CNT_NUM * (size_t)B);
  w.knn = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * Vh);
  w.stream_part = (float *)take(sizeof(float) * FOHO_STREAM_PARTIALS * FOHO_MAX_STREAM_CTAS * (size_t)B);
  w.parity = (uint32_t *)take(sizeof(uint32_t) * (size_t)B * D * D * w.W);
  w.cand = (int *)take(sizeof(int) * (size_t)B * w.cap);
  const size_t vo = Vo > 0 ? (size_t)Vo : 1;
  w.ot = (float *)take(sizeof(float) * 3 * vo);
  w.g_ot = (float *)take(sizeof(float) * 3 * vo);
  w.knn_obj = (unsigned long long *)take(sizeof(unsigned long long) * (size_t)B * Vh);
  w.oinfo = (FohoObjInfo *)take(sizeof(FohoObjInfo) * (size_t)B);
  w.sph = (float4 *)take(sizeof(float4) * (size_t)B * Fh);
  w.E_hm = (float *)take(sizeof(float) * 3 * (size_t)B * Vh);
  w.E_hg = (float *)take(sizeof(float) * 3 * (size_t)B * Vh);
  w.pen_part = (float *)take(sizeof(float) * 2 * FOHO_VE_MAX_CTAS * (size_t)B);
  w.kpbuf = (float *)take(sizeof(float) * 64 * (size_t)B);
  w.fin_acc = (float *)take(sizeof(float) * FOHO_FIN_NRED * (size_t)B);
  w.cand_val = (float *)take(sizeof(float) * (size_t)B * w.cap);
  w.tri_idx = (int *)take(sizeof(int) * 8 * (size_t)B * Vh);
  w.tri_val = (float *)take(sizeof(float) * 8 * (size_t)B * Vh);
  w.total = off;
}

// kernel ids of the optional timeline trace (desc->trace: 2 x u64 per id, host-initialised to {~0, 0})
enum { TR_PREP = 0, TR_STREAM, TR_H2C, TR_C2H, TR_CHAMFER, TR_RASTER, TR_COMPACT, TR_VOXDIST, TR_FIN, TR_ASM, TR_KP, TR_NUM };

// ---- device reductions -------------------------------------------------------------
#if defined(__CUDACC__)
struct FohoTrace {
  unsigned long long *p;
  __device__ __forceinline__ static unsigned long long now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
  }
  __device__ __forceinline__ FohoTrace(unsigned long long *base, int id) : p(nullptr) {
    if (base && threadIdx.x == 0) { p = base + 2 * id; atomicMin(p, now()); }
  }
  __device__ __forceinline__ ~FohoTrace() { if (p) atomicMax(p + 1, now()); }
};
// Object-side part of the per-sample frame: everything that depends only on theta_o, T_h2m, c_o and
// the lattice (a5/a6 of SURVEY.md section 8a; pipelines.py:108-118,242-250).  Called by k_prep and,
// so that the dense stream does not have to wait for k_prep, by thread 0 of every stream CTA.
__device__ __forceinline__ void foho_object_frame(const float *__restrict__ th, const float *__restrict__ T,
                                                  const float *__restrict__ co, float bound, int D, FohoFrame &fr) {
  for (int a = 0; a < 3; ++a) fr.co[a] = co[a];
  fr.so = th[8]; fr.to[0] = th[9]; fr.to[1] = th[10]; fr.to[2] = th[11];
  quat_to_mat(th + 12, fr.Ro);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) fr.Ah[3 * r + c] = T[4 * r + c];
  fr.step = 2.0f * bound / (float)(D - 1);
  fr.s_h2m = sqrtf(fr.Ah[0] * fr.Ah[0] + fr.Ah[3] * fr.Ah[3] + fr.Ah[6] * fr.Ah[6]);
  foho_f3 nb = f3(-bound, -bound, -bound);
  foho_f3 u0 = mat3_mul(fr.Ah, nb);
  fr.u0[0] = u0.x + (T[3] - co[0]); fr.u0[1] = u0.y + (T[7] - co[1]); fr.u0[2] = u0.z + (T[11] - co[2]);
  float Ahs[9], RA[9];
  for (int k = 0; k < 9; ++k) Ahs[k] = fr.Ah[k] * fr.step;
  mat3_matmul(fr.Ro, Ahs, RA);
  for (int k = 0; k < 9; ++k) fr.A[k] = fr.so * RA[k];
  foho_f3 ru = mat3_mul(fr.Ro, f3(fr.u0[0], fr.u0[1], fr.u0[2]));
  fr.bc[0] = fr.so * ru.x + fr.to[0]; fr.bc[1] = fr.so * ru.y + fr.to[1]; fr.bc[2] = fr.so * ru.z + fr.to[2];
  mat3_inverse(fr.A, fr.Ainv);
  mat3_inverse(Ahs, fr.Ahs_inv);
  fr.kappa = fr.so * fr.s_h2m * fr.step;
  foho_f3 babs = f3(fr.bc[0] + co[0], fr.bc[1] + co[1], fr.bc[2] + co[2]);
  foho_f3 e = mat3_tmul(fr.A, babs);
  fr.e[0] = e.x; fr.e[1] = e.y; fr.e[2] = e.z;
  fr.f = dot3(babs, babs);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum of N values per thread; result valid in thread 0 (and returned to all
// threads of warp 0).  smem must hold N * 32 floats.
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float *smem) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) smem[i * 32 + wid] = v[i];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float x = lane < nw ? smem[i * 32 + lane] : 0.f;
      v[i] = warp_sum(x);
    }
  }
}
#endif

// ---- per-device launch state (function attributes are per device; one process may drive several)
enum { FA_STREAM_TMA = 0, FA_CHAMFER, FA_C2H, FA_C2H_WALK, FA_VOXDIST, FA_NUM };
struct FohoDeviceState {
  int sm_count;
  size_t smem_attr[FA_NUM];      // largest dynamic shared-memory size opted into so far, per kernel
  bool carveout[FA_NUM];         // max-shared carve-out preference already set
};
FohoDeviceState *foho_device_state();     // of the current device; nullptr on a CUDA error
// opt `func` into `smem` bytes of dynamic shared memory (if it is more than before) and, when asked, into
// the largest shared-memory carve-out, so that kernels sharing an SM do not wait for it to drain
int foho_func_attrs(const void *func, int id, size_t smem, bool max_carveout);

// kernels implemented in the other translation units
// shared_sm: the sparse kernels run beside the stream (it then leaves them shared memory)
int foho_launch_stream(const foho_guidance_desc *d, const FohoWorkspace &ws, int *grid_x_out, bool shared_sm, cudaStream_t st);
// explicit object-mesh terms (guidance_objmesh.cu): `pre` runs before k_finalize_verts (it adds the contact
// gradient to G_hm), `post` after it (it adds to grad_theta[8..15] and the terms).
int foho_launch_objmesh_pre(const foho_guidance_desc *d, const FohoWorkspace &ws, cudaStream_t st);
int foho_launch_objmesh_post(const foho_guidance_desc *d, const FohoWorkspace &ws, cudaStream_t st);
// structured chamfer search (guidance_chamfer.cu); used when desc->accel is set
// grouped exact point->mesh distance of the candidate voxels (guidance_voxdist.cu); needs desc->accel
int foho_launch_voxdist_staged(const foho_guidance_desc *d, const FohoWorkspace &ws, cudaStream_t st);
void foho_accel_layout(FohoAccel &a, char *base, int B, int P);
int foho_sort_u64(unsigned long long *keys, int P2, int batch, cudaStream_t st);   // guidance_chamfer.cu
int foho_launch_chamfer_h2c(const foho_guidance_desc *d, const FohoWorkspace &ws, cudaStream_t st);
int foho_launch_chamfer_c2h(const foho_guidance_desc *d, const FohoWorkspace &ws, cudaStream_t st);
