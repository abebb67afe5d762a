// Explicit object-mesh terms of the guidance evaluation (REF rows a5/a6/a7/a10 of SURVEY.md 8a):
// the losses the reference evaluates on the FlexiCubes surface every inner iteration
// (third_party_patches/hy3dgen/shapegen/pipelines.py:1509-1541, 1561-1576):
//
//   moge_obj   = verts @ T_h2m[:3,:3]^T + T_h2m[:3,3]                     (:242-250, call :1520)
//   ot         = (s_o (moge_obj - c)) R_o^T + c + t_o,  c = bbox centre of moge_obj   (:108-118, :1523-1526)
//   distance   = mean_i clamp(min_j |h_i - ot_j|^2 - 0.01, 0)             (:1529-1541, pytorch3d knn_points K=1)
//   verts reg  = mean(ot^2)                                               (:1570)
//   edge loss  = mean_e |ot_e0 - ot_e1|^2                                 (:1575, pytorch3d mesh_edge_loss)
//   w_int      = 1e-5 if mean_i d2_i < 1e-3 and the step is late, else 1e-9   (:1561-1564)
//
// and their gradients to the hand leaves (through G_hm), the object leaves, and the object
// vertices themselves (grad_obj_verts: what FlexiCubes / the decoder would back-propagate).
// The bbox centre moves with the vertices; as in torch autograd its gradient is routed to the
// arg-min / arg-max vertex of every axis.
//
//   k_obj_prep   one CTA per sample: T_h2m, bbox centre (+arg indices), similarity, zeroing
//   k_obj_knn    brute-force 1-NN hand -> object vertices, object tiles in shared memory
//   k_obj_terms  contact / vertex / edge terms and dE/d(ot)
//   k_obj_chain  (after k_assemble) dE/d(ot) -> object leaves + grad_obj_verts, loss assembly
#include "foho_common.cuh"

namespace {

constexpr int OBJ_PREP_THREADS = 1024;
constexpr int OBJ_KNN_THREADS = 256;
constexpr int OBJ_KNN_TILE = 512;
constexpr int OBJ_HV = 4;   // hand vertices per thread per pass

__device__ __forceinline__ unsigned int orderable(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
  unsigned int lo = (unsigned int)v, hi = (unsigned int)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ foho_f3 obj_to_moge(const float *Ah, const float *T, const float *v) {
  foho_f3 m = mat3_mul(Ah, f3(v[0], v[1], v[2]));
  return f3(m.x + T[3], m.y + T[7], m.z + T[11]);
}

// ----------------------------------------------------------------------------- k_obj_prep
__global__ void __launch_bounds__(OBJ_PREP_THREADS) k_obj_prep(foho_guidance_desc d, FohoWorkspace ws) {
  __shared__ unsigned long long red[6][32];
  __shared__ FohoObjInfo info;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int v0 = d.obj_vert_offsets[b], v1 = d.obj_vert_offsets[b + 1];
  const FohoFrame &fr = ws.frames[b];
  const float *T = d.T_h2m + (size_t)b * 16;
  // keys: min -> (ord(f), j) minimised; max -> (ord(f), ~j) maximised: first index wins ties
  unsigned long long kmin[3] = {~0ull, ~0ull, ~0ull}, kmax[3] = {0ull, 0ull, 0ull};
  for (int j = v0 + tid; j < v1; j += blockDim.x) {
    const foho_f3 m = obj_to_moge(fr.Ah, T, d.obj_verts + 3 * (size_t)j);
    const float c[3] = {m.x, m.y, m.z};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const unsigned long long o = (unsigned long long)orderable(c[a]) << 32;
      const unsigned long long lo = o | (unsigned int)j, hi = o | (0xFFFFFFFFu - (unsigned int)j);
      kmin[a] = lo < kmin[a] ? lo : kmin[a];
      kmax[a] = hi > kmax[a] ? hi : kmax[a];
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long x = shfl_xor_u64(kmin[a], o), y = shfl_xor_u64(kmax[a], o);
      kmin[a] = x < kmin[a] ? x : kmin[a];
      kmax[a] = y > kmax[a] ? y : kmax[a];
    }
  if (lane == 0)
#pragma unroll
    for (int a = 0; a < 3; ++a) { red[a][wid] = kmin[a]; red[3 + a][wid] = kmax[a]; }
  __syncthreads();
  if (tid == 0) {
    const int nw = blockDim.x >> 5;
    for (int a = 0; a < 3; ++a) {
      unsigned long long lo = red[a][0], hi = red[3 + a][0];
      for (int w = 1; w < nw; ++w) { lo = red[a][w] < lo ? red[a][w] : lo; hi = red[3 + a][w] > hi ? red[3 + a][w] : hi; }
      if (v1 > v0) {
        info.c[a] = (from_orderable((unsigned int)(lo >> 32)) + from_orderable((unsigned int)(hi >> 32))) / 2.0f;
        info.amin[a] = (int)(unsigned int)lo;
        info.amax[a] = (int)(0xFFFFFFFFu - (unsigned int)hi);
      } else {
        info.c[a] = 0.f; info.amin[a] = -1; info.amax[a] = -1;
      }
    }
    info.v0 = v0; info.v1 = v1;
    info.e0 = d.obj_edge_offsets ? d.obj_edge_offsets[b] : 0;
    info.e1 = d.obj_edge_offsets ? d.obj_edge_offsets[b + 1] : 0;
    info.pad = 0;
    ws.oinfo[b] = info;
  }
  __syncthreads();
  // ot - c_o = s_o R_o (om - c) + ((c - c_o) + t_o)
  const foho_f3 c = f3(info.c[0], info.c[1], info.c[2]);
  const foho_f3 off = f3((c.x - fr.co[0]) + fr.to[0], (c.y - fr.co[1]) + fr.to[1], (c.z - fr.co[2]) + fr.to[2]);
  for (int j = v0 + tid; j < v1; j += blockDim.x) {
    const foho_f3 m = obj_to_moge(fr.Ah, T, d.obj_verts + 3 * (size_t)j);
    const foho_f3 r = mat3_mul(fr.Ro, fr.so * (m - c));
    float *o = ws.ot + 3 * (size_t)j, *g = ws.g_ot + 3 * (size_t)j;
    o[0] = r.x + off.x; o[1] = r.y + off.y; o[2] = r.z + off.z;
    g[0] = 0.f; g[1] = 0.f; g[2] = 0.f;
    if (d.obj_moge) {                        // absolute MoGe coordinates for the renderer
      float *om = d.obj_moge + 3 * (size_t)j;
      om[0] = o[0] + fr.co[0]; om[1] = o[1] + fr.co[1]; om[2] = o[2] + fr.co[2];
    }
  }
  for (int i = tid; i < d.Vh; i += blockDim.x) ws.knn_obj[(size_t)b * d.Vh + i] = ~0ull;
}

// ----------------------------------------------------------------------------- k_obj_knn
__global__ void __launch_bounds__(OBJ_KNN_THREADS) k_obj_knn(foho_guidance_desc d, FohoWorkspace ws) {
  __shared__ float4 so[OBJ_KNN_TILE];
  const int b = blockIdx.y, Vh = d.Vh;
  const FohoObjInfo &info = ws.oinfo[b];
  const int base = info.v0 + blockIdx.x * OBJ_KNN_TILE;
  if (base >= info.v1) return;
  const int m = min(OBJ_KNN_TILE, info.v1 - base);
  for (int k = threadIdx.x; k < m; k += blockDim.x) {
    const float *o = ws.ot + 3 * (size_t)(base + k);
    so[k] = make_float4(o[0], o[1], o[2], 0.f);
  }
  __syncthreads();
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;
  unsigned long long *knn = ws.knn_obj + (size_t)b * Vh;
  for (int i0 = 0; i0 < Vh; i0 += OBJ_KNN_THREADS * OBJ_HV) {
    float hx[OBJ_HV], hy[OBJ_HV], hz[OBJ_HV], best[OBJ_HV];
    int bj[OBJ_HV];
#pragma unroll
    for (int u = 0; u < OBJ_HV; ++u) {
      const int i = i0 + threadIdx.x + u * OBJ_KNN_THREADS;
      const bool ok = i < Vh;
      hx[u] = ok ? hmc[3 * i] : 0.f; hy[u] = ok ? hmc[3 * i + 1] : 0.f; hz[u] = ok ? hmc[3 * i + 2] : 0.f;
      best[u] = INFINITY; bj[u] = 0;
    }
    for (int j = 0; j < m; ++j) {
      const float4 o = so[j];
#pragma unroll
      for (int u = 0; u < OBJ_HV; ++u) {
        const float dx = hx[u] - o.x, dy = hy[u] - o.y, dz = hz[u] - o.z;
        const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d2 < best[u]) { best[u] = d2; bj[u] = j; }
      }
    }
#pragma unroll
    for (int u = 0; u < OBJ_HV; ++u) {
      const int i = i0 + threadIdx.x + u * OBJ_KNN_THREADS;
      if (i < Vh && best[u] < INFINITY)
        atomicMin(knn + i, ((unsigned long long)__float_as_uint(best[u]) << 32) | (unsigned int)(base + bj[u]));
    }
  }
}

// ----------------------------------------------------------------------------- k_obj_terms
__global__ void __launch_bounds__(256) k_obj_terms(foho_guidance_desc d, FohoWorkspace ws) {
  __shared__ float red[4 * 32];
  const int b = blockIdx.y, Vh = d.Vh;
  const FohoObjInfo info = ws.oinfo[b];
  const int n = info.v1 - info.v0, ne = info.e1 - info.e0;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};          // dist, mean_d2, vreg, edge
  if (n > 0) {
    const FohoFrame &fr = ws.frames[b];
    const float *hmc = ws.hmc + (size_t)b * Vh * 3;
    float *Ghm = ws.G_hm + (size_t)b * Vh * 3;
    const foho_weights &W = d.w;
    const int items = Vh + n + ne;
    const float cd = 2.f * W.w_dist / (float)Vh, cv = 2.f * W.w_vreg / (3.f * (float)n);
    const float ce = ne > 0 ? 2.f * W.w_edge / (float)ne : 0.f;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < items; t += gridDim.x * blockDim.x) {
      if (t < Vh) {
        const unsigned long long key = ws.knn_obj[(size_t)b * Vh + t];
        if (key == ~0ull) continue;
        const int j = (int)(unsigned int)key;
        const float *o = ws.ot + 3 * (size_t)j;
        const foho_f3 df = f3(hmc[3 * t] - o[0], hmc[3 * t + 1] - o[1], hmc[3 * t + 2] - o[2]);
        const float d2 = fmaf(df.z, df.z, fmaf(df.y, df.y, df.x * df.x));
        acc[1] += d2;
        const float a = d2 - W.dist_margin;
        if (a >= 0.f) {
          acc[0] += a;
          const foho_f3 g = cd * df;
          Ghm[3 * t] += g.x; Ghm[3 * t + 1] += g.y; Ghm[3 * t + 2] += g.z;       // one thread per hand vertex
          float *go = ws.g_ot + 3 * (size_t)j;
          atomicAdd(go, -g.x); atomicAdd(go + 1, -g.y); atomicAdd(go + 2, -g.z);
        }
      } else if (t < Vh + n) {
        const int j = info.v0 + (t - Vh);
        const float *o = ws.ot + 3 * (size_t)j;
        const foho_f3 p = f3(o[0] + fr.co[0], o[1] + fr.co[1], o[2] + fr.co[2]);
        acc[2] += dot3(p, p);
        float *go = ws.g_ot + 3 * (size_t)j;
        atomicAdd(go, cv * p.x); atomicAdd(go + 1, cv * p.y); atomicAdd(go + 2, cv * p.z);
      } else {
        const int e = info.e0 + (t - Vh - n);
        const int ia = d.obj_edges[2 * (size_t)e], ib = d.obj_edges[2 * (size_t)e + 1];
        const float *pa = ws.ot + 3 * (size_t)ia, *pb = ws.ot + 3 * (size_t)ib;
        const foho_f3 df = f3(pa[0] - pb[0], pa[1] - pb[1], pa[2] - pb[2]);
        acc[3] += dot3(df, df);
        const foho_f3 g = ce * df;
        float *ga = ws.g_ot + 3 * (size_t)ia, *gb = ws.g_ot + 3 * (size_t)ib;
        atomicAdd(ga, g.x); atomicAdd(ga + 1, g.y); atomicAdd(ga + 2, g.z);
        atomicAdd(gb, -g.x); atomicAdd(gb + 1, -g.y); atomicAdd(gb + 2, -g.z);
      }
    }
  }
  block_sum<4>(acc, red);
  if (threadIdx.x == 0) {
    float *ac = ws.acc + (size_t)b * ACC_NUM;
    if (acc[0] != 0.f) atomicAdd(ac + ACC_DIST, acc[0]);
    if (acc[1] != 0.f) atomicAdd(ac + ACC_MEAN_D2, acc[1]);
    if (acc[2] != 0.f) atomicAdd(ac + ACC_VREG, acc[2]);
    if (acc[3] != 0.f) atomicAdd(ac + ACC_EDGE, acc[3]);
  }
}

// ----------------------------------------------------------------------------- k_obj_chain
__global__ void __launch_bounds__(OBJ_PREP_THREADS) k_obj_chain(foho_guidance_desc d, FohoWorkspace ws) {
  __shared__ float red[12 * 32];
  __shared__ float gc_s[3];
  const int b = blockIdx.x, tid = threadIdx.x;
  const FohoObjInfo info = ws.oinfo[b];
  const int n = info.v1 - info.v0, ne = info.e1 - info.e0;
  const FohoFrame &fr = ws.frames[b];
  const float *T = d.T_h2m + (size_t)b * 16;
  const foho_f3 c = f3(info.c[0], info.c[1], info.c[2]);
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  for (int j = info.v0 + tid; j < info.v1; j += blockDim.x) {
    float *g = ws.g_ot + 3 * (size_t)j;
    if (d.grad_obj_ext) {                    // renderer terms join here: dE/d(ot) += dE_ext/d(ot)
      const float *ge = d.grad_obj_ext + 3 * (size_t)j;
      g[0] += ge[0]; g[1] += ge[1]; g[2] += ge[2];
    }
    const foho_f3 w = obj_to_moge(fr.Ah, T, d.obj_verts + 3 * (size_t)j) - c;
    acc[0] += g[0]; acc[1] += g[1]; acc[2] += g[2];
    acc[3] += g[0] * w.x; acc[4] += g[0] * w.y; acc[5] += g[0] * w.z;
    acc[6] += g[1] * w.x; acc[7] += g[1] * w.y; acc[8] += g[1] * w.z;
    acc[9] += g[2] * w.x; acc[10] += g[2] * w.y; acc[11] += g[2] * w.z;
  }
  block_sum<12>(acc, red);
  if (tid == 0) {
    const foho_weights &W = d.w;
    float *go = d.grad_theta + (size_t)b * 16;
    float *Tm = d.terms + (size_t)b * FOHO_NUM_TERMS;
    const float *ac = ws.acc + (size_t)b * ACC_NUM;
    float dso = 0.f, GRo[9];
    for (int k = 0; k < 9; ++k) { dso += fr.Ro[k] * acc[3 + k]; GRo[k] = fr.so * acc[3 + k]; }
    const foho_f3 Sg = f3(acc[0], acc[1], acc[2]);
    const foho_f3 rs = fr.so * mat3_tmul(fr.Ro, Sg);
    gc_s[0] = Sg.x - rs.x; gc_s[1] = Sg.y - rs.y; gc_s[2] = Sg.z - rs.z;     // dE/dc = (I - s R)^T sum g
    float gq[4];
    quat_to_mat_backward(d.theta + (size_t)b * 16 + 12, GRo, gq);
    if (n > 0) {
      go[8] += dso; go[9] += Sg.x; go[10] += Sg.y; go[11] += Sg.z;
      go[12] += gq[0]; go[13] += gq[1]; go[14] += gq[2]; go[15] += gq[3];
    }
    const float invV = 1.f / (float)d.Vh;
    const float L_dist = n > 0 ? ac[ACC_DIST] * invV : 0.f;
    const float mean_d2 = n > 0 ? ac[ACC_MEAN_D2] * invV : 0.f;
    const float L_vreg = n > 0 ? ac[ACC_VREG] / (3.f * (float)n) : 0.f;
    const float L_edge = ne > 0 ? ac[ACC_EDGE] / (float)ne : 0.f;
    Tm[FOHO_T_DIST] = L_dist; Tm[FOHO_T_VREG] = L_vreg; Tm[FOHO_T_EDGE] = L_edge; Tm[FOHO_T_MEAN_D2] = mean_d2;
    float total = Tm[FOHO_T_TOTAL] + W.w_dist * L_dist + W.w_vreg * L_vreg + W.w_edge * L_edge;
    // pipelines.py:1561-1564 (k_assemble used w_int_lo)
    if (n > 0 && mean_d2 < 0.001f && d.late_step) total += (W.w_int_hi - W.w_int_lo) * Tm[FOHO_T_COUNT];
    Tm[FOHO_T_TOTAL] = total;
  }
  __syncthreads();
  if (!d.grad_obj_verts) return;
  const foho_f3 gc = f3(0.5f * gc_s[0], 0.5f * gc_s[1], 0.5f * gc_s[2]);
  for (int j = info.v0 + tid; j < info.v1; j += blockDim.x) {
    const float *g = ws.g_ot + 3 * (size_t)j;
    foho_f3 gm = fr.so * mat3_tmul(fr.Ro, f3(g[0], g[1], g[2]));
    if (j == info.amin[0]) gm.x += gc.x;
    if (j == info.amax[0]) gm.x += gc.x;
    if (j == info.amin[1]) gm.y += gc.y;
    if (j == info.amax[1]) gm.y += gc.y;
    if (j == info.amin[2]) gm.z += gc.z;
    if (j == info.amax[2]) gm.z += gc.z;
    const foho_f3 gv = mat3_tmul(fr.Ah, gm);
    float *o = d.grad_obj_verts + 3 * (size_t)j;
    o[0] = gv.x; o[1] = gv.y; o[2] = gv.z;
  }
}

}  // namespace

int foho_launch_objmesh_pre(const foho_guidance_desc *dp, const FohoWorkspace &ws, cudaStream_t st) {
  const foho_guidance_desc &d = *dp;
  k_obj_prep<<<d.B, OBJ_PREP_THREADS, 0, st>>>(d, ws);
  FOHO_LAUNCH_CHECK();
  const int nchunk = (d.Vo_total + OBJ_KNN_TILE - 1) / OBJ_KNN_TILE;     // upper bound per sample
  k_obj_knn<<<dim3(nchunk, d.B), OBJ_KNN_THREADS, 0, st>>>(d, ws);
  FOHO_LAUNCH_CHECK();
  long long items = (long long)d.Vh + d.Vo_total + d.Eo_total;
  int nblk = (int)((items + 1023) / 1024);
  if (nblk > 128) nblk = 128;
  if (nblk < 1) nblk = 1;
  k_obj_terms<<<dim3(nblk, d.B), 256, 0, st>>>(d, ws);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}

int foho_launch_objmesh_post(const foho_guidance_desc *dp, const FohoWorkspace &ws, cudaStream_t st) {
  k_obj_chain<<<dp->B, OBJ_PREP_THREADS, 0, st>>>(*dp, ws);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
