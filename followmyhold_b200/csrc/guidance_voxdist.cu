// Exact point->mesh distance of the candidate voxels (NS row a14: SDF_h of the voxels inside both the
// hand and the object; a8 semantics of third_party/utilz/kaolin_sdf_ops.py:88-109 evaluated on the
// Hunyuan lattice), staged search.
//
// k_raster writes each face's bounding sphere (lattice units, current pose).  A CTA stages the spheres
// and the posed vertices in shared memory once and its warps then serve runs of consecutive candidates
// (k_compact emits them column by column, so a run is spatially coherent):
//   bound   : the first candidate of a run takes the nearest vertex (25 smem reads per lane); every
//             later one the previous answer plus the step between the two voxels (the distance
//             function is 1-Lipschitz), so the bound is within a voxel of the answer;
//   filter  : all lanes sweep the face spheres against the bound and append the survivors (~2 % of the
//             faces: the candidates sit 1-3 voxels under the skin) to a per-warp list;
//   exact   : whenever 32 survivors are queued, one Ericson closest-point evaluation per lane, fully
//             converged, and the bound tightens for the rest of the sweep.
#include "foho_common.cuh"

namespace {

constexpr int VT_THREADS = 256;
constexpr int VT_WARPS = VT_THREADS / 32;
constexpr int VT_QUEUE = 160;            // per-warp survivor queue (>= 31 + 128)

struct VtBest { float d2, wa, wb, wc; int s; };

__global__ void __launch_bounds__(VT_THREADS) k_voxdist_staged(foho_guidance_desc d, FohoWorkspace ws, FohoAccel acc) {
  FohoTrace trace_(ws.trace, TR_VOXDIST);
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.y, D = d.D, Vh = d.Vh, Fh = d.Fh;
  const int *cnt = ws.cnt + (size_t)b * CNT_NUM;
  int n = cnt[CNT_NCAND];
  if (n > ws.cap) n = ws.cap;
  // consecutive runs of candidates per warp
  const int total_warps = gridDim.x * VT_WARPS;
  const int run = (n + total_warps - 1) / total_warps;
  if (run == 0 || (int)(blockIdx.x * VT_WARPS) * run >= n) return;
  const int NF = (Fh + 31) & ~31;
  float4 *sph = reinterpret_cast<float4 *>(sm_raw);             // [NF] face spheres (pad: r = -1)
  float4 *sv = sph + NF;                                        // [Vh] posed vertices, lattice units
  int *queue = reinterpret_cast<int *>(sv + Vh);                // [VT_WARPS][VT_QUEUE]
  ushort4 *sfv = reinterpret_cast<ushort4 *>(queue + VT_WARPS * VT_QUEUE);   // [NF] vertex ids of the faces (Vh <= 65535)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  {
    const float4 *gs = ws.sph + (size_t)b * Fh;
#pragma unroll 8
    for (int i = threadIdx.x; i < NF; i += blockDim.x) sph[i] = i < Fh ? gs[i] : make_float4(0.f, 0.f, 0.f, -1.f);
    const float *hg = ws.hg + (size_t)b * Vh * 3;
#pragma unroll 4
    for (int i = threadIdx.x; i < Vh; i += blockDim.x) sv[i] = make_float4(hg[3 * i], hg[3 * i + 1], hg[3 * i + 2], 0.f);
    const int4 *gf = acc.face_sv + (size_t)b * FOHO_ACCEL_FACES;
#pragma unroll 8
    for (int i = threadIdx.x; i < Fh; i += blockDim.x) {
      const int4 f = gf[i];
      sfv[i] = make_ushort4((unsigned short)f.x, (unsigned short)f.y, (unsigned short)f.z, 0);
    }
  }
  __syncthreads();

  const FohoFrame &fr = ws.frames[b];
  const float kappa = fr.kappa;
  const float N = (float)D * (float)D * (float)D;
  const float *S = d.sdf + (size_t)b * D * D * D;
  float *Ghg = ws.G_hg + (size_t)b * Vh * 3;
  const int *cand = ws.cand + (size_t)b * ws.cap;
  float *cval = ws.cand_val + (size_t)b * ws.cap;   // dE/dS of the candidate, added to G by k_assemble
  int *q = queue + wid * VT_QUEUE;
  float acc_int = 0.f, acc_gk = 0.f;
  const int c0 = (blockIdx.x * VT_WARPS + wid) * run;
  const int c1 = min(n, c0 + run);
  foho_f3 pprev = f3(0.f, 0.f, 0.f);
  float dprev = -1.f;                                          // distance of the previous candidate of this run
  // the voxel index and the field value of up to 32 candidates of the run are fetched together (lane k
  // holds candidate c0+k): one round trip to DRAM per 32 candidates instead of two per candidate
  int vpre = 0;
  float spre = 0.f;
  for (int c = c0; c < c1; ++c) {
    const int k = (c - c0) & 31;
    if (k == 0) {
      const int cc = c + lane;
      vpre = cc < c1 ? cand[cc] : 0;
      spre = cc < c1 ? S[vpre] : 0.f;
    }
    const int v = __shfl_sync(0xffffffffu, vpre, k);
    const float sval = __shfl_sync(0xffffffffu, spre, k);
    const int Z = v % D, Y = (v / D) % D, X = v / (D * D);
    const foho_f3 p = f3((float)X, (float)Y, (float)Z);
    // ---- bound
    float wbest;                                               // warp-uniform squared bound on the answer
    {
      const foho_f3 dp = p - pprev;
      const float jump = sqrtf(dot3(dp, dp));
      if (dprev >= 0.f && jump <= 4.f) {
        const float u = (dprev + jump) * 1.00001f + 1e-6f;
        wbest = u * u;
      } else {
        float ub2 = INFINITY;
        for (int i = lane; i < Vh; i += 32) {
          const float4 w = sv[i];
          const float dx = w.x - p.x, dy = w.y - p.y, dz = w.z - p.z;
          ub2 = fminf(ub2, fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
        }
        wbest = warp_min(ub2) * 1.00001f + 1e-12f;
      }
    }
    VtBest bst; bst.d2 = INFINITY; bst.wa = bst.wb = bst.wc = 0.f; bst.s = -1;
    int nq = 0;                                                // warp-uniform queue fill
    auto drain = [&](int count) {                              // exact test of queue[0..count), one per lane
      if (lane < count) {
        const int s = q[lane];
        const ushort4 f = sfv[s];
        float wa, wb, wc;
        // translate by -p first: differences of nearby lattice coordinates are (nearly) exact in fp32
        const float4 A = sv[f.x], Bv = sv[f.y], C = sv[f.z];
        const float d2 = closest_point_triangle(f3(0.f, 0.f, 0.f), f3(A.x, A.y, A.z) - p, f3(Bv.x, Bv.y, Bv.z) - p,
                                                f3(C.x, C.y, C.z) - p, wa, wb, wc);
        if (d2 < bst.d2) { bst.d2 = d2; bst.wa = wa; bst.wb = wb; bst.wc = wc; bst.s = s; }
      }
      wbest = fminf(wbest, warp_min(bst.d2));
    };
    // ---- filter (+ exact whenever a full warp of survivors is queued); 4 independent sphere tests per
    //      lane and step keep the loads and the arithmetic of a step in flight together
    float wroot = sqrtf(wbest);
    for (int s0 = 0; s0 < NF; s0 += 128) {
      bool keep[4];
      unsigned m[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int s = s0 + u * 32 + lane;
        const float4 s4 = s < NF ? sph[s] : make_float4(0.f, 0.f, 0.f, -1.f);
        const float dx = s4.x - p.x, dy = s4.y - p.y, dz = s4.z - p.z;
        const float dc2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const float t = (s4.w + wroot) * 1.00001f;            // |centre - p| > r + bound  =>  face cannot win
        keep[u] = s4.w >= 0.f && !(dc2 > t * t);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) m[u] = __ballot_sync(0xffffffffu, keep[u]);
      if ((m[0] | m[1] | m[2] | m[3]) == 0u) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (keep[u]) q[nq + __popc(m[u] & ((1u << lane) - 1u))] = s0 + u * 32 + lane;
        nq += __popc(m[u]);
      }
      __syncwarp();
      while (nq >= 32) {
        // oldest 32 entries sit at the tail end: take the last 32 so nothing has to move
        nq -= 32;
        const int s = q[nq + lane];
        {
          const ushort4 f = sfv[s];
          float wa, wb, wc;
          const float4 A = sv[f.x], Bv = sv[f.y], C = sv[f.z];
          const float d2 = closest_point_triangle(f3(0.f, 0.f, 0.f), f3(A.x, A.y, A.z) - p, f3(Bv.x, Bv.y, Bv.z) - p,
                                                  f3(C.x, C.y, C.z) - p, wa, wb, wc);
          if (d2 < bst.d2) { bst.d2 = d2; bst.wa = wa; bst.wb = wb; bst.wc = wc; bst.s = s; }
        }
        const float nb = fminf(wbest, warp_min(bst.d2));
        if (nb < wbest) { wbest = nb; wroot = sqrtf(wbest); }
        __syncwarp();
      }
    }
    if (nq > 0) drain(nq);
    __syncwarp();
    // ---- the lane holding the warp's best finishes the candidate (ties -> lowest lane)
    const float mb = warp_min(bst.d2);
    const unsigned vote = __ballot_sync(0xffffffffu, bst.s >= 0 && bst.d2 == mb);
    if (vote == 0u) {                                       // cannot happen for a non-empty mesh
      if (lane == 0) cval[c] = 0.f;
      dprev = -1.f;
      continue;
    }
    dprev = sqrtf(mb); pprev = p;
    if (lane == __ffs(vote) - 1) {
      const float s = sval;                                  // S[v] < 0 by construction
      const float dist = sqrtf(bst.d2);
      const float ns = -s;
      acc_int += ns * dist;
      const float coef = d.w.w_ivol * ns / N;               // dE/d(kappa*dist)
      acc_gk += coef * dist;
      cval[c] = -d.w.w_ivol * kappa * dist / N;
      if (dist > 0.f) {
        const ushort4 f = sfv[bst.s];
        const float4 A = sv[f.x], Bv = sv[f.y], C = sv[f.z];
        const foho_f3 a = f3(A.x, A.y, A.z) - p, bb = f3(Bv.x, Bv.y, Bv.z) - p, cc = f3(C.x, C.y, C.z) - p;
        const foho_f3 qq = f3(bst.wa * a.x + bst.wb * bb.x + bst.wc * cc.x, bst.wa * a.y + bst.wb * bb.y + bst.wc * cc.y,
                              bst.wa * a.z + bst.wb * bb.z + bst.wc * cc.z);       // closest point relative to p
        const float inv = 1.f / dist;
        const foho_f3 dir = (-inv) * qq;
        const float k = -coef * kappa;                      // d(dist)/dv_k = -w_k dir
        atomicAdd(Ghg + 3 * f.x, k * bst.wa * dir.x); atomicAdd(Ghg + 3 * f.x + 1, k * bst.wa * dir.y); atomicAdd(Ghg + 3 * f.x + 2, k * bst.wa * dir.z);
        atomicAdd(Ghg + 3 * f.y, k * bst.wb * dir.x); atomicAdd(Ghg + 3 * f.y + 1, k * bst.wb * dir.y); atomicAdd(Ghg + 3 * f.y + 2, k * bst.wb * dir.z);
        atomicAdd(Ghg + 3 * f.z, k * bst.wc * dir.x); atomicAdd(Ghg + 3 * f.z + 1, k * bst.wc * dir.y); atomicAdd(Ghg + 3 * f.z + 2, k * bst.wc * dir.z);
      }
    }
  }
  acc_int = warp_sum(acc_int);
  acc_gk = warp_sum(acc_gk);
  if (lane == 0 && (acc_int != 0.f || acc_gk != 0.f)) {
    atomicAdd(ws.acc + (size_t)b * ACC_NUM + ACC_INT, acc_int);
    atomicAdd(ws.acc + (size_t)b * ACC_NUM + ACC_GKAPPA, acc_gk);
  }
}

}  // namespace

int foho_launch_voxdist_staged(const foho_guidance_desc *dp, const FohoWorkspace &ws, cudaStream_t st) {
  const foho_guidance_desc &d = *dp;
  FohoAccel a;
  foho_accel_layout(a, (char *)d.accel, d.B, d.P);
  if (a.total > d.accel_bytes) return FOHO_E_WORKSPACE;
  const int NF = (d.Fh + 31) & ~31;
  const size_t smem = ((size_t)NF + (size_t)d.Vh) * sizeof(float4) + VT_WARPS * VT_QUEUE * sizeof(int) + (size_t)NF * sizeof(ushort4);
  if (smem > 200 * 1024) return FOHO_E_SHAPE;
  {
    int rc = foho_func_attrs((const void *)k_voxdist_staged, FA_VOXDIST, smem, true);
    if (rc != FOHO_OK) return rc;
  }
  k_voxdist_staged<<<dim3(64, d.B), VT_THREADS, smem, st>>>(d, ws, a);
  FOHO_LAUNCH_CHECK();
  return FOHO_OK;
}
