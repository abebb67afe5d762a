// Sparse part of the guidance evaluation: everything that is not the dense volume stream.
//
//   k_prep            a5/a6  leaves -> frames, transformed hand verts (MoGe + lattice), bbox, zeroing
//   k_raster          a8     +z ray-parity voxelisation of the hand on the lattice (sign rule); face spheres
//   k_compact         a9     voxels inside hand & object -> candidate list (= the REF count)
//   k_voxdist         a14    exact point->mesh distance for the candidates (flat sweep; the staged search of
//                            guidance_voxdist.cu is used when the per-image face order exists)
//   k_chamfer         a15    brute-force 1-NN both ways between hand verts and the MoGe cloud (no accel buffer;
//                            the structured searches live in guidance_chamfer.cu)
//   k_keypoints       a11    key-point regression, projection, MSE and its gradient
//   k_vertex_early    a13    trilinear vertex samples, penalties, corner lists, key-point back-projection
//   k_finalize_verts  a12    chain rule from the per-vertex gradients to 26 moments of the 16 leaves
//   k_assemble        a10/a12  deferred dE/dS scatter; stream moments + vertex sums -> leaf gradients, loss terms
//
// Reference seams: third_party_patches/hy3dgen/shapegen/pipelines.py:108-118 (a6),
// :121-135 (a11), :231-239 (a9), :242-250 (a5), :1480-1600 (inner iteration);
// third_party/utilz/kaolin_sdf_ops.py:88-109 (a8).  Term definitions: DESIGN.md.
#include "foho_common.cuh"
#include <mutex>

namespace {

// ----------------------------------------------------------------------------- k_prep
__global__ void __launch_bounds__(256) k_prep(foho_guidance_desc d, FohoWorkspace ws) {
  FohoTrace trace_(ws.trace, TR_PREP);
  __shared__ FohoFrame fr;
  __shared__ float red[6 * 32];
  const int b = blockIdx.x, tid = threadIdx.x, Vh = d.Vh, D = d.D;
  const float *rest = d.hand_rest + (size_t)b * Vh * 3;

  // bbox centre of the rest hand (pipelines.py:111)
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = tid; i < Vh; i += blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float v = rest[3 * i + a];
      mn[a] = fminf(mn[a], v);
      mx[a] = fmaxf(mx[a], v);
    }
  }
  const int lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = warp_min(mn[a]); mx[a] = warp_max(mx[a]); }
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { red[a * 32 + wid] = mn[a]; red[(3 + a) * 32 + wid] = mx[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    const float *th = d.theta + (size_t)b * 16;
    const float *T = d.T_h2m + (size_t)b * 16;
    const float *co = d.obj_center + (size_t)b * 3;
    for (int a = 0; a < 3; ++a) {
      float lo = red[a * 32], hi = red[(3 + a) * 32];
      for (int w = 1; w < nw; ++w) { lo = fminf(lo, red[a * 32 + w]); hi = fmaxf(hi, red[(3 + a) * 32 + w]); }
      fr.ch[a] = (lo + hi) / 2.0f;
      fr.chc[a] = fr.ch[a] - co[a];
    }
    fr.sh = th[0]; fr.th[0] = th[1]; fr.th[1] = th[2]; fr.th[2] = th[3];
    quat_to_mat(th + 4, fr.Rh);
    foho_object_frame(th, T, co, d.bound, D, fr);
  }
  __syncthreads();

  // transformed hand verts (centred MoGe) and lattice coordinates; lattice bbox
  float *hmc = ws.hmc + (size_t)b * Vh * 3, *hg = ws.hg + (size_t)b * Vh * 3;
  float *Ghm = ws.G_hm + (size_t)b * Vh * 3, *Ghg = ws.G_hg + (size_t)b * Vh * 3;
  for (int a = 0; a < 3; ++a) { mn[a] = INFINITY; mx[a] = -INFINITY; }
  for (int i = tid; i < Vh; i += blockDim.x) {
    foho_f3 w = f3(rest[3 * i] - fr.ch[0], rest[3 * i + 1] - fr.ch[1], rest[3 * i + 2] - fr.ch[2]);
    foho_f3 rw = mat3_mul(fr.Rh, fr.sh * w);                      // (s (v-c)) R^T  (row-vector form of :116)
    foho_f3 m = f3(rw.x + (fr.chc[0] + fr.th[0]), rw.y + (fr.chc[1] + fr.th[1]), rw.z + (fr.chc[2] + fr.th[2]));
    hmc[3 * i] = m.x; hmc[3 * i + 1] = m.y; hmc[3 * i + 2] = m.z;
    foho_f3 g = mat3_mul(fr.Ainv, f3(m.x - fr.bc[0], m.y - fr.bc[1], m.z - fr.bc[2]));
    hg[3 * i] = g.x; hg[3 * i + 1] = g.y; hg[3 * i + 2] = g.z;
    mn[0] = fminf(mn[0], g.x); mn[1] = fminf(mn[1], g.y); mn[2] = fminf(mn[2], g.z);
    mx[0] = fmaxf(mx[0], g.x); mx[1] = fmaxf(mx[1], g.y); mx[2] = fmaxf(mx[2], g.z);
    if (d.hand_moge) {
      float *o = d.hand_moge + ((size_t)b * Vh + i) * 3;
      o[0] = m.x + fr.co[0]; o[1] = m.y + fr.co[1]; o[2] = m.z + fr.co[2];
    }
    if (d.hand_grid) {
      float *o = d.hand_grid + ((size_t)b * Vh + i) * 3;
      o[0] = g.x; o[1] = g.y; o[2] = g.z;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) { Ghm[3 * i + a] = 0.f; Ghg[3 * i + a] = 0.f; }
    ws.knn[(size_t)b * Vh + i] = 0xFFFFFFFFFFFFFFFFull;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = warp_min(mn[a]); mx[a] = warp_max(mx[a]); }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { red[a * 32 + wid] = mn[a]; red[(3 + a) * 32 + wid] = mx[a]; }
  }
  __syncthreads();
  if (tid == 0) {
    for (int a = 0; a < 3; ++a) {
      float lo = red[a * 32], hi = red[(3 + a) * 32];
      for (int w = 1; w < nw; ++w) { lo = fminf(lo, red[a * 32 + w]); hi = fmaxf(hi, red[(3 + a) * 32 + w]); }
      float cl = ceilf(lo), fh = floorf(hi);
      int ilo = cl <= 0.f ? 0 : (cl >= (float)D ? D : (int)cl);
      int ihi = fh >= (float)(D - 1) ? D - 1 : (fh < 0.f ? -1 : (int)fh);
      if (!(lo == lo) || !(hi == hi)) { ilo = 1; ihi = 0; }     // NaN leaves -> empty
      fr.lo[a] = ilo; fr.hi[a] = ihi;
    }
    for (int k = 0; k < ACC_NUM; ++k) ws.acc[(size_t)b * ACC_NUM + k] = 0.f;
    for (int k = 0; k < CNT_NUM; ++k) ws.cnt[(size_t)b * CNT_NUM + k] = 0;
  }
  __syncthreads();
  // publish frame
  {
    const int nwords = sizeof(FohoFrame) / 4;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&fr);
    uint32_t *dst = reinterpret_cast<uint32_t *>(ws.frames + b);
    for (int k = tid; k < nwords; k += blockDim.x) dst[k] = src[k];
  }
  // clear the parity columns of the bbox
  const int nx = fr.hi[0] - fr.lo[0] + 1, ny = fr.hi[1] - fr.lo[1] + 1;
  if (nx > 0 && ny > 0 && fr.hi[2] >= fr.lo[2]) {
    uint32_t *par = ws.parity + (size_t)b * D * D * ws.W;
    const int items = nx * ny * ws.W;
    for (int k = tid; k < items; k += blockDim.x) {
      int w = k % ws.W, c = k / ws.W;
      int X = fr.lo[0] + c / ny, Y = fr.lo[1] + c % ny;
      par[((size_t)X * D + Y) * ws.W + w] = 0u;
    }
  }
}

// ----------------------------------------------------------------------------- k_raster
// eight lanes per hand face, striding over the columns of the face's xy bounding box: XOR the "below the
// crossing" prefix into every column the projection covers (rule: foho_math.cuh::column_hits_triangle).
// (A thread per face leaves the few large faces -- the wrist cap fan -- as a long serial tail.)
constexpr int RASTER_THREADS = 256;
constexpr int RASTER_LANES = 8;          // lanes per face: a typical face covers ~3x3 columns
__global__ void __launch_bounds__(RASTER_THREADS) k_raster(foho_guidance_desc d, FohoWorkspace ws, const int *__restrict__ face_rank) {
  FohoTrace trace_(ws.trace, TR_RASTER);
  const int b = blockIdx.y, D = d.D, lane = threadIdx.x & (RASTER_LANES - 1);
  const int f = (blockIdx.x * RASTER_THREADS + threadIdx.x) / RASTER_LANES;
  if (f >= d.Fh) return;
  const FohoFrame &fr = ws.frames[b];
  const float *hg = ws.hg + (size_t)b * d.Vh * 3;
  const int ia = d.hand_faces[3 * f], ib = d.hand_faces[3 * f + 1], ic = d.hand_faces[3 * f + 2];
  const foho_f3 a = f3(hg[3 * ia], hg[3 * ia + 1], hg[3 * ia + 2]);
  const foho_f3 bb = f3(hg[3 * ib], hg[3 * ib + 1], hg[3 * ib + 2]);
  const foho_f3 c = f3(hg[3 * ic], hg[3 * ic + 1], hg[3 * ic + 2]);
  if (lane == 0) {
    // bounding sphere of the face for k_voxdist's culling
    const foho_f3 m = (1.f / 3.f) * (a + bb + c);
    const foho_f3 da = a - m, db = bb - m, dc = c - m;
    const float r2 = fmaxf(dot3(da, da), fmaxf(dot3(db, db), dot3(dc, dc)));
    // with the per-image face order (accel) the sphere goes to the face's position in that order
    const int slot = face_rank ? face_rank[(size_t)b * FOHO_ACCEL_FACES + f] : f;
    ws.sph[(size_t)b * d.Fh + slot] = make_float4(m.x, m.y, m.z, sqrtf(r2) * 1.00001f + 1e-6f);
  }
  if (fr.hi[0] < fr.lo[0] || fr.hi[1] < fr.lo[1] || fr.hi[2] < fr.lo[2]) return;
  float fxmin = ceilf(fminf(a.x, fminf(bb.x, c.x))), fxmax = floorf(fmaxf(a.x, fmaxf(bb.x, c.x)));
  float fymin = ceilf(fminf(a.y, fminf(bb.y, c.y))), fymax = floorf(fmaxf(a.y, fmaxf(bb.y, c.y)));
  if (!(fxmin <= fxmax) || !(fymin <= fymax)) return;
  int xmin = fxmin <= 0.f ? 0 : (fxmin >= (float)D ? D : (int)fxmin);
  int xmax = fxmax >= (float)(D - 1) ? D - 1 : (fxmax < 0.f ? -1 : (int)fxmax);
  int ymin = fymin <= 0.f ? 0 : (fymin >= (float)D ? D : (int)fymin);
  int ymax = fymax >= (float)(D - 1) ? D - 1 : (fymax < 0.f ? -1 : (int)fymax);
  const int nx = xmax - xmin + 1, ny = ymax - ymin + 1;
  if (nx <= 0 || ny <= 0) return;
  uint32_t *par = ws.parity + (size_t)b * D * D * ws.W;
  for (int k = lane; k < nx * ny; k += RASTER_LANES) {
    const int X = xmin + k / ny, Y = ymin + k % ny;
    float zc;
    if (!column_hits_triangle(ia, ib, ic, a, bb, c, (float)X, (float)Y, &zc)) continue;
    int nz = count_below(zc, D);
    if (nz <= 0) continue;
    uint32_t *col = par + ((size_t)X * D + Y) * ws.W;
    int full = nz >> 5, rem = nz & 31;
    for (int w = 0; w < full; ++w) atomicXor(col + w, 0xFFFFFFFFu);
    if (rem) atomicXor(col + full, (1u << rem) - 1u);
  }
}

// ----------------------------------------------------------------------------- k_compact
// one lane per (column, 32-voxel word) of the hand's lattice bbox: the word of the parity mask, the 128
// bytes of S under it, the hits; one slot reservation per warp of 32 words.
__global__ void __launch_bounds__(256) k_compact(foho_guidance_desc d, FohoWorkspace ws) {
  FohoTrace trace_(ws.trace, TR_COMPACT);
  const int b = blockIdx.y, D = d.D, lane = threadIdx.x & 31;
  const FohoFrame &fr = ws.frames[b];
  const int nx = fr.hi[0] - fr.lo[0] + 1, ny = fr.hi[1] - fr.lo[1] + 1;
  if (nx <= 0 || ny <= 0 || fr.hi[2] < fr.lo[2]) return;
  const uint32_t *par = ws.parity + (size_t)b * D * D * ws.W;
  const float *S = d.sdf + (size_t)b * D * D * D;
  int *cnt = ws.cnt + (size_t)b * CNT_NUM;
  int *cand = ws.cand + (size_t)b * ws.cap;
  // no crossing lies above the bbox, so words above hi[2] hold no bits (an OPEN mesh leaves bits all
  // the way down below a hole, so the scan starts at word 0)
  const int w0 = 0, w1 = fr.hi[2] >> 5, nw = w1 - w0 + 1;
  const int items = nx * ny * nw;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  // each lane fetches one word (32 independent loads per round), then the warp serves the non-zero ones
  for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; base < items; base += warps * 32) {
    const int k = base + lane;
    uint32_t word = 0u;
    int col = 0, wz = 0;
    if (k < items) {
      wz = w0 + k % nw;
      const int c = k / nw;
      col = (fr.lo[0] + c / ny) * D + (fr.lo[1] + c % ny);
      word = par[(size_t)col * ws.W + wz];
    }
    // pass 1: every lane tests its own word: the 32 field values of the word are one 128-byte line,
    //         fetched with eight independent 16-byte loads (one trip to DRAM for the whole warp's 32 words;
    //         a lane-per-bit scan needed one dependent trip per non-zero word)
    unsigned mymask = 0u;
    if (word != 0u) {
      const size_t v0 = (size_t)col * D + (size_t)wz * 32;
      const float *sp = S + v0;
      if ((D & 3) == 0 && wz * 32 + 32 <= D) {
        const float4 *s4 = reinterpret_cast<const float4 *>(sp);       // (col * D + wz * 32) % 4 == 0
        float4 q[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q[u] = s4[u];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          mymask |= (q[u].x < 0.f ? 1u : 0u) << (4 * u) | (q[u].y < 0.f ? 1u : 0u) << (4 * u + 1) |
                    (q[u].z < 0.f ? 1u : 0u) << (4 * u + 2) | (q[u].w < 0.f ? 1u : 0u) << (4 * u + 3);
        }
      } else {
        for (int z = 0; z < 32 && wz * 32 + z < D; ++z) mymask |= (sp[z] < 0.f ? 1u : 0u) << z;
      }
      mymask &= word;
    }
    // pass 2: one slot reservation for the whole batch, then every lane writes its own word's hits
    const int mine = __popc(mymask);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(cnt + CNT_NCAND, total);
    slot0 = __shfl_sync(0xffffffffu, slot0, 0) + incl - mine;
    unsigned mm = mymask;
    while (mm) {
      const int z = __ffs(mm) - 1;
      mm &= mm - 1;
      if (slot0 < ws.cap) cand[slot0] = col * D + wz * 32 + z;
      else atomicOr(cnt + CNT_FLAGS, 1);
      ++slot0;
    }
  }
}

// ----------------------------------------------------------------------------- k_voxdist
// one warp per candidate voxel: nearest-vertex upper bound, then bounding-sphere culled
// exact closest point over the faces.
__global__ void __launch_bounds__(256) k_voxdist(foho_guidance_desc d, FohoWorkspace ws) {
  FohoTrace trace_(ws.trace, TR_VOXDIST);
  const int b = blockIdx.y, D = d.D, Vh = d.Vh, Fh = d.Fh;
  const int *cnt = ws.cnt + (size_t)b * CNT_NUM;
  int n = cnt[CNT_NCAND];
  if (n > ws.cap) n = ws.cap;
  const int warps_per_cta = blockDim.x >> 5;
  if ((int)(blockIdx.x * warps_per_cta) >= n) return;
  // hand geometry straight from global memory: 9 KB of vertices, 25 KB of face spheres (written by
  // k_raster) and 18 KB of indices per sample stay in L1/L2; no per-CTA staging, no shared memory, so
  // the kernel fits beside the dense stream's CTAs.
  const float4 *__restrict__ sph = ws.sph + (size_t)b * Fh;
  const float *__restrict__ sv = ws.hg + (size_t)b * Vh * 3;
  const int *__restrict__ sf = d.hand_faces;
  const FohoFrame &fr = ws.frames[b];
  const float kappa = fr.kappa;
  const float N = (float)D * (float)D * (float)D;
  const float *S = d.sdf + (size_t)b * D * D * D;
  float *Ghg = ws.G_hg + (size_t)b * Vh * 3;
  const int *cand = ws.cand + (size_t)b * ws.cap;
  float *cval = ws.cand_val + (size_t)b * ws.cap;   // dE/dS of the candidate, added to G by k_assemble
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float acc_int = 0.f, acc_gk = 0.f;
  for (int c = blockIdx.x * warps_per_cta + wid; c < n; c += gridDim.x * warps_per_cta) {
    const int v = cand[c];
    const int Z = v % D, Y = (v / D) % D, X = v / (D * D);
    const foho_f3 p = f3((float)X, (float)Y, (float)Z);
    float ub2 = INFINITY;
    for (int i = lane; i < Vh; i += 32) {
      foho_f3 q = f3(sv[3 * i], sv[3 * i + 1], sv[3 * i + 2]) - p;
      ub2 = fminf(ub2, dot3(q, q));
    }
    ub2 = warp_min(ub2) * 1.00001f + 1e-12f;
    float best2 = ub2, bwa = 0.f, bwb = 0.f, bwc = 0.f;
    int bf = -1;
    for (int f = lane; f < Fh; f += 32) {
      float4 s = sph[f];
      foho_f3 q = f3(s.x, s.y, s.z) - p;
      float lb = sqrtf(dot3(q, q)) - s.w;
      if (lb > 0.f && lb * lb > best2) continue;
      int ia = sf[3 * f], ib = sf[3 * f + 1], ic = sf[3 * f + 2];
      float wa, wb, wc;
      // translate by -p first: differences of nearby lattice coordinates are (nearly) exact in fp32
      float d2 = closest_point_triangle(f3(0.f, 0.f, 0.f), f3(sv[3 * ia], sv[3 * ia + 1], sv[3 * ia + 2]) - p,
                                        f3(sv[3 * ib], sv[3 * ib + 1], sv[3 * ib + 2]) - p,
                                        f3(sv[3 * ic], sv[3 * ic + 1], sv[3 * ic + 2]) - p, wa, wb, wc);
      if (d2 < best2 || bf < 0) {
        if (d2 <= best2) { best2 = d2; bf = f; bwa = wa; bwb = wb; bwc = wc; }
      }
    }
    float key = bf >= 0 ? best2 : INFINITY;
    float mk = warp_min(key);
    unsigned vote = __ballot_sync(0xffffffffu, key == mk && bf >= 0);
    if (vote == 0u) {                                       // cannot happen for a non-empty mesh
      if (lane == 0) cval[c] = 0.f;
      continue;
    }
    int leader = __ffs(vote) - 1;
    if (lane == leader) {
      const float s = S[v];                                  // < 0 by construction
      const float dist = sqrtf(best2);
      const float ns = -s;
      acc_int += ns * dist;
      const float coef = d.w.w_ivol * ns / N;               // dE/d(kappa*dist)
      acc_gk += coef * dist;
      cval[c] = -d.w.w_ivol * kappa * dist / N;
      if (dist > 0.f) {
        int ia = sf[3 * bf], ib = sf[3 * bf + 1], ic = sf[3 * bf + 2];
        foho_f3 a = f3(sv[3 * ia], sv[3 * ia + 1], sv[3 * ia + 2]) - p;
        foho_f3 bb = f3(sv[3 * ib], sv[3 * ib + 1], sv[3 * ib + 2]) - p;
        foho_f3 cc = f3(sv[3 * ic], sv[3 * ic + 1], sv[3 * ic + 2]) - p;
        foho_f3 q = f3(bwa * a.x + bwb * bb.x + bwc * cc.x, bwa * a.y + bwb * bb.y + bwc * cc.y,
                       bwa * a.z + bwb * bb.z + bwc * cc.z);       // closest point relative to p
        float inv = 1.f / dist;
        foho_f3 dir = (-inv) * q;
        float k = -coef * kappa;                            // d(dist)/dv_k = -w_k dir
        atomicAdd(Ghg + 3 * ia, k * bwa * dir.x); atomicAdd(Ghg + 3 * ia + 1, k * bwa * dir.y); atomicAdd(Ghg + 3 * ia + 2, k * bwa * dir.z);
        atomicAdd(Ghg + 3 * ib, k * bwb * dir.x); atomicAdd(Ghg + 3 * ib + 1, k * bwb * dir.y); atomicAdd(Ghg + 3 * ib + 2, k * bwb * dir.z);
        atomicAdd(Ghg + 3 * ic, k * bwc * dir.x); atomicAdd(Ghg + 3 * ic + 1, k * bwc * dir.y); atomicAdd(Ghg + 3 * ic + 2, k * bwc * dir.z);
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

// ----------------------------------------------------------------------------- k_chamfer
constexpr int CH_THREADS = 256;
constexpr int CH_POINTS_PER_CTA = 2048;

__global__ void __launch_bounds__(CH_THREADS) k_chamfer(foho_guidance_desc d, FohoWorkspace ws) {
  FohoTrace trace_(ws.trace, TR_CHAMFER);
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int b = blockIdx.y, Vh = d.Vh, P = d.P;
  float4 *sh = reinterpret_cast<float4 *>(sm_raw);                               // [Vh]
  unsigned long long *best = reinterpret_cast<unsigned long long *>(sh + Vh);    // [Vh]
  float *gacc = reinterpret_cast<float *>(best + Vh);                            // [Vh*3]
  __shared__ float red[32];
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;
  for (int i = threadIdx.x; i < Vh; i += blockDim.x) {
    sh[i] = make_float4(hmc[3 * i], hmc[3 * i + 1], hmc[3 * i + 2], 0.f);
    best[i] = 0xFFFFFFFFFFFFFFFFull;
    gacc[3 * i] = 0.f; gacc[3 * i + 1] = 0.f; gacc[3 * i + 2] = 0.f;
  }
  __syncthreads();
  const FohoFrame &fr = ws.frames[b];
  const float cx = fr.co[0], cy = fr.co[1], cz = fr.co[2];
  const float *cloud = d.cloud + (size_t)b * P * 3;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int base = blockIdx.x * CH_POINTS_PER_CTA;
  const float coef = 2.f * d.w.w_ch / (float)P;
  float sum = 0.f;
  for (int off = wid * 64; off < CH_POINTS_PER_CTA; off += (CH_THREADS / 32) * 64) {
    const int ia = base + off + lane, ib = ia + 32;
    if (base + off >= P) break;
    const bool va = ia < P, vb = ib < P;
    float ax = INFINITY, ay = 0.f, az = 0.f, bx = INFINITY, by = 0.f, bz = 0.f;
    if (va) { ax = cloud[3 * ia] - cx; ay = cloud[3 * ia + 1] - cy; az = cloud[3 * ia + 2] - cz; }
    if (vb) { bx = cloud[3 * ib] - cx; by = cloud[3 * ib + 1] - cy; bz = cloud[3 * ib + 2] - cz; }
    float besta = INFINITY, bestb = INFINITY;
    int ja = 0, jb = 0;
    for (int j = 0; j < Vh; ++j) {
      const float4 h = sh[j];
      float dx = ax - h.x, dy = ay - h.y, dz = az - h.z;
      float da = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      dx = bx - h.x; dy = by - h.y; dz = bz - h.z;
      float db = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if (da < besta) { besta = da; ja = j; }
      if (db < bestb) { bestb = db; jb = j; }
      const float m = fminf(da, db);
      const unsigned mb = __float_as_uint(m);
      const unsigned wm = __reduce_min_sync(0xffffffffu, mb);
      const unsigned cur = (unsigned)(best[j] >> 32);
      if (wm <= cur && wm < 0x7f800000u) {
        unsigned vote = __ballot_sync(0xffffffffu, mb == wm);
        if (lane == __ffs(vote) - 1) {
          unsigned idx = (unsigned)(da <= db ? ia : ib);
          atomicMin(best + j, ((unsigned long long)wm << 32) | idx);
        }
      }
    }
    if (va) {
      sum += besta;
      const float4 h = sh[ja];
      atomicAdd(gacc + 3 * ja, coef * (h.x - ax)); atomicAdd(gacc + 3 * ja + 1, coef * (h.y - ay)); atomicAdd(gacc + 3 * ja + 2, coef * (h.z - az));
    }
    if (vb) {
      sum += bestb;
      const float4 h = sh[jb];
      atomicAdd(gacc + 3 * jb, coef * (h.x - bx)); atomicAdd(gacc + 3 * jb + 1, coef * (h.y - by)); atomicAdd(gacc + 3 * jb + 2, coef * (h.z - bz));
    }
  }
  sum = warp_sum(sum);
  if (lane == 0) red[wid] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < CH_THREADS / 32; ++w) t += red[w];
    atomicAdd(ws.acc + (size_t)b * ACC_NUM + ACC_CH_CLOUD, t);
  }
  float *Ghm = ws.G_hm + (size_t)b * Vh * 3;
  unsigned long long *gknn = ws.knn + (size_t)b * Vh;
  for (int i = threadIdx.x; i < Vh; i += blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float g = gacc[3 * i + a];
      if (g != 0.f) atomicAdd(Ghm + 3 * i + a, g);
    }
    unsigned long long k = best[i];
    if (k != 0xFFFFFFFFFFFFFFFFull) atomicMin(gknn + i, k);
  }
}

// ----------------------------------------------------------------------------- key-points, vertex passes, assembly
constexpr int FIN_THREADS = 512;
constexpr int FIN_NRED = FOHO_FIN_NRED;
__constant__ int c_tips[5] = {744, 320, 443, 554, 671};                          // pipelines.py:127
__constant__ int c_openpose[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};  // :128

// k_keypoints (REF a11, pipelines.py:121-135, 1490-1495): regressed + fingertip key-points, their
// screen projection, the MSE against the HaMeR 2-D key-points and its gradient w.r.t. the 21 3-D
// key-points.  Needs only k_prep's output, so it runs early on a side stream; k_finalize_verts picks
// up ws.kpbuf = [21*3 gradient | loss].
constexpr int KP_THREADS = 512;
__global__ void __launch_bounds__(KP_THREADS) k_keypoints(foho_guidance_desc d, FohoWorkspace ws) {
  FohoTrace trace_(ws.trace, TR_KP);
  __shared__ float kp3[21][3];     // concatenated order: 16 regressed + 5 tips
  __shared__ float kpl[21];
  const int b = blockIdx.x, tid = threadIdx.x, Vh = d.Vh;
  const int lane = tid & 31, wid = tid >> 5;
  const FohoFrame &fr = ws.frames[b];
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;
  float *out = ws.kpbuf + (size_t)b * 64;
  const float co[3] = {fr.co[0], fr.co[1], fr.co[2]};
  if (wid < 16) {
    const float *J = d.j_regressor + (size_t)wid * Vh;
    float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
#pragma unroll 8
    for (int i = lane; i < Vh; i += 32) {
      const float w = J[i];
      sx = fmaf(w, hmc[3 * i], sx); sy = fmaf(w, hmc[3 * i + 1], sy); sz = fmaf(w, hmc[3 * i + 2], sz);
      sw += w;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sw = warp_sum(sw);
    // rows of J sum to 1 for MANO but do not rely on it: add c_o * sum(J)
    if (lane == 0) { kp3[wid][0] = sx + sw * co[0]; kp3[wid][1] = sy + sw * co[1]; kp3[wid][2] = sz + sw * co[2]; }
  }
  if (tid < 5) {
    int i = c_tips[tid];
    kp3[16 + tid][0] = hmc[3 * i] + co[0]; kp3[16 + tid][1] = hmc[3 * i + 1] + co[1]; kp3[16 + tid][2] = hmc[3 * i + 2] + co[2];
  }
  __syncthreads();
  if (tid < 21) {
    const int c = c_openpose[tid];                 // out[tid] = cat[c]
    const float x = kp3[c][0], y = kp3[c][1], z = kp3[c][2];
    const float th = tanf(d.fov_deg * 0.017453292519943295f * 0.5f);
    const float H = (float)d.image_h, W = (float)d.image_w;
    const float sc = fminf(H, W) * 0.5f;
    const float xv = -x, yv = y, zv = -z;
    const float iz = 1.f / (zv * th);
    const float u = W * 0.5f - sc * xv * iz, v = H * 0.5f - sc * yv * iz;
    const float *t = d.kps_2d + ((size_t)b * 21 + tid) * 2;
    const float du = u - t[0], dv = v - t[1];
    kpl[tid] = (du * du + dv * dv) / 42.f;
    const float wk = d.w.w_hand * d.w.w_kp * 2.f / 42.f;
    const float gu = wk * du, gv = wk * dv;
    // u = W/2 + sc x/(zv th) ; v = H/2 - sc y/(zv th) ; zv = -z
    out[3 * c] = gu * sc * iz;
    out[3 * c + 1] = -gv * sc * iz;
    out[3 * c + 2] = (gu * (-sc * xv * iz / zv) + gv * (-sc * yv * iz / zv));
  }
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int k = 0; k < 21; ++k) t += kpl[k];
    out[63] = t;
  }
}

// k_vertex_early: the per-vertex work that needs nothing but k_prep's output (and, for the key-point
// back-projection, k_keypoints' 21 gradients): the a13 trilinear sample of S at every hand vertex with
// its penalties, the dE/dS corner contributions (-> ws.tri_*, applied by k_assemble), the field-gradient
// part of dE/d(lattice position) and the key-point / external part of dE/d(MoGe position).  Runs on a
// side stream long before the searches finish, so that k_finalize_verts only has to combine vectors.
constexpr int VE_THREADS = 256;
__global__ void __launch_bounds__(VE_THREADS) k_vertex_early(foho_guidance_desc d, FohoWorkspace ws) {
  FohoTrace trace_(ws.trace, TR_KP);
  __shared__ float gkp[21][3];     // dE/dkp3 in concatenated order
  __shared__ float red[2 * 32];
  const int b = blockIdx.y, tid = threadIdx.x, Vh = d.Vh, D = d.D;
  const bool use_kp = d.n_joints == 16 && d.j_regressor && d.kps_2d && Vh > 744;
  if (tid < 63) (&gkp[0][0])[tid] = use_kp ? ws.kpbuf[(size_t)b * 64 + tid] : 0.f;
  __syncthreads();
  const float *hg = ws.hg + (size_t)b * Vh * 3;
  const float *S = d.sdf + (size_t)b * D * D * D;
  int *tri_idx = ws.tri_idx + (size_t)b * Vh * 8;
  float *tri_val = ws.tri_val + (size_t)b * Vh * 8;
  float *Ehg = ws.E_hg + (size_t)b * Vh * 3, *Ehm = ws.E_hm + (size_t)b * Vh * 3;
  const float invV = 1.f / (float)Vh;
  const float Dm1 = (float)(D - 1);
  float v2[2] = {0.f, 0.f};        // pen, con sums of this CTA
  const int i = blockIdx.x * blockDim.x + tid;
  if (i < Vh) {
    // a13 trilinear sample with border clamp
    float g[3] = {hg[3 * i], hg[3 * i + 1], hg[3 * i + 2]};
    float fr_[3]; int i0[3]; bool live[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float gc = fminf(fmaxf(g[a], 0.f), Dm1);
      live[a] = (g[a] > 0.f) && (g[a] < Dm1);            // clamp has zero slope outside
      float fl = fminf(floorf(gc), (float)(D - 2));
      i0[a] = (int)fl;
      fr_[a] = gc - fl;
    }
    float c[2][2][2];
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dz = 0; dz < 2; ++dz) c[dx][dy][dz] = S[((size_t)(i0[0] + dx) * D + (i0[1] + dy)) * D + (i0[2] + dz)];
    // key-point back-projection through J and the tips (a11) while the samples are in flight
    float kx = 0.f, ky = 0.f, kz = 0.f;
    if (use_kp) {
      float w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = d.j_regressor[(size_t)j * Vh + i];
#pragma unroll
      for (int j = 0; j < 16; ++j) { kx = fmaf(w[j], gkp[j][0], kx); ky = fmaf(w[j], gkp[j][1], ky); kz = fmaf(w[j], gkp[j][2], kz); }
#pragma unroll
      for (int t = 0; t < 5; ++t)
        if (i == c_tips[t]) { kx += gkp[16 + t][0]; ky += gkp[16 + t][1]; kz += gkp[16 + t][2]; }
    }
    if (d.grad_hand_ext) {
      const float *e = d.grad_hand_ext + ((size_t)b * Vh + i) * 3;
      kx += e[0]; ky += e[1]; kz += e[2];
    }
    Ehm[3 * i] = kx; Ehm[3 * i + 1] = ky; Ehm[3 * i + 2] = kz;
    const float fx = fr_[0], fy = fr_[1], fz = fr_[2];
    float s = 0.f, dsx = 0.f, dsy = 0.f, dsz = 0.f;
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dz = 0; dz < 2; ++dz) {
          float wx = dx ? fx : 1.f - fx, wy = dy ? fy : 1.f - fy, wz = dz ? fz : 1.f - fz;
          float v = c[dx][dy][dz];
          s += wx * wy * wz * v;
          dsx += (dx ? 1.f : -1.f) * wy * wz * v;
          dsy += (dy ? 1.f : -1.f) * wx * wz * v;
          dsz += (dz ? 1.f : -1.f) * wx * wy * v;
        }
    v2[0] = foho_relu(-s);
    v2[1] = foho_relu(fabsf(s) - d.w.con_margin);
    float dLds = 0.f;
    if (s < 0.f) dLds -= d.w.w_pen * invV;
    if (fabsf(s) > d.w.con_margin) dLds += d.w.w_con * invV * (s > 0.f ? 1.f : -1.f);
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dz = 0; dz < 2; ++dz) {
          float wx = dx ? fx : 1.f - fx, wy = dy ? fy : 1.f - fy, wz = dz ? fz : 1.f - fz;
          const int k = dx * 4 + dy * 2 + dz;
          tri_idx[i * 8 + k] = ((i0[0] + dx) * D + (i0[1] + dy)) * D + (i0[2] + dz);
          tri_val[i * 8 + k] = dLds * (wx * wy * wz);
        }
    float e3[3] = {0.f, 0.f, 0.f};
    if (dLds != 0.f) {
      if (live[0]) e3[0] = dLds * dsx;
      if (live[1]) e3[1] = dLds * dsy;
      if (live[2]) e3[2] = dLds * dsz;
    }
    Ehg[3 * i] = e3[0]; Ehg[3 * i + 1] = e3[1]; Ehg[3 * i + 2] = e3[2];
  }
  block_sum<2>(v2, red);
  if (tid == 0) {
    // per-CTA partials, summed in fixed order by k_finalize_verts
    ws.pen_part[((size_t)b * FOHO_VE_MAX_CTAS + blockIdx.x) * 2] = v2[0];
    ws.pen_part[((size_t)b * FOHO_VE_MAX_CTAS + blockIdx.x) * 2 + 1] = v2[1];
  }
}

// k_finalize_verts: everything per hand vertex (runs beside the dense stream: it neither reads the
// stream's moments nor writes G -- its dE/dS corner contributions go to ws.tri_* and are applied by
// k_assemble once the stream has written G).
__global__ void __launch_bounds__(FIN_THREADS) k_finalize_verts(foho_guidance_desc d, FohoWorkspace ws, int ve_ctas) {
  FohoTrace trace_(ws.trace, TR_FIN);
  __shared__ FohoFrame fr;
  __shared__ float red[FIN_NRED * 32];
  const int b = blockIdx.x, tid = threadIdx.x, Vh = d.Vh;
  const bool use_kp = d.n_joints == 16 && d.j_regressor && d.kps_2d && Vh > 744;
  {
    const int nwords = sizeof(FohoFrame) / 4;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(ws.frames + b);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&fr);
    for (int k = tid; k < nwords; k += blockDim.x) dst[k] = src[k];
  }
  __syncthreads();
  const float *hmc = ws.hmc + (size_t)b * Vh * 3;

  // ---- per-vertex pass
  float acc[FIN_NRED];
#pragma unroll
  for (int k = 0; k < FIN_NRED; ++k) acc[k] = 0.f;
  // layout: 0..2 gt_h, 3 gs_h, 4..12 GR_h, 13..15 gt_o, 16 gs_o, 17..25 GR_o, 26 pen, 27 con, 28 ch_hand
  const float *Ghm = ws.G_hm + (size_t)b * Vh * 3;
  const float *Ghg = ws.G_hg + (size_t)b * Vh * 3;
  const float *Ehm = ws.E_hm + (size_t)b * Vh * 3;
  const float *Ehg = ws.E_hg + (size_t)b * Vh * 3;
  const float *rest = d.hand_rest + (size_t)b * Vh * 3;
  const float invV = 1.f / (float)Vh;
  for (int i = tid; i < Vh; i += blockDim.x) {
    // a15 hand -> cloud: issue the dependent load chain first
    unsigned long long key = 0xFFFFFFFFFFFFFFFFull;
    if (d.P > 0 && d.cloud) key = ws.knn[(size_t)b * Vh + i];
    const float ghg[3] = {Ghg[3 * i] + Ehg[3 * i], Ghg[3 * i + 1] + Ehg[3 * i + 1], Ghg[3 * i + 2] + Ehg[3 * i + 2]};
    // gradient w.r.t. the (centred) MoGe position of this vertex
    const foho_f3 m = f3(hmc[3 * i], hmc[3 * i + 1], hmc[3 * i + 2]);
    foho_f3 gm = f3(Ghm[3 * i] + Ehm[3 * i], Ghm[3 * i + 1] + Ehm[3 * i + 1], Ghm[3 * i + 2] + Ehm[3 * i + 2]);
    if (key != 0xFFFFFFFFFFFFFFFFull) {
      const float *p = d.cloud + ((size_t)b * d.P + (unsigned)(key & 0xFFFFFFFFull)) * 3;
      foho_f3 df = m - f3(p[0] - fr.co[0], p[1] - fr.co[1], p[2] - fr.co[2]);
      acc[28] += dot3(df, df);
      gm = gm + (2.f * d.w.w_ch * invV) * df;
    }
    // lattice -> object-centred coordinates
    const foho_f3 gxp = mat3_tmul(fr.Ahs_inv, f3(ghg[0], ghg[1], ghg[2]));       // dE/dx'
    const foho_f3 r = f3(m.x - fr.to[0], m.y - fr.to[1], m.z - fr.to[2]);
    const float iso = 1.f / fr.so;
    const foho_f3 xp = iso * mat3_tmul(fr.Ro, r);                               // x' = R_o^T r / s_o
    const foho_f3 rg = iso * mat3_mul(fr.Ro, gxp);                              // R_o gx' / s_o
    gm = gm + rg;
    acc[13] -= rg.x; acc[14] -= rg.y; acc[15] -= rg.z;
    acc[16] -= dot3(gxp, xp) * iso;
    acc[17] += r.x * gxp.x * iso; acc[18] += r.x * gxp.y * iso; acc[19] += r.x * gxp.z * iso;
    acc[20] += r.y * gxp.x * iso; acc[21] += r.y * gxp.y * iso; acc[22] += r.y * gxp.z * iso;
    acc[23] += r.z * gxp.x * iso; acc[24] += r.z * gxp.y * iso; acc[25] += r.z * gxp.z * iso;
    // hand leaves
    const foho_f3 w = f3(rest[3 * i] - fr.ch[0], rest[3 * i + 1] - fr.ch[1], rest[3 * i + 2] - fr.ch[2]);
    const foho_f3 rw = mat3_mul(fr.Rh, w);
    acc[0] += gm.x; acc[1] += gm.y; acc[2] += gm.z;
    acc[3] += dot3(gm, rw);
    const float sh = fr.sh;
    acc[4] += gm.x * sh * w.x; acc[5] += gm.x * sh * w.y; acc[6] += gm.x * sh * w.z;
    acc[7] += gm.y * sh * w.x; acc[8] += gm.y * sh * w.y; acc[9] += gm.y * sh * w.z;
    acc[10] += gm.z * sh * w.x; acc[11] += gm.z * sh * w.y; acc[12] += gm.z * sh * w.z;
  }
  block_sum<FIN_NRED>(acc, red);
  if (tid == 0) {
    float *fa = ws.fin_acc + (size_t)b * FIN_NRED;
    acc[29] = use_kp ? ws.kpbuf[(size_t)b * 64 + 63] : 0.f;
    for (int k = 0; k < ve_ctas; ++k) {
      acc[26] += ws.pen_part[((size_t)b * FOHO_VE_MAX_CTAS + k) * 2];
      acc[27] += ws.pen_part[((size_t)b * FOHO_VE_MAX_CTAS + k) * 2 + 1];
    }
#pragma unroll
    for (int k = 0; k < FIN_NRED; ++k) fa[k] = acc[k];
  }
}

// k_assemble: after the dense stream AND the sparse chain: (1) adds the deferred sparse dE/dS
// contributions (hand-voxel candidates, trilinear corners) to G, (2) block 0 of each sample folds the
// stream moments and the vertex sums into the 16 leaf gradients and the loss terms.
constexpr int ASM_THREADS = 256;
__global__ void __launch_bounds__(ASM_THREADS) k_assemble(foho_guidance_desc d, FohoWorkspace ws, int stream_gx, int do_voxels) {
  FohoTrace trace_(ws.trace, TR_ASM);
  const int b = blockIdx.y, tid = threadIdx.x, Vh = d.Vh, D = d.D;
  float *G = d.grad_sdf + (size_t)b * D * D * D;
  {
    // index and value of every queued contribution are fetched together (independent loads, one round
    // trip to memory -- the lists were written ~150 us ago and the stream has pushed them out of L2)
    const int *tri_idx = ws.tri_idx + (size_t)b * Vh * 8;
    const float *tri_val = ws.tri_val + (size_t)b * Vh * 8;
    int n = 0;
    const int *cand = ws.cand + (size_t)b * ws.cap;
    const float *cval = ws.cand_val + (size_t)b * ws.cap;
    if (do_voxels) {
      n = ws.cnt[(size_t)b * CNT_NUM + CNT_NCAND];
      if (n > ws.cap) n = ws.cap;
    }
    const int stride = gridDim.x * blockDim.x;
    for (int k0 = blockIdx.x * blockDim.x + tid; k0 < Vh * 8 || k0 < n; k0 += 4 * stride) {
      int ti[4], ci[4];
      float tv[4], cv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * stride;
        const bool t = k < Vh * 8, c = k < n;
        ti[u] = t ? tri_idx[k] : 0; tv[u] = t ? tri_val[k] : 0.f;
        ci[u] = c ? cand[k] : 0;    cv[u] = c ? cval[k] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (tv[u] != 0.f) atomicAdd(G + ti[u], tv[u]);
        if (cv[u] != 0.f) atomicAdd(G + ci[u], cv[u]);
      }
    }
  }
  if (blockIdx.x != 0) return;
  // block 0: stage everything the leaf assembly reads (one round trip for the whole block instead of a
  // chain of dependent loads in one thread: this kernel sits after the dense stream, on the critical path)
  __shared__ FohoFrame fr;
  __shared__ float s_acc[FIN_NRED];
  __shared__ double s_mom[6];
  {
    const int nwords = sizeof(FohoFrame) / 4;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(ws.frames + b);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&fr);
    for (int k = tid; k < nwords; k += blockDim.x) dst[k] = src[k];
    if (tid < FIN_NRED) s_acc[tid] = ws.fin_acc[(size_t)b * FIN_NRED + tid];
    // stream moments: thread m < 6 sums moment m over the stream CTAs in fixed order (double)
    if (tid >= 32 && tid < 38) {
      const int m = tid - 32;
      const float *part = ws.stream_part + (size_t)b * FOHO_MAX_STREAM_CTAS * FOHO_STREAM_PARTIALS;
      double sum = 0;
#pragma unroll 8
      for (int k = 0; k < stream_gx; ++k) sum += part[(size_t)k * FOHO_STREAM_PARTIALS + m];
      s_mom[m] = sum;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const bool use_kp = d.n_joints == 16 && d.j_regressor && d.kps_2d && Vh > 744;
    const float invV = 1.f / (float)Vh;
    float acc[FIN_NRED];
    for (int k = 0; k < FIN_NRED; ++k) acc[k] = s_acc[k];
    const float kp_loss = acc[29];
    const foho_weights &W = d.w;
    const double N = (double)D * D * D;
    const double M0 = s_mom[0], M1x = s_mom[1], M1y = s_mom[2], M1z = s_mom[3], M2 = s_mom[4], cobj = s_mom[5];
    const double so = fr.so, hs = (double)fr.s_h2m * fr.step;
    double Ahs[9];
    for (int k = 0; k < 9; ++k) Ahs[k] = (double)fr.Ah[k] * fr.step;
    double X1[3], Au[3];
    for (int a = 0; a < 3; ++a) {
      X1[a] = Ahs[3 * a] * M1x + Ahs[3 * a + 1] * M1y + Ahs[3 * a + 2] * M1z + (double)fr.u0[a] * M0;
      Au[a] = Ahs[a] * fr.u0[0] + Ahs[3 + a] * fr.u0[1] + Ahs[6 + a] * fr.u0[2];        // (Ahs^T u0)[a]
    }
    const double u02 = (double)fr.u0[0] * fr.u0[0] + (double)fr.u0[1] * fr.u0[1] + (double)fr.u0[2] * fr.u0[2];
    const double X2 = hs * hs * M2 + 2.0 * (Au[0] * M1x + Au[1] * M1y + Au[2] * M1z) + u02 * M0;
    double dv[3], RX1[3];
    for (int a = 0; a < 3; ++a) {
      dv[a] = (double)fr.co[a] + fr.to[a];
      RX1[a] = (double)fr.Ro[3 * a] * X1[0] + (double)fr.Ro[3 * a + 1] * X1[1] + (double)fr.Ro[3 * a + 2] * X1[2];
    }
    const double dRX = dv[0] * RX1[0] + dv[1] * RX1[1] + dv[2] * RX1[2];
    const double d2 = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
    const double Smom = so * so * X2 + 2.0 * so * dRX + d2 * M0;
    const double L_mom = Smom / N;
    const double cm = (double)W.w_mom / N;

    float gth[3] = {acc[0], acc[1], acc[2]};
    float gsh = acc[3];
    float GRh[9], GRo[9];
    for (int k = 0; k < 9; ++k) { GRh[k] = acc[4 + k]; GRo[k] = acc[17 + k]; }
    float gto[3] = {acc[13], acc[14], acc[15]};
    float gso = acc[16];
    // kappa path and the moment term
    const float *ac = ws.acc + (size_t)b * ACC_NUM;
    gso += ac[ACC_GKAPPA] * fr.s_h2m * fr.step;
    gso += (float)(cm * (2.0 * so * X2 + 2.0 * dRX));
    for (int a = 0; a < 3; ++a) {
      gto[a] += (float)(cm * (2.0 * so * RX1[a] + 2.0 * dv[a] * M0));
      for (int c = 0; c < 3; ++c) GRo[3 * a + c] += (float)(cm * 2.0 * so * dv[a] * X1[c]);
    }
    // translation regularisers (pipelines.py:1498,1571)
    const float tr_h = (fr.th[0] * fr.th[0] + fr.th[1] * fr.th[1] + fr.th[2] * fr.th[2]) / 3.f;
    const float tr_o = (fr.to[0] * fr.to[0] + fr.to[1] * fr.to[1] + fr.to[2] * fr.to[2]) / 3.f;
    for (int a = 0; a < 3; ++a) {
      gth[a] += W.w_hand * W.w_treg_h * (2.f / 3.f) * fr.th[a];
      gto[a] += W.w_treg_o * (2.f / 3.f) * fr.to[a];
    }
    const float *th = d.theta + (size_t)b * 16;
    float gqh[4], gqo[4];
    quat_to_mat_backward(th + 4, GRh, gqh);
    quat_to_mat_backward(th + 12, GRo, gqo);
    float *go = d.grad_theta + (size_t)b * 16;
    go[0] = gsh; go[1] = gth[0]; go[2] = gth[1]; go[3] = gth[2];
    go[4] = gqh[0]; go[5] = gqh[1]; go[6] = gqh[2]; go[7] = gqh[3];
    go[8] = gso; go[9] = gto[0]; go[10] = gto[1]; go[11] = gto[2];
    go[12] = gqo[0]; go[13] = gqo[1]; go[14] = gqo[2]; go[15] = gqo[3];

    const int *cn = ws.cnt + (size_t)b * CNT_NUM;
    float *T = d.terms + (size_t)b * FOHO_NUM_TERMS;
    const float L_pen = acc[26] * invV, L_con = acc[27] * invV;
    const float L_int = (float)((double)fr.kappa * ac[ACC_INT] / N);
    const float count = (float)cn[CNT_NCAND] / 1000.f;
    const float L_ch = (d.P > 0 && d.cloud) ? acc[28] * invV + ac[ACC_CH_CLOUD] / (float)d.P : 0.f;
    const float L_kp = use_kp ? kp_loss : 0.f;
    const float w_int = W.w_int_lo;   // REF switch needs the mesh term a7 (mean d2); volume-only path stays at lo
    T[FOHO_T_PEN] = L_pen; T[FOHO_T_CON] = L_con; T[FOHO_T_INT] = L_int; T[FOHO_T_COUNT] = count;
    T[FOHO_T_MOM] = (float)L_mom; T[FOHO_T_CH] = L_ch; T[FOHO_T_KP] = L_kp;
    T[FOHO_T_TREG_H] = tr_h; T[FOHO_T_TREG_O] = tr_o;
    T[FOHO_T_DIST] = 0.f; T[FOHO_T_VREG] = 0.f; T[FOHO_T_EDGE] = 0.f; T[FOHO_T_MEAN_D2] = 0.f;
    T[FOHO_T_NCAND] = (float)cn[CNT_NCAND];
    T[FOHO_T_FLAGS] = (float)cn[CNT_FLAGS];
    if (d.sticky_flags && cn[CNT_FLAGS]) atomicOr(d.sticky_flags + b, cn[CNT_FLAGS]);
    (void)cobj;
    T[FOHO_T_TOTAL] = w_int * count + W.w_treg_o * tr_o + W.w_hand * (W.w_kp * L_kp + W.w_treg_h * tr_h) +
                      W.w_pen * L_pen + W.w_con * L_con + W.w_ivol * L_int + W.w_ch * L_ch + W.w_mom * (float)L_mom;
  }
}

}  // namespace

// ----------------------------------------------------------------------------- fork/join context
// Three high-priority side streams + five events per (device, lane), created on first use (do the first call
// outside a stream capture).  A mutex serialises callers that share them.
struct ForkCtx {
  cudaStream_t side[3];
  cudaEvent_t fork, prep, join_a, join_b, join_c;
  std::mutex mu;
};
constexpr int FORK_LANES = 4;
static ForkCtx *fork_ctx(int lane) {
  static ForkCtx *all[64 * FORK_LANES] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (lane < 0 || lane >= FORK_LANES) lane = 0;
  ForkCtx **ctx = all + lane * 64;
  std::lock_guard<std::mutex> lock(mu);
  if (!ctx[dev]) {
    ForkCtx *c = new ForkCtx();
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);      // hi = numerically lowest = highest priority
    bool ok = true;
    for (int i = 0; i < 3; ++i) ok = ok && cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, hi) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->prep, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->join_a, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->join_b, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&c->join_c, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { delete c; return nullptr; }
    ctx[dev] = c;
  }
  return ctx[dev];
}

FohoDeviceState *foho_device_state() {
  static FohoDeviceState *st[64] = {nullptr};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!st[dev]) {
    FohoDeviceState *s = new FohoDeviceState();
    if (cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { delete s; return nullptr; }
    for (int i = 0; i < FA_NUM; ++i) { s->smem_attr[i] = 0; s->carveout[i] = false; }
    st[dev] = s;
  }
  return st[dev];
}

int foho_func_attrs(const void *func, int id, size_t smem, bool max_carveout) {
  FohoDeviceState *ds = foho_device_state();
  if (!ds) return (int)cudaGetLastError();
  if (smem > 48 * 1024 && smem > ds->smem_attr[id]) {
    FOHO_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ds->smem_attr[id] = smem;
  }
  if (max_carveout && !ds->carveout[id]) {
    FOHO_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    ds->carveout[id] = true;
  }
  return FOHO_OK;
}

// ----------------------------------------------------------------------------- C-ABI
extern "C" int foho_abi_version(void) { return FOHO_ABI_VERSION; }

extern "C" const char *foho_status_string(int s) {
  switch (s) {
    case FOHO_OK: return "ok";
    case FOHO_E_NULL: return "required pointer is NULL";
    case FOHO_E_SHAPE: return "size out of the supported range";
    case FOHO_E_WORKSPACE: return "workspace too small or misaligned";
    case FOHO_E_ARG: return "invalid argument";
    case FOHO_E_DRIVER: return "CUDA driver entry point unavailable or tensor-map encode failed";
    default: return s > 0 ? cudaGetErrorString((cudaError_t)s) : "unknown foho status";
  }
}

extern "C" void foho_default_weights(foho_weights *w) {
  if (!w) return;
  w->w_dist = 10.f; w->w_vreg = 1e-3f; w->w_edge = 1.f; w->w_treg_o = 1e-3f; w->w_hand = 1e-3f;
  w->w_kp = 1e-4f; w->w_treg_h = 1e-2f; w->w_int_lo = 1e-9f; w->w_int_hi = 1e-5f; w->dist_margin = 0.01f;
  w->w_pen = 10.f; w->w_con = 10.f; w->w_ivol = 10.f; w->w_ch = 10.f; w->w_mom = 1e-3f; w->con_margin = 0.01f;
}

extern "C" size_t foho_guidance_workspace_bytes(int32_t B, int32_t D, int32_t Vh, int32_t Fh, int32_t P, int32_t Vo) {
  if (B < 1 || D < 2 || D > 1024 || Vh < 1 || Fh < 1) return 0;
  FohoWorkspace w;
  foho_ws_layout(w, nullptr, B, D, Vh, Fh, P, Vo);
  return w.total;
}

extern "C" int foho_guidance_energy_fwd_bwd(const foho_guidance_desc *dp, void *cuda_stream) {
  if (!dp) return FOHO_E_NULL;
  const foho_guidance_desc &d = *dp;
  if (!d.sdf || !d.grad_sdf || !d.hand_rest || !d.hand_faces || !d.T_h2m || !d.obj_center || !d.theta ||
      !d.grad_theta || !d.terms || !d.workspace)
    return FOHO_E_NULL;
  if (d.B < 1 || d.D < 2 || d.D > 1024 || d.Vh < 1 || d.Vh > 4096 || d.Fh < 1 || d.Fh > 8192 || d.P < 0)
    return FOHO_E_SHAPE;
  if (d.P > 0 && !d.cloud) return FOHO_E_NULL;
  if (d.n_joints != 0 && d.n_joints != 16) return FOHO_E_ARG;
  if (!(d.bound > 0.f)) return FOHO_E_ARG;
  if (d.Vo_total < 0 || d.Eo_total < 0) return FOHO_E_SHAPE;
  if (d.Vo_total > 0 && (!d.obj_verts || !d.obj_vert_offsets)) return FOHO_E_NULL;
  if (d.Eo_total > 0 && (d.Vo_total == 0 || !d.obj_edges || !d.obj_edge_offsets)) return FOHO_E_NULL;
  if (((uintptr_t)d.workspace & 255) != 0) return FOHO_E_WORKSPACE;
  if (((uintptr_t)d.sdf & 15) != 0 || ((uintptr_t)d.grad_sdf & 15) != 0) return FOHO_E_ARG;
  FohoWorkspace ws;
  foho_ws_layout(ws, (char *)d.workspace, d.B, d.D, d.Vh, d.Fh, d.P, d.Vo_total);
  if (ws.total > d.workspace_bytes) return FOHO_E_WORKSPACE;
  ws.trace = (unsigned long long *)d.trace;
  cudaStream_t st = (cudaStream_t)cuda_stream;

  const int sm = d.stage_mask == 0 ? 0x3f : d.stage_mask;
  const bool overlap = d.stage_mask == 0 && d.serial == 0;
  int gx = 0;

  // Stream layout.  overlap: the dense stream depends on nothing but the inputs, so it starts at once
  // on the caller's stream while k_prep and the sparse chains run on three library-owned side streams:
  //   caller : fork ------------------ k_stream --------------------------------- wait(A) k_assemble [obj post]
  //   side A : wait(fork) k_prep rec(P) k_chamfer_c2h*     wait(B) wait(C) [obj pre] k_finalize_verts rec(A)
  //   side B :                  wait(P) k_raster k_compact k_voxdist rec(B)
  //   side C :                  wait(P) k_keypoints k_chamfer_h2c rec(C)
  // Only event record / wait is used, so the same sequence is legal inside a stream capture.
  ForkCtx *fc = nullptr;
  cudaStream_t sa = st, sb = st, sc = st;
  if (overlap) {
    fc = fork_ctx(d.lane);
    if (!fc) return (int)cudaGetLastError();
    sa = fc->side[0]; sb = fc->side[1]; sc = fc->side[2];
    fc->mu.lock();
  }
  struct Unlock { ForkCtx *f; ~Unlock() { if (f) f->mu.unlock(); } } unlock{fc};
  const bool use_kp = d.n_joints == 16 && d.j_regressor && d.kps_2d && d.Vh > 744;

  if (overlap) {
    FOHO_CUDA_TRY(cudaEventRecord(fc->fork, st));
    FOHO_CUDA_TRY(cudaStreamWaitEvent(sa, fc->fork, 0));
  }
  if (sm & 1) {
    k_prep<<<d.B, 256, 0, sa>>>(d, ws);
    FOHO_LAUNCH_CHECK();
  }
  if (overlap) {
    FOHO_CUDA_TRY(cudaEventRecord(fc->prep, sa));
    FOHO_CUDA_TRY(cudaStreamWaitEvent(sb, fc->prep, 0));
    FOHO_CUDA_TRY(cudaStreamWaitEvent(sc, fc->prep, 0));
  }
  if (sm & 2) {
    int rc = foho_launch_stream(dp, ws, &gx, overlap, st);
    if (rc != FOHO_OK) return rc;
  }
  const bool accel = (sm & 4) && d.P > 0 && d.accel;
  if (accel) {
    if (d.Vh > FOHO_ACCEL_HV) return FOHO_E_SHAPE;
    if (((uintptr_t)d.accel & 255) != 0) return FOHO_E_WORKSPACE;
    int rc = foho_launch_chamfer_c2h(dp, ws, sa);
    if (rc != FOHO_OK) return rc;
  } else if ((sm & 4) && d.P > 0) {
    const size_t smem = (size_t)d.Vh * (16 + 8 + 12);
    if (smem > 200 * 1024) return FOHO_E_SHAPE;
    {
      int rc = foho_func_attrs((const void *)k_chamfer, FA_CHAMFER, smem, false);
      if (rc != FOHO_OK) return rc;
    }
    const int nchunk = (d.P + CH_POINTS_PER_CTA - 1) / CH_POINTS_PER_CTA;
    k_chamfer<<<dim3(nchunk, d.B), CH_THREADS, smem, sa>>>(d, ws);
    FOHO_LAUNCH_CHECK();
  }
  if (sm & 8) {
    const bool face_tree = d.accel && d.Fh <= FOHO_ACCEL_FACES;
    const int *face_rank = nullptr;
    if (face_tree) {
      if (((uintptr_t)d.accel & 255) != 0) return FOHO_E_WORKSPACE;
      FohoAccel a;
      foho_accel_layout(a, (char *)d.accel, d.B, d.P);
      if (a.total > d.accel_bytes) return FOHO_E_WORKSPACE;
      face_rank = a.face_rank;
    }
    k_raster<<<dim3((d.Fh * RASTER_LANES + RASTER_THREADS - 1) / RASTER_THREADS, d.B), RASTER_THREADS, 0, sb>>>(d, ws, face_rank);
    FOHO_LAUNCH_CHECK();
    k_compact<<<dim3(32, d.B), 256, 0, sb>>>(d, ws);
    FOHO_LAUNCH_CHECK();
    if (face_tree) {
      int rc = foho_launch_voxdist_staged(dp, ws, sb);
      if (rc != FOHO_OK) return rc;
    } else {
      k_voxdist<<<dim3(64, d.B), 256, 0, sb>>>(d, ws);
      FOHO_LAUNCH_CHECK();
    }
  }
  if (overlap) FOHO_CUDA_TRY(cudaEventRecord(fc->join_b, sb));
  const int ve_ctas = (d.Vh + VE_THREADS - 1) / VE_THREADS;
  if (ve_ctas > FOHO_VE_MAX_CTAS) return FOHO_E_SHAPE;
  if (sm & 16) {
    if (use_kp) {
      k_keypoints<<<d.B, KP_THREADS, 0, sc>>>(d, ws);
      FOHO_LAUNCH_CHECK();
    }
    k_vertex_early<<<dim3(ve_ctas, d.B), VE_THREADS, 0, sc>>>(d, ws);
    FOHO_LAUNCH_CHECK();
  }
  if (accel) {
    int rc = foho_launch_chamfer_h2c(dp, ws, sc);
    if (rc != FOHO_OK) return rc;
  }
  if (overlap) {
    FOHO_CUDA_TRY(cudaEventRecord(fc->join_c, sc));
    FOHO_CUDA_TRY(cudaStreamWaitEvent(sa, fc->join_b, 0));
    FOHO_CUDA_TRY(cudaStreamWaitEvent(sa, fc->join_c, 0));
  }
  const bool obj_mesh = (sm & 32) && d.Vo_total > 0;
  if (obj_mesh) {
    int rc = foho_launch_objmesh_pre(dp, ws, sa);
    if (rc != FOHO_OK) return rc;
  }
  if (sm & 16) {
    k_finalize_verts<<<d.B, FIN_THREADS, 0, sa>>>(d, ws, ve_ctas);
    FOHO_LAUNCH_CHECK();
  }
  if (overlap) {
    FOHO_CUDA_TRY(cudaEventRecord(fc->join_a, sa));
    FOHO_CUDA_TRY(cudaStreamWaitEvent(st, fc->join_a, 0));
  }
  if (sm & 16) {
    k_assemble<<<dim3(16, d.B), ASM_THREADS, 0, st>>>(d, ws, (sm & 2) ? gx : 0, (sm & 8) ? 1 : 0);
    FOHO_LAUNCH_CHECK();
  }
  if (obj_mesh && (sm & 16)) {
    int rc = foho_launch_objmesh_post(dp, ws, st);
    if (rc != FOHO_OK) return rc;
  }
  return FOHO_OK;
}
