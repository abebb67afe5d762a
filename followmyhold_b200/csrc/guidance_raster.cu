// Row f2, first part: the renderer of the guidance loop and its image-space losses,
// forward and backward to the mesh vertices, as one chain of sm_100a kernels.
//
// What the reference runs every inner iteration (third_party_patches/hy3dgen/shapegen/pipelines.py):
//   norms = renderer(mesh); depth = renderer.rasterizer(mesh).zbuf                      :273-274
//   min/max normalisation of normals (white background included) and of the disparity    :276-287
//   10 * normal_alignment_loss + 10 * l1(disparity) + 10 * bce(silhouette)                :178-187,1567-1569,1580-1583
// with pytorch3d's naive rasteriser (faces_per_pixel = 1), FoVPerspectiveCameras(R = diag(-1,1,-1), T = 0) and the
// PhongNormalShader of pipelines.py:74-92 (pixel colour = SUM of the top face's three vertex normals), camera and
// settings of src/foho/guidance/run.py:84-116.  pytorch3d is not vendored: its arithmetic is restated from memory
// (oracle/raster_oracle.py, PARITY UNPINNED); the kernels follow that restatement formula for formula.
//
// Kernels (one launch each per call, meshes of all B images packed):
//   k_rs_verts      project vertices (NDC xy, view depth), clear accumulators
//   k_rs_facenrm    face cross products -> vertex normal sums (64-bit fixed-point atomics: order independent)
//   k_rs_vertnrm    normalise the vertex normals
//   k_rs_bin        faces -> 16x16-pixel tiles (per-tile lists)
//   k_rs_raster     one CTA per tile: nearest face per pixel (ties -> smaller face index), depth, raw normal colour,
//                   per-image min / max of normals and disparity (ordered-int atomics)
//   k_rs_loss       per pixel: normalised maps, the three losses, d loss / d (raw normal, depth), min / max adjoint sums
//   k_rs_bwd_pix    per pixel: gradients to the three vertex normals and, through the perspective-correct depth
//                   interpolation, to the projected vertices -> world vertices
//   k_rs_bwd_vnrm   vertex-normal normalisation backward
//   k_rs_bwd_face   cross-product backward
//   k_rs_final      fixed point -> float gradients, loss terms
#include "foho_common.cuh"

namespace {

constexpr int TILE = 16;
constexpr float K_EPS = 1e-8f;
constexpr float BG_DEPTH = 10.f, EPS_RANGE = 1e-6f;
constexpr double FX_N = 17592186044416.0;       // 2^44: vertex-normal sums (|cross| << 1)
constexpr double FX_G = 1099511627776.0;        // 2^40: gradients and loss sums

struct RsImage {                 // per-image scalars, device
  int nmin, nmax, dmin, dmax;    // ordered-int encodings of min / max of the raw normal colours and the disparity
  int cnt_nlo, cnt_nhi, cnt_dlo, cnt_dhi;   // how many pixel components sit on each extremum (ties share the adjoint)
  long long s_ln, s_ld, s_ls;    // fixed-point sums of the three losses
  long long g_nlo, g_nhi, g_dlo, g_dhi;     // fixed-point adjoints of the four extrema
  int n_valid, overflow, pad0, pad1;
};

struct RsWork {
  float *ndc;                    // [Vt,3] x_ndc, y_ndc, z_view
  long long *nacc;               // [Vt,3] fixed-point vertex-normal sums
  float *vn;                     // [Vt,4] unit vertex normal, |sum|
  long long *gn;                 // [Vt,3] fixed-point dE/d(unit vertex normal)
  long long *gp;                 // [Vt,3] fixed-point dE/d(world vertex)
  float *gs;                     // [Vt,3] dE/d(vertex normal sum)
  int *tile_cnt;                 // [B,tiles]
  int *tile_list;                // [B,tiles,cap]
  int *p2f;                      // [B,H,W] packed face index or -1
  float *zbuf;                   // [B,H,W]
  float *nraw;                   // [B,H,W,3]
  float *gpix;                   // [B,H,W,4] dE/d(raw normal xyz), dE/d(depth)
  RsImage *img;                  // [B]
  int tiles_x, tiles_y, cap;
};

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ void fx_add(long long *p, double v, double scale) {
  atomicAdd(reinterpret_cast<unsigned long long *>(p), (unsigned long long)__double2ll_rn(v * scale));
}
__device__ __forceinline__ float edge_fn(float px, float py, float ax, float ay, float bx, float by) {
  return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}
__device__ __forceinline__ int image_of(const int *offsets, int B, int i) {      // offsets [B+1], i in [0, offsets[B])
  int lo = 0, hi = B;
  while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (i >= offsets[mid]) lo = mid; else hi = mid; }
  return lo;
}

// Two vertex / face sets per image: set 1 = the first V1 vertices / F1 faces with static per-image ranges (the hand),
// set 2 = what follows, with per-image ranges whose counts live on the device (the extracted object mesh: dynamic shapes
// without a host sync -- launches are capacity sized, threads beyond the count leave).  Without set 2 everything is set 1.
__device__ __forceinline__ int vert_image(const foho_raster_desc &d, int i, bool *valid) {
  const int V1 = d.vert_offsets2 ? d.V1 : d.V_total;
  if (i < V1) { *valid = true; return image_of(d.vert_offsets, d.B, i); }
  const int j = i - V1;
  *valid = j < d.vert_offsets2[d.B];
  return *valid ? image_of(d.vert_offsets2, d.B, j) : 0;
}
__device__ __forceinline__ int face_image(const foho_raster_desc &d, int f, bool *valid) {
  const int F1 = d.face_offsets2 ? d.F1 : d.F_total;
  if (f < F1) { *valid = !d.skip_set1; return image_of(d.face_offsets, d.B, f); }
  const int j = f - F1;
  *valid = j < d.face_offsets2[d.B];
  return *valid ? image_of(d.face_offsets2, d.B, j) : 0;
}

// ---------------------------------------------------------------- geometry
__global__ void k_rs_verts(foho_raster_desc d, RsWork w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d.B) {
    RsImage &im = w.img[i];
    im.nmin = im.dmin = 0x7fffffff; im.nmax = im.dmax = (int)0x80000000;
    im.cnt_nlo = im.cnt_nhi = im.cnt_dlo = im.cnt_dhi = 0;
    im.s_ln = im.s_ld = im.s_ls = 0; im.g_nlo = im.g_nhi = im.g_dlo = im.g_dhi = 0;
    im.n_valid = 0; im.overflow = 0;
  }
  if (i < d.B * w.tiles_x * w.tiles_y) w.tile_cnt[i] = 0;
  if (i >= d.V_total) return;
#pragma unroll
  for (int a = 0; a < 3; ++a) { w.nacc[3 * i + a] = 0; w.gn[3 * i + a] = 0; w.gp[3 * i + a] = 0; }
  bool valid;
  const int b = vert_image(d, i, &valid);
  if (!valid) { w.ndc[3 * i] = 0.f; w.ndc[3 * i + 1] = 0.f; w.ndc[3 * i + 2] = -1.f; return; }
  const float t = tanf(d.fov_deg[b] * 0.00872664625997164788f);       // tan(fov / 2)
  // view = X R + T with R = diag(-1, 1, -1), T = 0 (guidance/run.py:84-90)
  const float xv = -d.verts[3 * i], yv = d.verts[3 * i + 1], zv = -d.verts[3 * i + 2];
  w.ndc[3 * i] = xv / (zv * t); w.ndc[3 * i + 1] = yv / (zv * t); w.ndc[3 * i + 2] = zv;
}

__global__ void k_rs_facenrm(foho_raster_desc d, RsWork w) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= d.F_total) return;
  bool valid;
  face_image(d, f, &valid);
  if (!valid) return;
  const int i0 = d.faces[3 * f], i1 = d.faces[3 * f + 1], i2 = d.faces[3 * f + 2];
  const float *v = d.verts;
  const float ax = v[3 * i1] - v[3 * i0], ay = v[3 * i1 + 1] - v[3 * i0 + 1], az = v[3 * i1 + 2] - v[3 * i0 + 2];
  const float bx = v[3 * i2] - v[3 * i0], by = v[3 * i2 + 1] - v[3 * i0 + 1], bz = v[3 * i2 + 2] - v[3 * i0 + 2];
  const double c[3] = {(double)ay * bz - (double)az * by, (double)az * bx - (double)ax * bz, (double)ax * by - (double)ay * bx};
#pragma unroll
  for (int a = 0; a < 3; ++a) { fx_add(w.nacc + 3 * i0 + a, c[a], FX_N); fx_add(w.nacc + 3 * i1 + a, c[a], FX_N); fx_add(w.nacc + 3 * i2 + a, c[a], FX_N); }
}

__global__ void k_rs_vertnrm(foho_raster_desc d, RsWork w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.V_total) return;
  const float sx = (float)((double)w.nacc[3 * i] / FX_N), sy = (float)((double)w.nacc[3 * i + 1] / FX_N), sz = (float)((double)w.nacc[3 * i + 2] / FX_N);
  const float len = sqrtf(sx * sx + sy * sy + sz * sz);
  const float inv = 1.f / fmaxf(len, 1e-6f);                         // F.normalize(eps = 1e-6)
  w.vn[4 * i] = sx * inv; w.vn[4 * i + 1] = sy * inv; w.vn[4 * i + 2] = sz * inv; w.vn[4 * i + 3] = len;
}

// ---------------------------------------------------------------- binning + rasterisation
__global__ void k_rs_bin(foho_raster_desc d, RsWork w) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= d.F_total) return;
  bool valid;
  const int b = face_image(d, f, &valid);
  if (!valid) return;
  const int i0 = d.faces[3 * f], i1 = d.faces[3 * f + 1], i2 = d.faces[3 * f + 2];
  const float x0 = w.ndc[3 * i0], y0 = w.ndc[3 * i0 + 1], z0 = w.ndc[3 * i0 + 2];
  const float x1 = w.ndc[3 * i1], y1 = w.ndc[3 * i1 + 1], z1 = w.ndc[3 * i1 + 2];
  const float x2 = w.ndc[3 * i2], y2 = w.ndc[3 * i2 + 1], z2 = w.ndc[3 * i2 + 2];
  if (fmaxf(z0, fmaxf(z1, z2)) < 0.f) return;                       // behind the camera
  if (!(isfinite(x0) && isfinite(x1) && isfinite(x2) && isfinite(y0) && isfinite(y1) && isfinite(y2))) return;
  const float area = edge_fn(x2, y2, x0, y0, x1, y1);
  if (fabsf(area) <= K_EPS) return;
  // pixel (row i, col j) centre: x = 1 - (2j+1)/W, y = 1 - (2i+1)/H  ->  j = ((1 - x) W - 1) / 2
  const float xmin = fminf(x0, fminf(x1, x2)), xmax = fmaxf(x0, fmaxf(x1, x2));
  const float ymin = fminf(y0, fminf(y1, y2)), ymax = fmaxf(y0, fmaxf(y1, y2));
  int j0 = (int)floorf(((1.f - xmax) * d.W - 1.f) * 0.5f) - 1, j1 = (int)ceilf(((1.f - xmin) * d.W - 1.f) * 0.5f) + 1;
  int r0 = (int)floorf(((1.f - ymax) * d.H - 1.f) * 0.5f) - 1, r1 = (int)ceilf(((1.f - ymin) * d.H - 1.f) * 0.5f) + 1;
  j0 = max(j0, 0); r0 = max(r0, 0); j1 = min(j1, d.W - 1); r1 = min(r1, d.H - 1);
  if (j0 > j1 || r0 > r1) return;
  for (int ty = r0 / TILE; ty <= r1 / TILE; ++ty)
    for (int tx = j0 / TILE; tx <= j1 / TILE; ++tx) {
      const int tile = (b * w.tiles_y + ty) * w.tiles_x + tx;
      const int slot = atomicAdd(w.tile_cnt + tile, 1);
      if (slot < w.cap) w.tile_list[(long long)tile * w.cap + slot] = f;
      else w.img[b].overflow = 1;
    }
}

__global__ void __launch_bounds__(TILE * TILE) k_rs_raster(foho_raster_desc d, RsWork w) {
  const int b = blockIdx.z, ty = blockIdx.y, tx = blockIdx.x;
  const int tile = (b * w.tiles_y + ty) * w.tiles_x + tx;
  const int t = threadIdx.x;
  const int row = ty * TILE + t / TILE, col = tx * TILE + t % TILE;
  const bool in_img = row < d.H && col < d.W;
  const float px = 1.f - (2.f * col + 1.f) / d.W, py = 1.f - (2.f * row + 1.f) / d.H;
  __shared__ float sv[64][9];
  __shared__ int sf[64];
  __shared__ int red[4][8];
  const int n = min(w.tile_cnt[tile], w.cap);
  float best_z = INFINITY;
  int best_f = -1;
  for (int base = 0; base < n; base += 64) {
    const int m = min(64, n - base);
    __syncthreads();
    if (t < m) {
      const int f = w.tile_list[(long long)tile * w.cap + base + t];
      sf[t] = f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int vi = d.faces[3 * f + k];
        sv[t][3 * k] = w.ndc[3 * vi]; sv[t][3 * k + 1] = w.ndc[3 * vi + 1]; sv[t][3 * k + 2] = w.ndc[3 * vi + 2];
      }
    }
    __syncthreads();
    if (in_img)
      for (int k = 0; k < m; ++k) {
        const float x0 = sv[k][0], y0 = sv[k][1], z0 = sv[k][2], x1 = sv[k][3], y1 = sv[k][4], z1 = sv[k][5];
        const float x2 = sv[k][6], y2 = sv[k][7], z2 = sv[k][8];
        const float a = edge_fn(x2, y2, x0, y0, x1, y1) + K_EPS;
        float w0 = edge_fn(px, py, x1, y1, x2, y2) / a, w1 = edge_fn(px, py, x2, y2, x0, y0) / a, w2 = edge_fn(px, py, x0, y0, x1, y1) / a;
        const float t0 = w0 * z1 * z2, t1 = z0 * w1 * z2, t2 = z0 * z1 * w2;
        const float den = fmaxf(t0 + t1 + t2, K_EPS);
        w0 = t0 / den; w1 = t1 / den; w2 = t2 / den;
        const float pz = w0 * z0 + w1 * z1 + w2 * z2;
        if (w0 > 0.f && w1 > 0.f && w2 > 0.f && pz >= 0.f) {
          const int f = sf[k];
          if (pz < best_z || (pz == best_z && f < best_f)) { best_z = pz; best_f = f; }
        }
      }
  }
  float nx = 1.f, ny = 1.f, nz = 1.f, disp = 1.f / (BG_DEPTH + EPS_RANGE);     // white background, depth 10
  if (best_f >= 0) {
    const int i0 = d.faces[3 * best_f], i1 = d.faces[3 * best_f + 1], i2 = d.faces[3 * best_f + 2];
    nx = w.vn[4 * i0] + w.vn[4 * i1] + w.vn[4 * i2];
    ny = w.vn[4 * i0 + 1] + w.vn[4 * i1 + 1] + w.vn[4 * i2 + 1];
    nz = w.vn[4 * i0 + 2] + w.vn[4 * i1 + 2] + w.vn[4 * i2 + 2];
    disp = 1.f / (best_z + EPS_RANGE);
  }
  int lo_n = 0x7fffffff, hi_n = (int)0x80000000, lo_d = 0x7fffffff, hi_d = (int)0x80000000;
  if (in_img) {
    const long long pix = ((long long)b * d.H + row) * d.W + col;
    w.p2f[pix] = best_f;
    w.zbuf[pix] = best_f >= 0 ? best_z : -1.f;
    w.nraw[3 * pix] = nx; w.nraw[3 * pix + 1] = ny; w.nraw[3 * pix + 2] = nz;
    lo_n = min(f2ord(nx), min(f2ord(ny), f2ord(nz))); hi_n = max(f2ord(nx), max(f2ord(ny), f2ord(nz)));
    lo_d = hi_d = f2ord(disp);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lo_n = min(lo_n, __shfl_xor_sync(0xffffffffu, lo_n, o)); hi_n = max(hi_n, __shfl_xor_sync(0xffffffffu, hi_n, o));
    lo_d = min(lo_d, __shfl_xor_sync(0xffffffffu, lo_d, o)); hi_d = max(hi_d, __shfl_xor_sync(0xffffffffu, hi_d, o));
  }
  if ((t & 31) == 0) { red[0][t >> 5] = lo_n; red[1][t >> 5] = hi_n; red[2][t >> 5] = lo_d; red[3][t >> 5] = hi_d; }
  __syncthreads();
  if (t == 0) {
    for (int k = 1; k < TILE * TILE / 32; ++k) {
      red[0][0] = min(red[0][0], red[0][k]); red[1][0] = max(red[1][0], red[1][k]);
      red[2][0] = min(red[2][0], red[2][k]); red[3][0] = max(red[3][0], red[3][k]);
    }
    atomicMin(&w.img[b].nmin, red[0][0]); atomicMax(&w.img[b].nmax, red[1][0]);
    atomicMin(&w.img[b].dmin, red[2][0]); atomicMax(&w.img[b].dmax, red[3][0]);
  }
}

// ---------------------------------------------------------------- losses (forward) and their pixel adjoints
// pass 0: count the pixel components on the extrema (ties share the min / max adjoint evenly, like torch's full
// reductions); pass 1: losses, direct adjoints, extremum adjoint sums
__global__ void k_rs_loss(foho_raster_desc d, RsWork w, int pass) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const long long npix = (long long)d.H * d.W;
  const bool live = pix < npix;
  const long long gp = (long long)b * npix + pix;
  RsImage &im = w.img[b];
  const float nlo = ord2f(im.nmin), nhi = ord2f(im.nmax), dlo = ord2f(im.dmin), dhi = ord2f(im.dmax);
  const float rn_r = nhi - nlo + EPS_RANGE, rd_r = dhi - dlo + EPS_RANGE;
  float n[3] = {0.f, 0.f, 0.f}, z = -1.f;
  int f = -1;
  if (live) { f = w.p2f[gp]; z = w.zbuf[gp]; n[0] = w.nraw[3 * gp]; n[1] = w.nraw[3 * gp + 1]; n[2] = w.nraw[3 * gp + 2]; }
  const bool hit = f >= 0;
  const float zz = hit ? z : BG_DEPTH;
  const float disp = 1.f / (zz + EPS_RANGE);
  if (pass == 0) {
    int c[4] = {0, 0, 0, 0};
    if (live) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { c[0] += n[a] == nlo; c[1] += n[a] == nhi; }
      c[2] = disp == dlo; c[3] = disp == dhi;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      c[k] = __reduce_add_sync(0xffffffffu, c[k]);
    }
    if ((threadIdx.x & 31) == 0) {
      if (c[0]) atomicAdd(&im.cnt_nlo, c[0]);
      if (c[1]) atomicAdd(&im.cnt_nhi, c[1]);
      if (c[2]) atomicAdd(&im.cnt_dlo, c[2]);
      if (c[3]) atomicAdd(&im.cnt_dhi, c[3]);
    }
    return;
  }
  double l_n = 0.0, l_d = 0.0, l_s = 0.0, a_nlo = 0.0, a_nhi = 0.0, a_dlo = 0.0, a_dhi = 0.0;
  if (live) {
    // ---- normals: rn = (n - lo) / r on covered pixels, 0 elsewhere (:279-281); loss = mean_valid (1 - cos) (:178-187)
    const float inv_valid = 1.f / (float)max(d.n_valid[b], 1);
    float g_rn[3] = {0.f, 0.f, 0.f};
    float rn[3] = {0.f, 0.f, 0.f};
    if (hit) { rn[0] = (n[0] - nlo) / rn_r; rn[1] = (n[1] - nlo) / rn_r; rn[2] = (n[2] - nlo) / rn_r; }
    const bool valid = d.gt_mask[gp] != 0;
    if (valid) {
      const float *g = d.gt_normals + 3 * gp;
      const float ln = fmaxf(sqrtf(rn[0] * rn[0] + rn[1] * rn[1] + rn[2] * rn[2]), 1e-12f);
      const float lg = fmaxf(sqrtf(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]), 1e-12f);
      const float u[3] = {rn[0] / ln, rn[1] / ln, rn[2] / ln}, gg[3] = {g[0] / lg, g[1] / lg, g[2] / lg};
      const float cs = u[0] * gg[0] + u[1] * gg[1] + u[2] * gg[2];
      l_n = (double)(1.f - cs) * inv_valid;
      if (hit) {       // d(x/|x|) = (I - u u^T) / |x|; uncovered pixels have rn = 0 * mask: no gradient reaches n
        const float du[3] = {-gg[0] * inv_valid, -gg[1] * inv_valid, -gg[2] * inv_valid};
        const float dot = du[0] * u[0] + du[1] * u[1] + du[2] * u[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) g_rn[a] = d.w_normal * (du[a] - dot * u[a]) / ln;
      }
    }
    float gpn[3] = {0.f, 0.f, 0.f};
    if (hit)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float y = rn[a];
        gpn[a] = g_rn[a] / rn_r;
        a_nlo += (double)(-(1.f - y) / rn_r * g_rn[a]);
        a_nhi += (double)(-y / rn_r * g_rn[a]);
      }
    // ---- disparity: rd = (disp - lo) / r, L1 against the target (:283-285, 1568)
    const float rd = (disp - dlo) / rd_r;
    const float diff = rd - d.gt_disp[gp];
    l_d = (double)fabsf(diff) / (double)npix;
    const float g_rd = d.w_disp * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f)) / (float)npix;
    a_dlo = (double)(-(1.f - rd) / rd_r * g_rd);
    a_dhi = (double)(-rd / rd_r * g_rd);
    // background depth is a constant (:283): only covered pixels pass the gradient on to z
    const float g_z_direct = hit ? -(g_rd / rd_r) / ((zz + EPS_RANGE) * (zz + EPS_RANGE)) : 0.f;
    // ---- silhouette: alpha is 0 / 1 with sigma = 1e-8 (run.py:91-94): value only, logs clamped at -100 like torch
    const float tt = d.gt_sil[gp];
    l_s = (double)(hit ? 100.f * (1.f - tt) : 100.f * tt) / (double)npix;
    w.gpix[4 * gp] = gpn[0]; w.gpix[4 * gp + 1] = gpn[1]; w.gpix[4 * gp + 2] = gpn[2]; w.gpix[4 * gp + 3] = g_z_direct;
  }
  // block sums -> one fixed-point atomic per quantity per warp
  double v[7] = {l_n, l_d, l_s, a_nlo, a_nhi, a_dlo, a_dhi};
#pragma unroll
  for (int k = 0; k < 7; ++k)
#pragma unroll
    for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  if ((threadIdx.x & 31) == 0) {
    fx_add(&im.s_ln, v[0], FX_G); fx_add(&im.s_ld, v[1], FX_G); fx_add(&im.s_ls, v[2], FX_G);
    fx_add(&im.g_nlo, v[3], FX_G); fx_add(&im.g_nhi, v[4], FX_G); fx_add(&im.g_dlo, v[5], FX_G); fx_add(&im.g_dhi, v[6], FX_G);
  }
}

// ---------------------------------------------------------------- backward to the vertices
__global__ void k_rs_bwd_pix(foho_raster_desc d, RsWork w) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const long long npix = (long long)d.H * d.W;
  if (pix >= npix) return;
  const long long gp = (long long)b * npix + pix;
  const int f = w.p2f[gp];
  if (f < 0) return;
  const RsImage &im = w.img[b];
  const float nlo = ord2f(im.nmin), nhi = ord2f(im.nmax), dlo = ord2f(im.dmin), dhi = ord2f(im.dmax);
  const float rd_r = dhi - dlo + EPS_RANGE;
  const float e_nlo = (float)((double)im.g_nlo / FX_G) / (float)max(im.cnt_nlo, 1), e_nhi = (float)((double)im.g_nhi / FX_G) / (float)max(im.cnt_nhi, 1);
  const float e_dlo = (float)((double)im.g_dlo / FX_G) / (float)max(im.cnt_dlo, 1), e_dhi = (float)((double)im.g_dhi / FX_G) / (float)max(im.cnt_dhi, 1);
  const int i0 = d.faces[3 * f], i1 = d.faces[3 * f + 1], i2 = d.faces[3 * f + 2];
  // ---- raw normal colour = n[i0] + n[i1] + n[i2]
  float g[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float nv = w.nraw[3 * gp + a];
    g[a] = w.gpix[4 * gp + a] + (nv == nlo ? e_nlo : 0.f) + (nv == nhi ? e_nhi : 0.f);
  }
  if (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      fx_add(w.gn + 3 * i0 + a, g[a], FX_G); fx_add(w.gn + 3 * i1 + a, g[a], FX_G); fx_add(w.gn + 3 * i2 + a, g[a], FX_G);
    }
  // ---- depth: pz = S / Q, S = b0 + b1 + b2, Q = sum b_i / z_i, b_i = edge_i / (area + eps)
  const float z = w.zbuf[gp];
  const float disp = 1.f / (z + EPS_RANGE);
  float gz = w.gpix[4 * gp + 3];
  const float g_disp_ext = (disp == dlo ? e_dlo : 0.f) + (disp == dhi ? e_dhi : 0.f);
  gz += -g_disp_ext / ((z + EPS_RANGE) * (z + EPS_RANGE));
  (void)rd_r;
  if (gz == 0.f) return;
  const int row = (int)(pix / d.W), col = (int)(pix % d.W);
  const float px = 1.f - (2.f * col + 1.f) / d.W, py = 1.f - (2.f * row + 1.f) / d.H;
  const int vi[3] = {i0, i1, i2};
  float X[3], Y[3], Z[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { X[k] = w.ndc[3 * vi[k]]; Y[k] = w.ndc[3 * vi[k] + 1]; Z[k] = w.ndc[3 * vi[k] + 2]; }
  const float a = edge_fn(X[2], Y[2], X[0], Y[0], X[1], Y[1]) + K_EPS;
  const float e[3] = {edge_fn(px, py, X[1], Y[1], X[2], Y[2]), edge_fn(px, py, X[2], Y[2], X[0], Y[0]), edge_fn(px, py, X[0], Y[0], X[1], Y[1])};
  const float bb[3] = {e[0] / a, e[1] / a, e[2] / a};
  const float S = bb[0] + bb[1] + bb[2], Q = bb[0] / Z[0] + bb[1] / Z[1] + bb[2] / Z[2];
  float g_b[3], g_Z[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    g_b[k] = gz * (1.f / Q - S / (Q * Q * Z[k]));
    g_Z[k] = gz * S * bb[k] / (Q * Q * Z[k] * Z[k]);
  }
  const float g_e[3] = {g_b[0] / a, g_b[1] / a, g_b[2] / a};
  const float g_a = -(g_b[0] * bb[0] + g_b[1] * bb[1] + g_b[2] * bb[2]) / a;
  float gX[3] = {0.f, 0.f, 0.f}, gY[3] = {0.f, 0.f, 0.f};
  // edge(p; A, B) = (px-Ax)(By-Ay) - (py-Ay)(Bx-Ax):  d/dAx = (py-Ay) - (By-Ay), d/dAy = (Bx-Ax) - (px-Ax), d/dBx = -(py-Ay), d/dBy = (px-Ax)
  auto edge_bwd = [&](float qx, float qy, int A, int Bv, float gcoef, bool q_is_vertex, int Qv) {
    gX[A] += gcoef * ((qy - Y[A]) - (Y[Bv] - Y[A]));
    gY[A] += gcoef * ((X[Bv] - X[A]) - (qx - X[A]));
    gX[Bv] += gcoef * (-(qy - Y[A]));
    gY[Bv] += gcoef * (qx - X[A]);
    if (q_is_vertex) { gX[Qv] += gcoef * (Y[Bv] - Y[A]); gY[Qv] += gcoef * (-(X[Bv] - X[A])); }
  };
  edge_bwd(px, py, 1, 2, g_e[0], false, 0);
  edge_bwd(px, py, 2, 0, g_e[1], false, 0);
  edge_bwd(px, py, 0, 1, g_e[2], false, 0);
  edge_bwd(X[2], Y[2], 0, 1, g_a, true, 2);        // area = edge(v2; v0, v1)
  const float tfov = tanf(d.fov_deg[b] * 0.00872664625997164788f);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    // x_ndc = x_v / (z_v t), y_ndc = y_v / (z_v t); view = (-x, y, -z) of the world vertex
    const float g_xv = gX[k] / (Z[k] * tfov), g_yv = gY[k] / (Z[k] * tfov);
    const float g_zv = g_Z[k] - (gX[k] * X[k] + gY[k] * Y[k]) / Z[k];
    fx_add(w.gp + 3 * vi[k], -g_xv, FX_G); fx_add(w.gp + 3 * vi[k] + 1, g_yv, FX_G); fx_add(w.gp + 3 * vi[k] + 2, -g_zv, FX_G);
  }
}

__global__ void k_rs_bwd_vnrm(foho_raster_desc d, RsWork w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d.V_total) return;
  const float g[3] = {(float)((double)w.gn[3 * i] / FX_G), (float)((double)w.gn[3 * i + 1] / FX_G), (float)((double)w.gn[3 * i + 2] / FX_G)};
  const float n[3] = {w.vn[4 * i], w.vn[4 * i + 1], w.vn[4 * i + 2]};
  const float len = w.vn[4 * i + 3];
  float gs[3] = {0.f, 0.f, 0.f};
  if (len > 1e-6f) {
    const float dot = g[0] * n[0] + g[1] * n[1] + g[2] * n[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) gs[a] = (g[a] - n[a] * dot) / len;
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) gs[a] = g[a] * 1e6f;                 // clamped denominator: n = s / 1e-6
  }
  w.gs[3 * i] = gs[0]; w.gs[3 * i + 1] = gs[1]; w.gs[3 * i + 2] = gs[2];
}

__global__ void k_rs_bwd_face(foho_raster_desc d, RsWork w) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= d.F_total) return;
  bool valid;
  face_image(d, f, &valid);
  if (!valid) return;
  const int i0 = d.faces[3 * f], i1 = d.faces[3 * f + 1], i2 = d.faces[3 * f + 2];
  const float gc[3] = {w.gs[3 * i0] + w.gs[3 * i1] + w.gs[3 * i2], w.gs[3 * i0 + 1] + w.gs[3 * i1 + 1] + w.gs[3 * i2 + 1],
                       w.gs[3 * i0 + 2] + w.gs[3 * i1 + 2] + w.gs[3 * i2 + 2]};
  if (gc[0] == 0.f && gc[1] == 0.f && gc[2] == 0.f) return;
  const float *v = d.verts;
  const float a[3] = {v[3 * i1] - v[3 * i0], v[3 * i1 + 1] - v[3 * i0 + 1], v[3 * i1 + 2] - v[3 * i0 + 2]};
  const float bb[3] = {v[3 * i2] - v[3 * i0], v[3 * i2 + 1] - v[3 * i0 + 1], v[3 * i2 + 2] - v[3 * i0 + 2]};
  // c = a x b:  dE/da = b x gc, dE/db = gc x a
  const float ga[3] = {bb[1] * gc[2] - bb[2] * gc[1], bb[2] * gc[0] - bb[0] * gc[2], bb[0] * gc[1] - bb[1] * gc[0]};
  const float gb[3] = {gc[1] * a[2] - gc[2] * a[1], gc[2] * a[0] - gc[0] * a[2], gc[0] * a[1] - gc[1] * a[0]};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    fx_add(w.gp + 3 * i1 + k, ga[k], FX_G); fx_add(w.gp + 3 * i2 + k, gb[k], FX_G); fx_add(w.gp + 3 * i0 + k, -(ga[k] + gb[k]), FX_G);
  }
}

__global__ void k_rs_final(foho_raster_desc d, RsWork w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < d.B) {
    const RsImage &im = w.img[i];
    float *L = d.losses + 8 * i;
    const float ln = (float)((double)im.s_ln / FX_G), ld = (float)((double)im.s_ld / FX_G), ls = (float)((double)im.s_ls / FX_G);
    L[0] = ln; L[1] = ld; L[2] = ls; L[3] = d.w_normal * ln + d.w_disp * ld + d.w_sil * ls;
    L[4] = ord2f(im.nmin); L[5] = ord2f(im.nmax); L[6] = ord2f(im.dmax); L[7] = (float)im.overflow;
  }
  if (i < d.V_total && d.grad_verts)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float g = (float)((double)w.gp[3 * i + a] / FX_G);
      d.grad_verts[3 * i + a] = d.accumulate_grad ? d.grad_verts[3 * i + a] + g : g;
    }
}

size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

size_t carve(const foho_raster_desc &d, char *base, RsWork *w) {
  const int tx = (d.W + TILE - 1) / TILE, ty = (d.H + TILE - 1) / TILE;
  const int cap = d.tile_cap > 0 ? d.tile_cap : 1024;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return base ? base + o : (char *)nullptr; };
  const size_t Vt = d.V_total, npix = (size_t)d.B * d.H * d.W;
  RsWork r;
  r.ndc = (float *)take(Vt * 3 * 4); r.nacc = (long long *)take(Vt * 3 * 8); r.vn = (float *)take(Vt * 4 * 4);
  r.gn = (long long *)take(Vt * 3 * 8); r.gp = (long long *)take(Vt * 3 * 8); r.gs = (float *)take(Vt * 3 * 4);
  r.tile_cnt = (int *)take((size_t)d.B * tx * ty * 4); r.tile_list = (int *)take((size_t)d.B * tx * ty * cap * 4);
  r.p2f = (int *)take(npix * 4); r.zbuf = (float *)take(npix * 4); r.nraw = (float *)take(npix * 12); r.gpix = (float *)take(npix * 16);
  r.img = (RsImage *)take((size_t)d.B * sizeof(RsImage));
  r.tiles_x = tx; r.tiles_y = ty; r.cap = cap;
  if (w) *w = r;
  return off;
}

}  // namespace

extern "C" size_t foho_raster_workspace_bytes(const foho_raster_desc *d) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->V_total <= 0 || d->F_total <= 0) return 0;
  return carve(*d, nullptr, nullptr);
}

extern "C" int foho_raster_losses_fwd_bwd(const foho_raster_desc *dp, void *cuda_stream) {
  if (!dp) return FOHO_E_NULL;
  const foho_raster_desc &d = *dp;
  if ((d.vert_offsets2 == nullptr) != (d.face_offsets2 == nullptr)) return FOHO_E_ARG;
  if (d.vert_offsets2 && (d.V1 < 0 || d.V1 > d.V_total || d.F1 < 0 || d.F1 > d.F_total)) return FOHO_E_SHAPE;
  if (!d.verts || !d.faces || !d.vert_offsets || !d.face_offsets || !d.fov_deg || !d.gt_normals || !d.gt_mask || !d.gt_disp || !d.gt_sil ||
      !d.n_valid || !d.losses || !d.workspace)
    return FOHO_E_NULL;
  if (d.B <= 0 || d.H <= 0 || d.W <= 0 || d.V_total <= 0 || d.F_total <= 0 || d.H > 4096 || d.W > 4096) return FOHO_E_SHAPE;
  if (d.workspace_bytes < foho_raster_workspace_bytes(dp) || (reinterpret_cast<uintptr_t>(d.workspace) & 255)) return FOHO_E_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  RsWork w;
  carve(d, reinterpret_cast<char *>(d.workspace), &w);
  const int T = 256;
  const int nprep = max(max(d.V_total, d.B * w.tiles_x * w.tiles_y), d.B);
  const long long npix = (long long)d.H * d.W;
  k_rs_verts<<<(nprep + T - 1) / T, T, 0, st>>>(d, w);
  k_rs_facenrm<<<(d.F_total + T - 1) / T, T, 0, st>>>(d, w);
  k_rs_vertnrm<<<(d.V_total + T - 1) / T, T, 0, st>>>(d, w);
  k_rs_bin<<<(d.F_total + T - 1) / T, T, 0, st>>>(d, w);
  k_rs_raster<<<dim3(w.tiles_x, w.tiles_y, d.B), TILE * TILE, 0, st>>>(d, w);
  const dim3 gpix((unsigned)((npix + T - 1) / T), d.B);
  k_rs_loss<<<gpix, T, 0, st>>>(d, w, 0);
  k_rs_loss<<<gpix, T, 0, st>>>(d, w, 1);
  if (d.grad_verts) {
    k_rs_bwd_pix<<<gpix, T, 0, st>>>(d, w);
    k_rs_bwd_vnrm<<<(d.V_total + T - 1) / T, T, 0, st>>>(d, w);
    k_rs_bwd_face<<<(d.F_total + T - 1) / T, T, 0, st>>>(d, w);
  }
  k_rs_final<<<(max(d.V_total, d.B) + T - 1) / T, T, 0, st>>>(d, w);
  if (d.out_p2f) FOHO_CUDA_TRY(cudaMemcpyAsync(d.out_p2f, w.p2f, (size_t)d.B * npix * 4, cudaMemcpyDeviceToDevice, st));
  if (d.out_zbuf) FOHO_CUDA_TRY(cudaMemcpyAsync(d.out_zbuf, w.zbuf, (size_t)d.B * npix * 4, cudaMemcpyDeviceToDevice, st));
  if (d.out_nraw) FOHO_CUDA_TRY(cudaMemcpyAsync(d.out_nraw, w.nraw, (size_t)d.B * npix * 12, cudaMemcpyDeviceToDevice, st));
  FOHO_LAUNCH_CHECK();
  return 0;
}
