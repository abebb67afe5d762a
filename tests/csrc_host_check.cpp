// Host build of the __host__ __device__ helpers in followmyhold_b200/csrc/foho_math.cuh so
// the CPU test-suite can check the exact arithmetic the kernels use (no GPU on the build box).
// TEST INFRASTRUCTURE ONLY -- never linked into libfoho_b200.so.
#include "../followmyhold_b200/csrc/foho_math.cuh"
#include <string.h>

extern "C" {

void host_quat_to_mat(const float *q, float *R) { quat_to_mat(q, R); }
void host_quat_backward(const float *q, const float *GR, float *gq) { quat_to_mat_backward(q, GR, gq); }

float host_closest_point(const float *p, const float *a, const float *b, const float *c, float *w3) {
  float wa, wb, wc;
  float d2 = closest_point_triangle(f3(p[0], p[1], p[2]), f3(a[0], a[1], a[2]), f3(b[0], b[1], b[2]),
                                    f3(c[0], c[1], c[2]), wa, wb, wc);
  w3[0] = wa; w3[1] = wb; w3[2] = wc;
  return d2;
}

// same loop structure as k_raster (one face at a time, XOR of the "below" prefix)
void host_raster_parity(const float *hg, const int *faces, int F, int D, unsigned char *inside) {
  memset(inside, 0, (size_t)D * D * D);
  for (int f = 0; f < F; ++f) {
    int ia = faces[3 * f], ib = faces[3 * f + 1], ic = faces[3 * f + 2];
    foho_f3 a = f3(hg[3 * ia], hg[3 * ia + 1], hg[3 * ia + 2]);
    foho_f3 b = f3(hg[3 * ib], hg[3 * ib + 1], hg[3 * ib + 2]);
    foho_f3 c = f3(hg[3 * ic], hg[3 * ic + 1], hg[3 * ic + 2]);
    float fxmin = ceilf(fminf(a.x, fminf(b.x, c.x))), fxmax = floorf(fmaxf(a.x, fmaxf(b.x, c.x)));
    float fymin = ceilf(fminf(a.y, fminf(b.y, c.y))), fymax = floorf(fmaxf(a.y, fmaxf(b.y, c.y)));
    if (!(fxmin <= fxmax) || !(fymin <= fymax)) continue;
    int xmin = fxmin <= 0.f ? 0 : (fxmin >= (float)D ? D : (int)fxmin);
    int xmax = fxmax >= (float)(D - 1) ? D - 1 : (fxmax < 0.f ? -1 : (int)fxmax);
    int ymin = fymin <= 0.f ? 0 : (fymin >= (float)D ? D : (int)fymin);
    int ymax = fymax >= (float)(D - 1) ? D - 1 : (fymax < 0.f ? -1 : (int)fymax);
    for (int X = xmin; X <= xmax; ++X)
      for (int Y = ymin; Y <= ymax; ++Y) {
        float zc;
        if (!column_hits_triangle(ia, ib, ic, a, b, c, (float)X, (float)Y, &zc)) continue;
        int nz = count_below(zc, D);
        for (int Z = 0; Z < nz; ++Z) inside[((size_t)X * D + Y) * D + Z] ^= 1;
      }
  }
}

void host_kabsch(const double *H9, double *R9) {
  double H[3][3], R[3][3];
  for (int i = 0; i < 9; ++i) H[i / 3][i % 3] = H9[i];
  kabsch_rotation(H, R);
  for (int i = 0; i < 9; ++i) R9[i] = R[i / 3][i % 3];
}

}  // extern "C"
