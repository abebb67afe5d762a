// Host build of the __host__ __device__ helpers in followmyhold_b200/csrc/foho_math.cuh so
// the CPU test-suite can check the exact arithmetic the kernels use (no GPU on the build box).
// TEST INFRASTRUCTURE ONLY -- never linked into libfoho_b200.so.
#include "../followmyhold_b200/csrc/foho_math.cuh"
#include "../followmyhold_b200/csrc/foho_adamw.cuh"
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

// fused AdamW / step_final arithmetic of k_update / k_update_f16 (foho_adamw.cuh), one tensor of n elements
// with learning rate `lr`; half tensors travel as their 16-bit patterns
static float h2f(uint16_t b) { _Float16 h; memcpy(&h, &b, 2); return (float)h; }
static uint16_t f2h(float f) { _Float16 h = (_Float16)f; uint16_t b; memcpy(&b, &h, 2); return b; }

void host_adamw_f32(float *p, const float *g, float *m, float *v, long n, float beta1, float beta2, float eps,
                    float weight_decay, float lr, int step, const float *x_t, float *x1, float sigma) {
  float lrs[6] = {lr, lr, lr, lr, lr, lr};
  foho_adam_scalars_t s = foho_adam_scalars(beta1, beta2, eps, weight_decay, lrs, lr, step);
  for (long i = 0; i < n; ++i) {
    foho_adamw_one<false>(p[i], g[i], m[i], v[i], s.decay_vel, s.neg_step_vel, s);
    if (x_t && x1) x1[i] = foho_step_final_one<false>(x_t[i], p[i], 1.f - sigma);
  }
}

void host_adamw_f16(uint16_t *p, const uint16_t *g, uint16_t *m, uint16_t *v, long n, float beta1, float beta2,
                    float eps, float weight_decay, float lr, int step, const uint16_t *x_t, uint16_t *x1, float sigma) {
  float lrs[6] = {lr, lr, lr, lr, lr, lr};
  foho_adam_scalars_t s = foho_adam_scalars(beta1, beta2, eps, weight_decay, lrs, lr, step);
  for (long i = 0; i < n; ++i) {
    float pf = h2f(p[i]), mf = h2f(m[i]), vf = h2f(v[i]);
    foho_adamw_one<true>(pf, h2f(g[i]), mf, vf, s.decay_vel, s.neg_step_vel, s);
    p[i] = f2h(pf); m[i] = f2h(mf); v[i] = f2h(vf);
    if (x_t && x1) x1[i] = f2h(foho_step_final_one<true>(h2f(x_t[i]), pf, 1.f - sigma));
  }
}

double host_as_written(float f) { return foho_as_written(f); }

}  // extern "C"
