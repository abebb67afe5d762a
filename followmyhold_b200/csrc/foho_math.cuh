// Host/device geometry helpers shared by the guidance kernels.
// Kept __host__ __device__ so tests/csrc_host_check.cpp can exercise the exact same
// arithmetic on the build box (which has no GPU).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FOHO_HD __host__ __device__ __forceinline__
#else
#define FOHO_HD inline
#endif

// max(x, 0) that keeps a NaN, like torch.relu (fmaxf returns 0 for a NaN input, and a NaN volume would slip past the
// NaN guard of pipelines.py:1442-1444,1590-1592)
FOHO_HD float foho_relu(float x) { return x > 0.f ? x : (x != x ? x : 0.f); }

// IEEE single ops that must never be contracted into FMAs: the inside/outside rule is
// specified bit-for-bit (oracle/guidance_oracle.py::raster_parity_inside).
#if defined(__CUDA_ARCH__)
#define FOHO_MUL(a, b) __fmul_rn((a), (b))
#define FOHO_ADD(a, b) __fadd_rn((a), (b))
#define FOHO_SUB(a, b) __fsub_rn((a), (b))
#define FOHO_DIV(a, b) __fdiv_rn((a), (b))
#else
static inline float foho_nofma_mul(float a, float b) { volatile float r = a * b; return r; }
static inline float foho_nofma_add(float a, float b) { volatile float r = a + b; return r; }
static inline float foho_nofma_sub(float a, float b) { volatile float r = a - b; return r; }
static inline float foho_nofma_div(float a, float b) { volatile float r = a / b; return r; }
#define FOHO_MUL(a, b) foho_nofma_mul((a), (b))
#define FOHO_ADD(a, b) foho_nofma_add((a), (b))
#define FOHO_SUB(a, b) foho_nofma_sub((a), (b))
#define FOHO_DIV(a, b) foho_nofma_div((a), (b))
#endif

struct foho_f3 { float x, y, z; };

FOHO_HD foho_f3 f3(float x, float y, float z) { foho_f3 r; r.x = x; r.y = y; r.z = z; return r; }
FOHO_HD foho_f3 operator-(foho_f3 a, foho_f3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
FOHO_HD foho_f3 operator+(foho_f3 a, foho_f3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
FOHO_HD foho_f3 operator*(float s, foho_f3 a) { return f3(s * a.x, s * a.y, s * a.z); }
FOHO_HD float dot3(foho_f3 a, foho_f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// y = M x, M row-major 3x3
FOHO_HD foho_f3 mat3_mul(const float *M, foho_f3 v) {
  return f3(M[0] * v.x + M[1] * v.y + M[2] * v.z, M[3] * v.x + M[4] * v.y + M[5] * v.z,
            M[6] * v.x + M[7] * v.y + M[8] * v.z);
}
// y = M^T x
FOHO_HD foho_f3 mat3_tmul(const float *M, foho_f3 v) {
  return f3(M[0] * v.x + M[3] * v.y + M[6] * v.z, M[1] * v.x + M[4] * v.y + M[7] * v.z,
            M[2] * v.x + M[5] * v.y + M[8] * v.z);
}
FOHO_HD void mat3_matmul(const float *A, const float *B, float *C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
FOHO_HD bool mat3_inverse(const float *A, float *I) {
  float c0 = A[4] * A[8] - A[5] * A[7], c1 = A[5] * A[6] - A[3] * A[8], c2 = A[3] * A[7] - A[4] * A[6];
  float det = A[0] * c0 + A[1] * c1 + A[2] * c2;
  if (det == 0.f) return false;
  float id = 1.f / det;
  I[0] = c0 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
  I[3] = c1 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
  I[6] = c2 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  return true;
}

// pytorch3d.transforms.quaternion_to_matrix semantics (real first, 2/(q.q) scaling;
// reference call sites third_party_patches/hy3dgen/shapegen/pipelines.py:1484,1524).
FOHO_HD void quat_to_mat(const float *q, float *R) {
  float r = q[0], i = q[1], j = q[2], k = q[3];
  float two_s = 2.f / (r * r + i * i + j * j + k * k);
  R[0] = 1.f - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
  R[3] = two_s * (i * j + k * r); R[4] = 1.f - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
  R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = 1.f - two_s * (i * i + j * j);
}

// gq[m] = sum_ab GR[a][b] dR[a][b]/dq[m] for the map above.
FOHO_HD void quat_to_mat_backward(const float *q, const float *GR, float *gq) {
  float r = q[0], i = q[1], j = q[2], k = q[3];
  float n = r * r + i * i + j * j + k * k;
  float two_s = 2.f / n;
  // R = I + two_s * P(q)
  float P[9] = {-(j * j + k * k), i * j - k * r, i * k + j * r,
                i * j + k * r, -(i * i + k * k), j * k - i * r,
                i * k - j * r, j * k + i * r, -(i * i + j * j)};
  float gP = 0.f;
  for (int a = 0; a < 9; ++a) gP += GR[a] * P[a];
  // d two_s / d q_m = -2 * two_s * q_m / n
  float dts = -2.f * two_s / n;
  // dP/dr
  float dr = GR[1] * (-k) + GR[2] * (j) + GR[3] * (k) + GR[5] * (-i) + GR[6] * (-j) + GR[7] * (i);
  float di = GR[1] * (j) + GR[2] * (k) + GR[3] * (j) + GR[4] * (-2.f * i) + GR[5] * (-r) + GR[6] * (k) + GR[7] * (r) +
             GR[8] * (-2.f * i);
  float dj = GR[0] * (-2.f * j) + GR[1] * (i) + GR[2] * (r) + GR[3] * (i) + GR[5] * (k) + GR[6] * (-r) + GR[7] * (k) +
             GR[8] * (-2.f * j);
  float dk = GR[0] * (-2.f * k) + GR[1] * (-r) + GR[2] * (i) + GR[3] * (r) + GR[4] * (-2.f * k) + GR[5] * (j) +
             GR[6] * (i) + GR[7] * (j);
  gq[0] = two_s * dr + dts * r * gP;
  gq[1] = two_s * di + dts * i * gP;
  gq[2] = two_s * dj + dts * j * gP;
  gq[3] = two_s * dk + dts * k * gP;
}

// Closest point on triangle (a,b,c) to p; returns squared distance and barycentric
// weights (Ericson, Real-Time Collision Detection 5.1.5).  Semantics of
// kaolin.metrics.trianglemesh.point_to_mesh_distance per face
// (third_party/utilz/kaolin_sdf_ops.py:100).
FOHO_HD float closest_point_triangle(foho_f3 p, foho_f3 a, foho_f3 b, foho_f3 c, float &wa, float &wb, float &wc) {
  foho_f3 ab = b - a, ac = c - a, ap = p - a;
  float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0.f && d2 <= 0.f) { wa = 1.f; wb = 0.f; wc = 0.f; return dot3(ap, ap); }
  foho_f3 bp = p - b;
  float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0.f && d4 <= d3) { wa = 0.f; wb = 1.f; wc = 0.f; return dot3(bp, bp); }
  float vc = d1 * d4 - d3 * d2;
  if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) {
    float v = d1 / (d1 - d3);
    wa = 1.f - v; wb = v; wc = 0.f;
  } else {
    foho_f3 cp = p - c;
    float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0.f && d5 <= d6) { wa = 0.f; wb = 0.f; wc = 1.f; return dot3(cp, cp); }
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) {
      float w = d2 / (d2 - d6);
      wa = 1.f - w; wb = 0.f; wc = w;
    } else {
      float va = d3 * d6 - d5 * d4;
      if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        wa = 0.f; wb = 1.f - w; wc = w;
      } else {
        float denom = 1.f / (va + vb + vc);
        float v = vb * denom, w = vc * denom;
        wa = 1.f - v - w; wb = v; wc = w;
      }
    }
  }
  foho_f3 q = f3(wa * a.x + wb * b.x + wc * c.x, wa * a.y + wb * b.y + wc * c.y, wa * a.z + wb * b.z + wc * c.z);
  foho_f3 d = p - q;
  return dot3(d, d);
}

// Canonical edge-function sign for the column (X,Y); lo/hi are the edge end points
// ordered by vertex index.  Ties broken by simulation of simplicity.  Returns the sign
// (+1/-1, 0 only for a degenerate edge) and the value through *e.
FOHO_HD int edge_sign_canonical(float lox, float loy, float hix, float hiy, float X, float Y, float *e) {
  float dx = FOHO_SUB(hix, lox), dy = FOHO_SUB(hiy, loy);
  float v = FOHO_SUB(FOHO_MUL(dx, FOHO_SUB(Y, loy)), FOHO_MUL(dy, FOHO_SUB(X, lox)));
  *e = v;
  if (v > 0.f) return 1;
  if (v < 0.f) return -1;
  if (dy != 0.f) return dy > 0.f ? -1 : 1;
  return dx > 0.f ? 1 : (dx < 0.f ? -1 : 0);
}

// Oriented edge u->v (vertex indices iu, iv) evaluated through the canonical form.
FOHO_HD int edge_sign_oriented(int iu, int iv, float ux, float uy, float vx, float vy, float X, float Y, float *e) {
  if (iu < iv) return edge_sign_canonical(ux, uy, vx, vy, X, Y, e);
  float ec;
  int s = -edge_sign_canonical(vx, vy, ux, uy, X, Y, &ec);
  *e = -ec;
  return s;
}

// Does the +z column through (X,Y) hit triangle (a,b,c)?  If so *zc is the crossing
// height.  Bit-exact twin of oracle.guidance_oracle.raster_parity_inside.
FOHO_HD bool column_hits_triangle(int ia, int ib, int ic, foho_f3 a, foho_f3 b, foho_f3 c, float X, float Y, float *zc) {
  float eab, ebc, eca;
  int sab = edge_sign_oriented(ia, ib, a.x, a.y, b.x, b.y, X, Y, &eab);
  int sbc = edge_sign_oriented(ib, ic, b.x, b.y, c.x, c.y, X, Y, &ebc);
  int sca = edge_sign_oriented(ic, ia, c.x, c.y, a.x, a.y, X, Y, &eca);
  if (!(sab == sbc && sbc == sca) || sab == 0) return false;
  float wa = ebc, wb = eca, wc = eab;
  float den = FOHO_ADD(FOHO_ADD(wa, wb), wc);
  if (den == 0.f) return false;
  float num = FOHO_ADD(FOHO_ADD(FOHO_MUL(wa, a.z), FOHO_MUL(wb, b.z)), FOHO_MUL(wc, c.z));
  *zc = FOHO_DIV(num, den);
  return true;
}

// Number of lattice heights Z in [0,D) with float(Z) < zc.
FOHO_HD int count_below(float zc, int D) {
  float c = ceilf(zc);
  if (!(c > 0.f)) return 0;   // also catches NaN
  if (c >= (float)D) return D;
  return (int)c;
}

// ---- float64 3x3 helpers for the similarity Procrustes of the ICP (icp.cu)
FOHO_HD void jacobi_eig3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

// Orthogonal polar factor of a 3x3 matrix with positive determinant by the scaled Newton iteration
// X <- (g X + X^-T / g) / 2, g ~ sqrt(|X^-1|_F / |X|_F) (Higham): a handful of iterations of one reciprocal each.
// The scaling only has to be roughly right (it is taken in single precision and switched off near the fixed point,
// from where the plain iteration squares the error each step), so the result is the polar factor to rounding.  For
// det(H) > 0 this IS U V^T of H = U S V^T.  Returns false (R untouched) when H is not safely orientation preserving
// or the iteration does not settle; the caller then takes the eigen-decomposition path.
FOHO_HD void cofactors3(const double X[3][3], double C[3][3]) {
  C[0][0] = X[1][1] * X[2][2] - X[1][2] * X[2][1]; C[0][1] = X[1][2] * X[2][0] - X[1][0] * X[2][2]; C[0][2] = X[1][0] * X[2][1] - X[1][1] * X[2][0];
  C[1][0] = X[0][2] * X[2][1] - X[0][1] * X[2][2]; C[1][1] = X[0][0] * X[2][2] - X[0][2] * X[2][0]; C[1][2] = X[0][1] * X[2][0] - X[0][0] * X[2][1];
  C[2][0] = X[0][1] * X[1][2] - X[0][2] * X[1][1]; C[2][1] = X[0][2] * X[1][0] - X[0][0] * X[1][2]; C[2][2] = X[0][0] * X[1][1] - X[0][1] * X[1][0];
}

FOHO_HD bool polar_rotation3(const double H[3][3], double R[3][3]) {
  double X[3][3], C[3][3], fro = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) fro += H[i][j] * H[i][j];
  if (!(fro > 0.0) || !(fro < 1e300)) return false;
  const double inv = 1.0 / sqrt(fro);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) X[i][j] = H[i][j] * inv;
  for (int it = 0; it < 40; ++it) {
    cofactors3(X, C);                              // X^-T = C / det
    const double det = X[0][0] * C[0][0] + X[0][1] * C[0][1] + X[0][2] * C[0][2];
    if (!(det > 1e-9)) return false;               // |X|_F = 1 at the start: this small means a (nearly) flat or mirrored H
    double nx = 0.0, nc = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { nx += X[i][j] * X[i][j]; nc += C[i][j] * C[i][j]; }
    const double rdet = 1.0 / det;
    // g^2 = |X^-1|_F / |X|_F = sqrt(nc / nx) / det
    float gf = sqrtf(sqrtf((float)nc / (float)nx) * (float)rdet);
    if (!(gf > 0.95f && gf < 1.05f)) gf = fminf(fmaxf(gf, 1e-3f), 1e3f); else gf = 1.f;
    const double ca = 0.5 * (double)gf, cb = 0.5 * rdet * (double)(1.f / gf);
    double change = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        const double v = ca * X[i][j] + cb * C[i][j];
        const double d = v - X[i][j];
        change += d * d;
        X[i][j] = v;
      }
    if (gf == 1.f && change <= 1e-14) {
      // |dX| <= 1e-7 on a plain step: the error left is ~ |dX|^2 / 2 <= 1e-14; one more plain step squares it again
      cofactors3(X, C);
      const double d2 = X[0][0] * C[0][0] + X[0][1] * C[0][1] + X[0][2] * C[0][2];
      const double r2 = 0.5 / d2;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = 0.5 * X[i][j] + r2 * C[i][j];
      return true;
    }
  }
  return false;
}

// Rotation of trimesh.registration.procrustes(reflection=False): with H = U S V^T,
// R = U diag(1,1,det(U V^T)) V^T.  Computed as R = U' V'^T where U', V' are the
// right-handed completions of the two leading singular pairs (identical result, no
// explicit sign logic; the smallest singular direction absorbs the flip).
FOHO_HD void kabsch_rotation(const double H[3][3], double R[3][3]) {
  if (polar_rotation3(H, R)) return;          // det(H) > 0 (every aligned pair of clouds): U V^T is the polar factor
  double K[3][3], V[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) K[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];   // H^T H
  jacobi_eig3(K, V);
  int o[3] = {0, 1, 2};
  double ev[3] = {K[0][0], K[1][1], K[2][2]};
  for (int a = 0; a < 2; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (ev[o[b]] > ev[o[a]]) { int t = o[a]; o[a] = o[b]; o[b] = t; }
  double v1[3], v2[3], v3[3], u1[3], u2[3], u3[3];
  for (int k = 0; k < 3; ++k) { v1[k] = V[k][o[0]]; v2[k] = V[k][o[1]]; }
  v3[0] = v1[1] * v2[2] - v1[2] * v2[1]; v3[1] = v1[2] * v2[0] - v1[0] * v2[2]; v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
  for (int k = 0; k < 3; ++k) {
    u1[k] = H[k][0] * v1[0] + H[k][1] * v1[1] + H[k][2] * v1[2];
    u2[k] = H[k][0] * v2[0] + H[k][1] * v2[1] + H[k][2] * v2[2];
  }
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int k = 0; k < 3; ++k) u1[k] /= n1;
  double dp = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];          // Gram-Schmidt guards near-degenerate pairs
  for (int k = 0; k < 3; ++k) u2[k] -= dp * u1[k];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  for (int k = 0; k < 3; ++k) u2[k] /= n2;
  u3[0] = u1[1] * u2[2] - u1[2] * u2[1]; u3[1] = u1[2] * u2[0] - u1[0] * u2[2]; u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[i][j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
}

