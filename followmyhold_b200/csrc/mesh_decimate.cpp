// Quadric edge-collapse decimation of a triangle mesh -- HOST code (no kernel in this file).
//
// Replaces `FaceReducer()(mesh)` of the reference's guidance stage (src/foho/guidance/run.py:161), i.e.
// hy3dgen.shapegen.postprocessors.reduce_face -> pymeshlab `meshing_decimation_quadric_edge_collapse(
// targetfacenum=40000, qualitythr=1.0, preserveboundary=True, boundaryweight=3, preservenormal=True,
// preservetopology=True, autoclean=True)`.  Neither hy3dgen nor MeshLab is in the reference tree: this is the
// published Garland-Heckbert algorithm with the same options (area-weighted face quadrics, boundary planes
// weighted by `boundary_weight`, optimal placement, link condition, normal-flip rejection), not a bit-level
// twin of MeshLab's queue order.  It runs once per image after the loop, on the CPU like the reference's.
#include "../../include/foho_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <utility>
#include <vector>

namespace {

struct Quadric {
  double a[10];   // [a0 a1 a2 a3; a1 a4 a5 a6; a2 a5 a7 a8; a3 a6 a8 a9]
  Quadric() { for (double &x : a) x = 0.0; }
  void add_plane(const double n[3], double d, double w) {
    a[0] += w * n[0] * n[0]; a[1] += w * n[0] * n[1]; a[2] += w * n[0] * n[2]; a[3] += w * n[0] * d;
    a[4] += w * n[1] * n[1]; a[5] += w * n[1] * n[2]; a[6] += w * n[1] * d;
    a[7] += w * n[2] * n[2]; a[8] += w * n[2] * d;
    a[9] += w * d * d;
  }
  void add(const Quadric &o) { for (int k = 0; k < 10; ++k) a[k] += o.a[k]; }
  double eval(const double p[3]) const {
    const double x = p[0], y = p[1], z = p[2];
    return a[0] * x * x + 2 * a[1] * x * y + 2 * a[2] * x * z + 2 * a[3] * x + a[4] * y * y + 2 * a[5] * y * z +
           2 * a[6] * y + a[7] * z * z + 2 * a[8] * z + a[9];
  }
  // minimiser of the quadric, false when the 3x3 block is (numerically) singular
  bool optimum(double p[3]) const {
    const double A00 = a[0], A01 = a[1], A02 = a[2], A11 = a[4], A12 = a[5], A22 = a[7];
    const double c00 = A11 * A22 - A12 * A12, c01 = A02 * A12 - A01 * A22, c02 = A01 * A12 - A02 * A11;
    const double det = A00 * c00 + A01 * c01 + A02 * c02;
    const double scale = std::fabs(A00) + std::fabs(A11) + std::fabs(A22);
    if (!(std::fabs(det) > 1e-9 * scale * scale * scale) || scale == 0.0) return false;
    const double c11 = A00 * A22 - A02 * A02, c12 = A01 * A02 - A00 * A12, c22 = A00 * A11 - A01 * A01;
    const double b0 = -a[3], b1 = -a[6], b2 = -a[8];
    p[0] = (c00 * b0 + c01 * b1 + c02 * b2) / det;
    p[1] = (c01 * b0 + c11 * b1 + c12 * b2) / det;
    p[2] = (c02 * b0 + c12 * b1 + c22 * b2) / det;
    return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]);
  }
};

inline void sub3(const double *a, const double *b, double *r) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
inline void cross3(const double *a, const double *b, double *r) {
  r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
inline double dot3d(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

struct Candidate {
  double cost;
  int32_t u, v;
  uint32_t su, sv;   // vertex stamps when the candidate was made
  bool operator<(const Candidate &o) const {      // min-heap on cost; ties by vertex ids: deterministic
    if (cost != o.cost) return cost > o.cost;
    if (u != o.u) return u > o.u;
    return v > o.v;
  }
};

struct Decimator {
  std::vector<double> P;                 // 3 per vertex
  std::vector<int32_t> F;                // 3 per face
  std::vector<char> fvalid, vvalid;
  std::vector<uint32_t> stamp;
  std::vector<Quadric> Q;
  std::vector<std::vector<int32_t>> vf;  // faces around a vertex (may hold dead faces; pruned on traversal)
  std::priority_queue<Candidate> heap;
  int64_t nfaces = 0;

  void face_normal(int32_t f, double n[3]) const {
    double e1[3], e2[3];
    sub3(&P[3 * F[3 * f + 1]], &P[3 * F[3 * f]], e1);
    sub3(&P[3 * F[3 * f + 2]], &P[3 * F[3 * f]], e2);
    cross3(e1, e2, n);
  }
  void prune(int32_t v) {
    auto &l = vf[v];
    size_t k = 0;
    for (size_t i = 0; i < l.size(); ++i)
      if (fvalid[l[i]]) l[k++] = l[i];
    l.resize(k);
  }
  bool face_has(int32_t f, int32_t v) const { return F[3 * f] == v || F[3 * f + 1] == v || F[3 * f + 2] == v; }

  // neighbours of v with the number of live faces shared with each (1 = boundary edge, 2 = interior)
  void ring(int32_t v, std::vector<std::pair<int32_t, int>> &out) {
    prune(v);
    out.clear();
    for (int32_t f : vf[v])
      for (int k = 0; k < 3; ++k) {
        const int32_t w = F[3 * f + k];
        if (w == v) continue;
        bool found = false;
        for (auto &pr : out)
          if (pr.first == w) { ++pr.second; found = true; break; }
        if (!found) out.emplace_back(w, 1);
      }
  }

  void push(int32_t u, int32_t v) {
    if (u > v) std::swap(u, v);
    Quadric q = Q[u];
    q.add(Q[v]);
    double p[3];
    double cost;
    if (q.optimum(p)) {
      cost = q.eval(p);
    } else {
      const double *cands[3] = {&P[3 * u], &P[3 * v], nullptr};
      double mid[3] = {0.5 * (P[3 * u] + P[3 * v]), 0.5 * (P[3 * u + 1] + P[3 * v + 1]), 0.5 * (P[3 * u + 2] + P[3 * v + 2])};
      cands[2] = mid;
      cost = q.eval(cands[0]);
      for (int k = 1; k < 3; ++k) cost = std::min(cost, q.eval(cands[k]));
    }
    if (!(cost > 0.0)) cost = 0.0;
    heap.push({cost, u, v, stamp[u], stamp[v]});
  }

  // position of the merged vertex for a legal collapse, chosen again at pop time (quadrics may have changed)
  void placement(int32_t u, int32_t v, bool ub, bool vb, double p[3]) const {
    if (ub != vb) {                       // one end on the boundary: the boundary keeps its shape
      const double *s = ub ? &P[3 * u] : &P[3 * v];
      p[0] = s[0]; p[1] = s[1]; p[2] = s[2];
      return;
    }
    Quadric q = Q[u];
    q.add(Q[v]);
    if (q.optimum(p)) {
      // keep the new vertex near the edge: a far-away optimum of an ill-conditioned quadric is not wanted
      double e[3], m[3] = {0.5 * (P[3 * u] + P[3 * v]), 0.5 * (P[3 * u + 1] + P[3 * v + 1]), 0.5 * (P[3 * u + 2] + P[3 * v + 2])};
      sub3(&P[3 * u], &P[3 * v], e);
      double r[3];
      sub3(p, m, r);
      if (dot3d(r, r) <= 4.0 * dot3d(e, e)) return;
    }
    const double mid[3] = {0.5 * (P[3 * u] + P[3 * v]), 0.5 * (P[3 * u + 1] + P[3 * v + 1]), 0.5 * (P[3 * u + 2] + P[3 * v + 2])};
    const double *best = &P[3 * u];
    double bc = q.eval(best);
    if (q.eval(&P[3 * v]) < bc) { best = &P[3 * v]; bc = q.eval(best); }
    if (q.eval(mid) < bc) best = mid;
    p[0] = best[0]; p[1] = best[1]; p[2] = best[2];
  }

  // would moving vertex `v` of face f to p flip the face or make it degenerate?  (v2 = the other collapse end)
  bool flips(int32_t f, int32_t v, const double p[3]) const {
    double n0[3], n1[3];
    face_normal(f, n0);
    const double *q[3];
    for (int k = 0; k < 3; ++k) q[k] = F[3 * f + k] == v ? p : &P[3 * F[3 * f + k]];
    double e1[3], e2[3];
    sub3(q[1], q[0], e1);
    sub3(q[2], q[0], e2);
    cross3(e1, e2, n1);
    const double l0 = std::sqrt(dot3d(n0, n0)), l1 = std::sqrt(dot3d(n1, n1));
    if (!(l1 > 1e-12 * (l0 + 1e-300))) return true;           // collapses to a sliver / point
    return dot3d(n0, n1) < 0.2 * l0 * l1;                      // normal turns by more than ~78 degrees
  }

  bool try_collapse(int32_t u, int32_t v) {
    std::vector<std::pair<int32_t, int>> ru, rv;
    ring(u, ru);
    ring(v, rv);
    int shared_faces = 0;
    bool ub = false, vb = false;
    for (auto &pr : ru) {
      if (pr.second == 1) ub = true;
      if (pr.second > 2) return false;                         // non-manifold fan: leave it alone
      if (pr.first == v) shared_faces = pr.second;
    }
    for (auto &pr : rv) {
      if (pr.second == 1) vb = true;
      if (pr.second > 2) return false;
    }
    if (shared_faces == 0) return false;                       // no longer an edge
    if (ub && vb && shared_faces != 1) return false;           // interior edge between two boundary vertices: would pinch
    // link condition: the common neighbours are exactly the apexes of the faces on the edge
    int common = 0;
    for (auto &a : ru)
      for (auto &b : rv)
        if (a.first == b.first) ++common;
    if (common != shared_faces) return false;
    if ((int64_t)vf[u].size() + (int64_t)vf[v].size() - 2 * shared_faces < 3 && !(ub || vb)) return false;   // would close a pillow
    double p[3];
    placement(u, v, ub, vb, p);
    for (int32_t f : vf[u])
      if (!face_has(f, v) && flips(f, u, p)) return false;
    for (int32_t f : vf[v])
      if (!face_has(f, u) && flips(f, v, p)) return false;
    // ---- collapse v into u
    P[3 * u] = p[0]; P[3 * u + 1] = p[1]; P[3 * u + 2] = p[2];
    Q[u].add(Q[v]);
    for (int32_t f : vf[v]) {
      if (face_has(f, u)) {
        fvalid[f] = 0;
        --nfaces;
      } else {
        for (int k = 0; k < 3; ++k)
          if (F[3 * f + k] == v) F[3 * f + k] = u;
        vf[u].push_back(f);
      }
    }
    vf[v].clear();
    vvalid[v] = 0;
    ++stamp[u];
    ++stamp[v];
    ring(u, ru);
    for (auto &pr : ru) push(u, pr.first);
    return true;
  }
};

}  // namespace

extern "C" int foho_mesh_decimate(const double *verts, int32_t V, const int32_t *faces, int32_t Fn, int32_t target_faces,
                                  double boundary_weight, double *out_verts, int32_t *out_V, int32_t *out_faces,
                                  int32_t *out_F) {
  if (!verts || !faces || !out_verts || !out_V || !out_faces || !out_F) return FOHO_E_NULL;
  if (V < 0 || Fn < 0 || target_faces < 0) return FOHO_E_SHAPE;
  for (int64_t k = 0; k < 3ll * Fn; ++k)
    if (faces[k] < 0 || faces[k] >= V) return FOHO_E_ARG;
  Decimator D;
  D.P.assign(verts, verts + 3ll * V);
  D.F.assign(faces, faces + 3ll * Fn);
  D.fvalid.assign(Fn, 1);
  D.vvalid.assign(V, 1);
  D.stamp.assign(V, 0u);
  D.Q.assign(V, Quadric());
  D.vf.assign(V, {});
  D.nfaces = 0;
  // faces with a repeated vertex carry no area: dropped up front
  for (int32_t f = 0; f < Fn; ++f) {
    const int32_t a = faces[3 * f], b = faces[3 * f + 1], c = faces[3 * f + 2];
    if (a == b || b == c || a == c) { D.fvalid[f] = 0; continue; }
    D.vf[a].push_back(f); D.vf[b].push_back(f); D.vf[c].push_back(f);
    ++D.nfaces;
  }
  if (D.nfaces > target_faces) {
    // face quadrics, area weighted
    std::vector<std::pair<int64_t, int32_t>> edges;       // (min*V+max, face)
    edges.reserve(3 * (size_t)Fn);
    for (int32_t f = 0; f < Fn; ++f) {
      if (!D.fvalid[f]) continue;
      double n[3];
      D.face_normal(f, n);
      const double l = std::sqrt(dot3d(n, n));
      if (l > 0.0) {
        const double un[3] = {n[0] / l, n[1] / l, n[2] / l};
        const double d = -dot3d(un, &D.P[3 * D.F[3 * f]]);
        for (int k = 0; k < 3; ++k) D.Q[D.F[3 * f + k]].add_plane(un, d, 0.5 * l);
      }
      for (int k = 0; k < 3; ++k) {
        int32_t a = D.F[3 * f + k], b = D.F[3 * f + (k + 1) % 3];
        if (a > b) std::swap(a, b);
        edges.emplace_back((int64_t)a * V + b, f);
      }
    }
    std::sort(edges.begin(), edges.end());
    // boundary planes: through the edge, perpendicular to its only face, weight boundary_weight * |e|^2
    for (size_t i = 0; i < edges.size();) {
      size_t j = i;
      while (j < edges.size() && edges[j].first == edges[i].first) ++j;
      const int32_t a = (int32_t)(edges[i].first / V), b = (int32_t)(edges[i].first % V);
      if (j - i == 1 && boundary_weight > 0.0) {
        double n[3], e[3], bn[3];
        D.face_normal(edges[i].second, n);
        sub3(&D.P[3 * b], &D.P[3 * a], e);
        cross3(e, n, bn);
        const double l = std::sqrt(dot3d(bn, bn));
        if (l > 0.0) {
          const double un[3] = {bn[0] / l, bn[1] / l, bn[2] / l};
          const double d = -dot3d(un, &D.P[3 * a]);
          const double w = boundary_weight * dot3d(e, e);
          D.Q[a].add_plane(un, d, w);
          D.Q[b].add_plane(un, d, w);
        }
      }
      i = j;
    }
    for (size_t i = 0; i < edges.size();) {
      size_t j = i;
      while (j < edges.size() && edges[j].first == edges[i].first) ++j;
      D.push((int32_t)(edges[i].first / V), (int32_t)(edges[i].first % V));
      i = j;
    }
    while (D.nfaces > target_faces && !D.heap.empty()) {
      const Candidate c = D.heap.top();
      D.heap.pop();
      if (!D.vvalid[c.u] || !D.vvalid[c.v] || D.stamp[c.u] != c.su || D.stamp[c.v] != c.sv) continue;
      D.try_collapse(c.u, c.v);
    }
  }
  // ---- compact: referenced live vertices in their original order, live faces in their original order
  std::vector<int32_t> remap(V, -1);
  int32_t nv = 0, nf = 0;
  for (int32_t f = 0; f < Fn; ++f) {
    if (!D.fvalid[f]) continue;
    for (int k = 0; k < 3; ++k) remap[D.F[3 * f + k]] = 0;
  }
  for (int32_t v = 0; v < V; ++v)
    if (remap[v] == 0) {
      remap[v] = nv;
      out_verts[3 * nv] = D.P[3 * v]; out_verts[3 * nv + 1] = D.P[3 * v + 1]; out_verts[3 * nv + 2] = D.P[3 * v + 2];
      ++nv;
    }
  for (int32_t f = 0; f < Fn; ++f) {
    if (!D.fvalid[f]) continue;
    for (int k = 0; k < 3; ++k) out_faces[3 * nf + k] = remap[D.F[3 * f + k]];
    ++nf;
  }
  *out_V = nv;
  *out_F = nf;
  return FOHO_OK;
}
