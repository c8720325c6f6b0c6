// lc3d_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A from-scratch, single-threaded CPU restatement of the PCL semantics used by the
// reference's fine-registration hot path (normative target
// PCL 1.8.1).  It exists to check the CUDA path and to be timed as the CPU
// baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it; the product (lowcost3dreconstruction_b200/,
// liblc3d.so, the CLI tools) never links, imports or calls anything in oracle/.
//
// PARITY UNPINNED: the arithmetic of the reference lives in PCL / FLANN / Eigen,
// which are un-vendored apt dependencies (install/tools_install.sh:36, version
// floating: PCL 1.7.2 / 1.8.1 / 1.10.0) absent from /root/reference and from this
// image; the reference ships no tests, fixtures or golden vectors.
// The restatement is anchored on the reference call sites cited at each function
// and validated against independent implementations (scipy cKDTree, numpy
// SVD/eigh, brute force) in tests/test_oracle_*.py.
//
// Build: g++ -O2 -ffp-contract=off (no FMA contraction: PCL distro builds evaluate
// float expressions with separate mul/add).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#include "../include/lc3d.h"

namespace {

struct P3 {
  float x, y, z;
};

inline const float* xyz_at(const lc3d_cloud* c, int64_t i) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(c->xyz) + i * c->xyz_stride);
}
inline const float* nrm_at(const lc3d_cloud* c, int64_t i) {
  return reinterpret_cast<const float*>(reinterpret_cast<const char*>(c->normal) +
                                        i * c->normal_stride);
}
inline uint32_t rgba_at(const lc3d_cloud* c, int64_t i) {
  return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(c->rgba) +
                                            i * c->rgba_stride);
}
inline float curv_at(const lc3d_cloud* c, int64_t i) {
  return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(c->curvature) +
                                         i * c->curvature_stride);
}
inline bool finite3(const float* p) {
  return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]);
}

// Squared L2 as flann::L2_Simple<float> evaluates it: result starts at
// 0 and accumulates diff*diff for x, y, z in float32, no FMA.
inline float dist2(const float* a, const float* b) {
  float r = 0.0f;
  float d = a[0] - b[0];
  r += d * d;
  d = a[1] - b[1];
  r += d * d;
  d = a[2] - b[2];
  r += d * d;
  return r;
}

// ------------------------------------------------------------------ kd-tree --
// Single kd-tree in the style of FLANN's KDTreeSingleIndex as PCL configures it
// (max_leaf_size 15, reordered data, exact search: checks=-1, eps=0) — SURVEY A.4.
// Results are the exact nearest neighbours under dist2(); ties are resolved to the
// lower point index so that results do not depend on tree shape (PCL's tie
// behaviour is traversal-defined and excluded from parity by the north star).
struct KdTree {
  struct Node {
    int32_t left, right;  // leaf: point range [left,right) in `order`
    int32_t child1, child2;
    int32_t divfeat;  // -1 for leaf
    float divlow, divhigh;
  };
  std::vector<Node> nodes;
  std::vector<int32_t> order;  // tree position -> original index
  std::vector<P3> pts;         // reordered copy
  float bb_lo[3], bb_hi[3];
  int64_t n = 0;
  static constexpr int kLeaf = 15;

  void build(const lc3d_cloud* c) {
    std::vector<P3> in;
    order.clear();
    in.reserve(c->n);
    for (int64_t i = 0; i < c->n; ++i) {
      const float* p = xyz_at(c, i);
      if (!finite3(p)) continue;  // KdTreeFLANN leaves non-finite points out
      order.push_back((int32_t)i);
      in.push_back({p[0], p[1], p[2]});
    }
    n = (int64_t)in.size();
    src_ = in.data();
    idx_.resize(n);
    std::iota(idx_.begin(), idx_.end(), 0);
    nodes.clear();
    nodes.reserve(n / 4 + 16);
    if (n > 0) {
      float lo[3], hi[3];
      bounds(0, (int32_t)n, lo, hi);
      for (int d = 0; d < 3; ++d) {
        bb_lo[d] = lo[d];
        bb_hi[d] = hi[d];
      }
      divide(0, (int32_t)n, lo, hi);
    }
    pts.resize(n);
    std::vector<int32_t> ord2(n);
    for (int64_t i = 0; i < n; ++i) {
      pts[i] = in[idx_[i]];
      ord2[i] = order[idx_[i]];
    }
    order.swap(ord2);
    idx_.clear();
    idx_.shrink_to_fit();
    src_ = nullptr;
  }

  // k-NN of q: writes up to k (index, d2) ascending by (d2, index); returns count.
  // max_d2 < 0: unbounded.
  int knn(const float* q, int k, int32_t* out_idx, float* out_d2) const {
    if (n == 0 || k <= 0) return 0;
    Result r;
    r.k = k;
    r.count = 0;
    r.idx = out_idx;
    r.d2 = out_d2;
    double dists[3] = {0, 0, 0};
    double mind = 0;
    for (int d = 0; d < 3; ++d) {
      if (q[d] < bb_lo[d]) dists[d] = sq((double)q[d] - (double)bb_lo[d]);
      if (q[d] > bb_hi[d]) dists[d] = sq((double)q[d] - (double)bb_hi[d]);
      mind += dists[d];
    }
    search(0, q, mind, dists, r);
    return r.count;
  }

  // Radius search as FLANN's RadiusResultSet does it (KdTreeFLANN::radiusSearch):
  // every indexed point with d2 < r2 STRICTLY, r2 = (float)(radius * radius); unordered.
  void radius(const float* q, float r2, std::vector<int32_t>& out) const {
    out.clear();
    if (n == 0) return;
    double dists[3] = {0, 0, 0};
    double mind = 0;
    for (int d = 0; d < 3; ++d) {
      if (q[d] < bb_lo[d]) dists[d] = sq((double)q[d] - (double)bb_lo[d]);
      if (q[d] > bb_hi[d]) dists[d] = sq((double)q[d] - (double)bb_hi[d]);
      mind += dists[d];
    }
    search_radius(0, q, mind, dists, r2, out);
  }

 private:
  void search_radius(int32_t ni, const float* q, double mind, double* dists, float r2,
                     std::vector<int32_t>& out) const {
    const Node& nd = nodes[ni];
    if (nd.divfeat < 0) {
      for (int32_t i = nd.left; i < nd.right; ++i)
        if (dist2(q, &pts[i].x) < r2) out.push_back(order[i]);
      return;
    }
    int f = nd.divfeat;
    double v = q[f];
    double d1 = v - (double)nd.divlow, d2 = v - (double)nd.divhigh;
    int32_t best, other;
    double cutd;
    if (d1 + d2 < 0) {
      best = nd.child1;
      other = nd.child2;
      cutd = sq(v - (double)nd.divhigh);
    } else {
      best = nd.child2;
      other = nd.child1;
      cutd = sq(v - (double)nd.divlow);
    }
    search_radius(best, q, mind, dists, r2, out);
    double save = dists[f];
    double m2 = mind + cutd - save;
    dists[f] = cutd;
    if (m2 * (1.0 - 1e-6) <= (double)r2) search_radius(other, q, m2, dists, r2, out);
    dists[f] = save;
  }

  const P3* src_ = nullptr;
  std::vector<int32_t> idx_;

  struct Result {
    int k, count;
    int32_t* idx;
    float* d2;
    inline double worst() const {
      return count < k ? std::numeric_limits<double>::infinity() : (double)d2[k - 1];
    }
    inline void add(float d, int32_t i) {
      if (count == k) {
        if (d > d2[k - 1] || (d == d2[k - 1] && i > idx[k - 1])) return;
      }
      int pos = count < k ? count : k - 1;
      while (pos > 0 && (d2[pos - 1] > d || (d2[pos - 1] == d && idx[pos - 1] > i))) {
        d2[pos] = d2[pos - 1];
        idx[pos] = idx[pos - 1];
        --pos;
      }
      d2[pos] = d;
      idx[pos] = i;
      if (count < k) ++count;
    }
  };
  static inline double sq(double v) { return v * v; }
  inline float coord(int32_t i, int d) const { return (&src_[i].x)[d]; }

  void bounds(int32_t l, int32_t r, float* lo, float* hi) const {
    for (int d = 0; d < 3; ++d) lo[d] = hi[d] = coord(idx_[l], d);
    for (int32_t i = l + 1; i < r; ++i)
      for (int d = 0; d < 3; ++d) {
        float v = coord(idx_[i], d);
        lo[d] = std::min(lo[d], v);
        hi[d] = std::max(hi[d], v);
      }
  }

  int32_t divide(int32_t l, int32_t r, float* lo, float* hi) {
    int32_t me = (int32_t)nodes.size();
    nodes.push_back(Node{});
    if (r - l <= kLeaf) {
      nodes[me].left = l;
      nodes[me].right = r;
      nodes[me].divfeat = -1;
      nodes[me].child1 = nodes[me].child2 = -1;
      bounds(l, r, lo, hi);
      return me;
    }
    // widest dimension of the (approximate) box, exact spread, mid-value cut
    // clamped into the data (sliding midpoint), balanced fallback to the median.
    int cut = 0;
    float span = hi[0] - lo[0];
    for (int d = 1; d < 3; ++d)
      if (hi[d] - lo[d] > span) {
        span = hi[d] - lo[d];
        cut = d;
      }
    float mn = coord(idx_[l], cut), mx = mn;
    for (int32_t i = l + 1; i < r; ++i) {
      float v = coord(idx_[i], cut);
      mn = std::min(mn, v);
      mx = std::max(mx, v);
    }
    float cutval = (mn + mx) / 2;
    // three-way partition: [< cutval][== cutval][> cutval]
    int32_t* a = idx_.data();
    int32_t lim1 = l, i = l, lim2 = r;
    while (i < lim2) {
      float v = coord(a[i], cut);
      if (v < cutval)
        std::swap(a[i++], a[lim1++]);
      else if (v > cutval)
        std::swap(a[i], a[--lim2]);
      else
        ++i;
    }
    int32_t half = l + (r - l) / 2, split;
    if (lim1 > half)
      split = lim1;
    else if (lim2 < half)
      split = lim2;
    else
      split = half;
    if (split == l || split == r) {  // all coordinates equal on `cut`: median by index
      split = half;
    }
    float llo[3], lhi[3], rlo[3], rhi[3];
    for (int d = 0; d < 3; ++d) {
      llo[d] = rlo[d] = lo[d];
      lhi[d] = rhi[d] = hi[d];
    }
    lhi[cut] = cutval;
    rlo[cut] = cutval;
    int32_t c1 = divide(l, split, llo, lhi);
    int32_t c2 = divide(split, r, rlo, rhi);
    nodes[me].divfeat = cut;
    nodes[me].child1 = c1;
    nodes[me].child2 = c2;
    nodes[me].divlow = lhi[cut];
    nodes[me].divhigh = rlo[cut];
    for (int d = 0; d < 3; ++d) {
      lo[d] = std::min(llo[d], rlo[d]);
      hi[d] = std::max(lhi[d], rhi[d]);
    }
    return me;
  }

  void search(int32_t ni, const float* q, double mind, double* dists, Result& r) const {
    const Node& nd = nodes[ni];
    if (nd.divfeat < 0) {
      for (int32_t i = nd.left; i < nd.right; ++i) r.add(dist2(q, &pts[i].x), order[i]);
      return;
    }
    int f = nd.divfeat;
    double v = q[f];
    double d1 = v - (double)nd.divlow, d2 = v - (double)nd.divhigh;
    int32_t best, other;
    double cutd;
    if (d1 + d2 < 0) {
      best = nd.child1;
      other = nd.child2;
      cutd = sq(v - (double)nd.divhigh);
    } else {
      best = nd.child2;
      other = nd.child1;
      cutd = sq(v - (double)nd.divlow);
    }
    search(best, q, mind, dists, r);
    double save = dists[f];
    double m2 = mind + cutd - save;
    dists[f] = cutd;
    // conservative (never prunes a subtree that could hold an equal-or-closer point)
    if (m2 * (1.0 - 1e-6) <= r.worst()) search(other, q, m2, dists, r);
    dists[f] = save;
  }
};

// --------------------------------------------------------- small dense math --
struct M4 {
  float m[16];  // row-major
};
inline M4 m4_identity() {
  M4 r;
  std::memset(r.m, 0, sizeof r.m);
  r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
  return r;
}
// Matrix4f product a*b, float32 (final_transformation_ = transformation_ * final_transformation_).
inline M4 m4_mul(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float s = a.m[i * 4 + 0] * b.m[0 * 4 + j];
      s += a.m[i * 4 + 1] * b.m[1 * 4 + j];
      s += a.m[i * 4 + 2] * b.m[2 * 4 + j];
      s += a.m[i * 4 + 3] * b.m[3 * 4 + j];
      r.m[i * 4 + j] = s;
    }
  return r;
}
// p' = T*(x,y,z,1) in float32: ((T0*x + T1*y) + T2*z) + T3   (SURVEY A.2 transformCloud)
inline void xform_point(const M4& T, const float* p, float* o) {
  float x = p[0], y = p[1], z = p[2];
  for (int r = 0; r < 3; ++r) {
    float s = T.m[r * 4 + 0] * x;
    s += T.m[r * 4 + 1] * y;
    s += T.m[r * 4 + 2] * z;
    s += T.m[r * 4 + 3];
    o[r] = s;
  }
}
inline void xform_normal(const M4& T, const float* p, float* o) {
  float x = p[0], y = p[1], z = p[2];
  for (int r = 0; r < 3; ++r) {
    float s = T.m[r * 4 + 0] * x;
    s += T.m[r * 4 + 1] * y;
    s += T.m[r * 4 + 2] * z;
    o[r] = s;
  }
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi), double.  A = V diag(w) V^T.
void jacobi_eig3(const double A[9], double w[3], double V[9]) {
  double a[9];
  std::memcpy(a, A, sizeof a);
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    double diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-300 || off <= 1e-34 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = a[p * 3 + q];
        if (apq == 0.0) continue;
        double app = a[p * 3 + p], aqq = a[q * 3 + q];
        double tau = (aqq - app) / (2.0 * apq);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < 3; ++k) {  // A <- A J
          double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = c * akp - s * akq;
          a[k * 3 + q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {  // A <- J^T A
          double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = c * apk - s * aqk;
          a[q * 3 + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - s * vkq;
          V[k * 3 + q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = a[0];
  w[1] = a[4];
  w[2] = a[8];
}

inline double det3(const double* m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
         m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// Rotation of Umeyama/Kabsch from the cross-covariance S = (1/n) sum (d-mu_d)(s-mu_s)^T:
// S = U D V^T, R = U diag(1,1,sign) V^T with sign = det(U) det(V)  (Eigen::umeyama).
// Computed via the eigen-decomposition of S^T S; the third left vector is u1 x u2,
// which folds det(U) into the construction (valid also for rank-2 S).
void kabsch_rotation(const double S[9], double R[9]) {
  double StS[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += S[k * 3 + i] * S[k * 3 + j];
      StS[i * 3 + j] = s;
    }
  double w[3], V[9];
  jacobi_eig3(StS, w, V);
  int o[3] = {0, 1, 2};
  std::sort(o, o + 3, [&](int a, int b) { return w[a] > w[b]; });
  double Vs[9], U[9];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) Vs[r * 3 + c] = V[r * 3 + o[c]];
  for (int c = 0; c < 2; ++c) {
    double u[3], nrm = 0;
    for (int r = 0; r < 3; ++r) {
      u[r] = S[r * 3 + 0] * Vs[0 * 3 + c] + S[r * 3 + 1] * Vs[1 * 3 + c] + S[r * 3 + 2] * Vs[2 * 3 + c];
      nrm += u[r] * u[r];
    }
    nrm = std::sqrt(nrm);
    for (int r = 0; r < 3; ++r) U[r * 3 + c] = nrm > 0 ? u[r] / nrm : (r == c ? 1.0 : 0.0);
  }
  {  // re-orthogonalise u2 against u1, then u3 = u1 x u2
    double dot = U[0] * U[1] + U[3] * U[4] + U[6] * U[7];
    double nrm = 0;
    for (int r = 0; r < 3; ++r) {
      U[r * 3 + 1] -= dot * U[r * 3 + 0];
      nrm += U[r * 3 + 1] * U[r * 3 + 1];
    }
    nrm = std::sqrt(nrm);
    if (nrm > 0)
      for (int r = 0; r < 3; ++r) U[r * 3 + 1] /= nrm;
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
  double dv = det3(Vs) < 0 ? -1.0 : 1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R[i * 3 + j] =
          U[i * 3 + 0] * Vs[j * 3 + 0] + U[i * 3 + 1] * Vs[j * 3 + 1] + dv * U[i * 3 + 2] * Vs[j * 3 + 2];
}

// x = A^-1 b for a dense 6x6 (Gaussian elimination, partial pivoting), double.
bool solve6(double A[36], double b[6], double x[6]) {
  int n = 6;
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
    if (A[piv * n + c] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < n; ++k) std::swap(A[c * n + k], A[piv * n + k]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < n; ++r) {
      double f = A[r * n + c] / A[c * n + c];
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= A[r * n + k] * x[k];
    x[r] = s / A[r * n + r];
  }
  return true;
}

struct Corr {
  int32_t q, m;
  float d2;
};

// TransformationEstimationSVD::estimateRigidTransformation via Umeyama.
// f32 = PCL-like float32 sums; otherwise double accumulation.
M4 estimate_svd(const std::vector<float>& X, const lc3d_cloud* tgt, const std::vector<Corr>& corr,
                bool f32) {
  const int64_t n = (int64_t)corr.size();
  double mu_s[3], mu_d[3], S[9];
  if (f32) {
    float ss[3] = {0, 0, 0}, sd[3] = {0, 0, 0};
    for (const Corr& c : corr) {
      const float* s = &X[3 * (size_t)c.q];
      const float* d = xyz_at(tgt, c.m);
      for (int k = 0; k < 3; ++k) {
        ss[k] += s[k];
        sd[k] += d[k];
      }
    }
    float inv = 1.0f / (float)n;
    float ms[3], md[3];
    for (int k = 0; k < 3; ++k) {
      ms[k] = ss[k] * inv;
      md[k] = sd[k] * inv;
    }
    float acc[9] = {0};
    for (const Corr& c : corr) {
      const float* s = &X[3 * (size_t)c.q];
      const float* d = xyz_at(tgt, c.m);
      float sc[3] = {s[0] - ms[0], s[1] - ms[1], s[2] - ms[2]};
      float dc[3] = {d[0] - md[0], d[1] - md[1], d[2] - md[2]};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) acc[i * 3 + j] += dc[i] * sc[j];
    }
    for (int i = 0; i < 9; ++i) S[i] = (double)(acc[i] * inv);
    for (int k = 0; k < 3; ++k) {
      mu_s[k] = ms[k];
      mu_d[k] = md[k];
    }
  } else {
    double ss[3] = {0, 0, 0}, sd[3] = {0, 0, 0}, acc[9] = {0};
    for (const Corr& c : corr) {
      const float* s = &X[3 * (size_t)c.q];
      const float* d = xyz_at(tgt, c.m);
      for (int k = 0; k < 3; ++k) {
        ss[k] += s[k];
        sd[k] += d[k];
      }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) acc[i * 3 + j] += (double)d[i] * (double)s[j];
    }
    for (int k = 0; k < 3; ++k) {
      mu_s[k] = ss[k] / (double)n;
      mu_d[k] = sd[k] / (double)n;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) S[i * 3 + j] = acc[i * 3 + j] / (double)n - mu_d[i] * mu_s[j];
  }
  double R[9];
  kabsch_rotation(S, R);
  M4 T = m4_identity();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.m[i * 4 + j] = (float)R[i * 3 + j];
    double t = mu_d[i] - (R[i * 3 + 0] * mu_s[0] + R[i * 3 + 1] * mu_s[1] + R[i * 3 + 2] * mu_s[2]);
    T.m[i * 4 + 3] = (float)t;
  }
  return T;
}

// TransformationEstimationPointToPlaneLLS: float32 products widened to
// double, 6x6 normal equations in double, Euler angles -> R = Rz(g) Ry(b) Rx(a).
M4 estimate_p2plane(const std::vector<float>& X, const lc3d_cloud* tgt, const std::vector<Corr>& corr,
                    bool* ok) {
  double ATA[36] = {0}, ATb[6] = {0};
  for (const Corr& c : corr) {
    const float* s = &X[3 * (size_t)c.q];
    const float* d = xyz_at(tgt, c.m);
    const float* nn = nrm_at(tgt, c.m);
    float sx = s[0], sy = s[1], sz = s[2], dx = d[0], dy = d[1], dz = d[2];
    float nx = nn[0], ny = nn[1], nz = nn[2];
    if (!finite3(s) || !finite3(d) || !finite3(nn)) continue;
    double J[6];
    J[0] = (double)(nz * sy - ny * sz);
    J[1] = (double)(nx * sz - nz * sx);
    J[2] = (double)(ny * sx - nx * sy);
    J[3] = nx;
    J[4] = ny;
    J[5] = nz;
    double r = (double)(nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz);
    for (int i = 0; i < 6; ++i) {
      for (int j = i; j < 6; ++j) ATA[i * 6 + j] += J[i] * J[j];
      ATb[i] += J[i] * r;
    }
  }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) ATA[i * 6 + j] = ATA[j * 6 + i];
  double x[6];
  *ok = solve6(ATA, ATb, x);
  M4 T = m4_identity();
  if (!*ok) return T;
  double al = x[0], be = x[1], ga = x[2];
  T.m[0] = (float)(std::cos(ga) * std::cos(be));
  T.m[1] = (float)(-std::sin(ga) * std::cos(al) + std::cos(ga) * std::sin(be) * std::sin(al));
  T.m[2] = (float)(std::sin(ga) * std::sin(al) + std::cos(ga) * std::sin(be) * std::cos(al));
  T.m[4] = (float)(std::sin(ga) * std::cos(be));
  T.m[5] = (float)(std::cos(ga) * std::cos(al) + std::sin(ga) * std::sin(be) * std::sin(al));
  T.m[6] = (float)(-std::cos(ga) * std::sin(al) + std::sin(ga) * std::sin(be) * std::cos(al));
  T.m[8] = (float)(-std::sin(be));
  T.m[9] = (float)(std::cos(be) * std::sin(al));
  T.m[10] = (float)(std::cos(be) * std::cos(al));
  T.m[3] = (float)x[3];
  T.m[7] = (float)x[4];
  T.m[11] = (float)x[5];
  return T;
}

// ------------------------------------------------------- eigen33 (float32) ---
// pcl::eigen33 smallest eigenpair in closed form, float32.
void roots2(float b, float c, float r[3]) {
  r[0] = 0.0f;
  float d = (float)(b * b - 4.0 * c);
  if (d < 0.0f) d = 0.0f;
  float sd = std::sqrt(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}
void roots3(const float m[9], float r[3]) {
  float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
             m[8] * m[1] * m[1];
  float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  float c2 = m[0] + m[4] + m[8];
  if (std::fabs(c0) < FLT_EPSILON) {
    roots2(c2, c1, r);
    return;
  }
  const float inv3 = (float)(1.0 / 3.0);
  const float sqrt3 = std::sqrt(3.0f);
  float c2_3 = c2 * inv3;
  float a_3 = (c1 - c2 * c2_3) * inv3;
  if (a_3 > 0.0f) a_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
  float q = half_b * half_b + a_3 * a_3 * a_3;
  if (q > 0.0f) q = 0.0f;
  float rho = std::sqrt(-a_3);
  float theta = std::atan2(std::sqrt(-q), half_b) * inv3;
  float ct = std::cos(theta), st = std::sin(theta);
  r[0] = c2_3 + 2.0f * rho * ct;
  r[1] = c2_3 - rho * (ct + sqrt3 * st);
  r[2] = c2_3 - rho * (ct - sqrt3 * st);
  if (r[0] >= r[1]) std::swap(r[0], r[1]);
  if (r[1] >= r[2]) {
    std::swap(r[1], r[2]);
    if (r[0] >= r[1]) std::swap(r[0], r[1]);
  }
  if (r[0] <= 0.0f) roots2(c2, c1, r);
}
inline void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
void eigen33_smallest(const float C[9], float* eigenvalue, float v[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(C[i]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float m[9];
  for (int i = 0; i < 9; ++i) m[i] = C[i] / scale;
  float r[3];
  roots3(m, r);
  *eigenvalue = r[0] * scale;
  m[0] -= r[0];
  m[4] -= r[0];
  m[8] -= r[0];
  float v1[3], v2[3], v3[3];
  cross3(&m[0], &m[3], v1);
  cross3(&m[0], &m[6], v2);
  cross3(&m[3], &m[6], v3);
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* best;
  float len;
  if (l1 >= l2 && l1 >= l3) {
    best = v1;
    len = l1;
  } else if (l2 >= l1 && l2 >= l3) {
    best = v2;
    len = l2;
  } else {
    best = v3;
    len = l3;
  }
  float s = std::sqrt(len);
  for (int k = 0; k < 3; ++k) v[k] = best[k] / s;
}

}  // namespace

// ============================================================= C interface ===
extern "C" {

struct orc_kdtree {
  KdTree t;
};

orc_kdtree* orc_kdtree_build(const lc3d_cloud* c) {
  orc_kdtree* k = new orc_kdtree;
  k->t.build(c);
  return k;
}
void orc_kdtree_free(orc_kdtree* k) { delete k; }

// Batch k-NN.  out arrays are m x k; missing neighbours are (-1, +inf).
void orc_kdtree_knn(const orc_kdtree* k, const lc3d_cloud* q, int32_t kk, int32_t* out_idx,
                    float* out_d2) {
  for (int64_t i = 0; i < q->n; ++i) {
    int32_t* oi = out_idx + i * kk;
    float* od = out_d2 + i * kk;
    int cnt = finite3(xyz_at(q, i)) ? k->t.knn(xyz_at(q, i), kk, oi, od) : 0;
    for (int j = cnt; j < kk; ++j) {
      oi[j] = -1;
      od[j] = std::numeric_limits<float>::infinity();
    }
  }
}

// 1-NN with PCL's correspondence gate: rejected iff (double)d2 > max_dist^2
// (CorrespondenceEstimation::determineCorrespondences).  max_dist<=0: no gate.
void orc_kdtree_nn(const orc_kdtree* k, const lc3d_cloud* q, double max_dist, int32_t* out_idx,
                   float* out_d2) {
  const double gate = max_dist > 0 ? max_dist * max_dist : std::numeric_limits<double>::infinity();
  for (int64_t i = 0; i < q->n; ++i) {
    int32_t idx = -1;
    float d2 = std::numeric_limits<float>::infinity();
    int cnt = finite3(xyz_at(q, i)) ? k->t.knn(xyz_at(q, i), 1, &idx, &d2) : 0;
    if (cnt == 0 || (double)d2 > gate) {
      idx = -1;
      d2 = std::numeric_limits<float>::infinity();
    }
    out_idx[i] = idx;
    out_d2[i] = d2;
  }
}

// pcl::IterativeClosestPoint::align + getFitnessScore as driven by
// pcl_tools/fine_registration.cpp:105-126.
// umeyama_f32 != 0: PCL-like float32 estimator sums; 0: double accumulation.
// iteration_log (or NULL): per-iteration {n_corr, mse} pairs, 2*max_iterations doubles.
int orc_icp_align(const lc3d_cloud* src, const lc3d_cloud* tgt, const lc3d_icp_params* p,
                  int32_t umeyama_f32, lc3d_icp_result* res, const lc3d_icp_outputs* out,
                  double* iteration_log) {
  std::memset(res, 0, sizeof *res);
  const int64_t n = src->n;
  KdTree tree;
  tree.build(tgt);
  const bool with_normals = src->normal != nullptr;
  std::vector<float> X(3 * (size_t)n), N;
  for (int64_t i = 0; i < n; ++i) std::memcpy(&X[3 * i], xyz_at(src, i), 12);
  if (with_normals) {
    N.resize(3 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) std::memcpy(&N[3 * i], nrm_at(src, i), 12);
  }
  M4 T = m4_identity(), Tfinal = m4_identity();
  const double max2 = p->max_correspondence_distance * p->max_correspondence_distance;
  // DefaultConvergenceCriteria as configured by ICP::computeTransformation (A.2/A.3)
  const double rot_thr = 1.0 - p->transformation_epsilon;
  const double transl_thr = p->transformation_epsilon;
  const double rel_mse = p->euclidean_fitness_epsilon;
  const double abs_mse = 1e-12;
  double prev_mse = std::numeric_limits<double>::max();
  int iterations = 0, state = LC3D_STATE_NOT_CONVERGED;
  bool converged = false;
  std::vector<Corr> corr;
  corr.reserve(n);
  do {
    corr.clear();
    for (int64_t i = 0; i < n; ++i) {
      const float* q = &X[3 * i];
      if (!finite3(q)) continue;
      int32_t idx;
      float d2;
      if (tree.knn(q, 1, &idx, &d2) == 0) continue;
      if ((double)d2 > max2) continue;
      corr.push_back({(int32_t)i, idx, d2});
    }
    if (out && p->dump_iteration == iterations) {
      if (out->corr_index)
        for (int64_t i = 0; i < n; ++i) out->corr_index[i] = -1;
      if (out->corr_dist2)
        for (int64_t i = 0; i < n; ++i) out->corr_dist2[i] = std::numeric_limits<float>::infinity();
      for (const Corr& c : corr) {
        if (out->corr_index) out->corr_index[c.q] = c.m;
        if (out->corr_dist2) out->corr_dist2[c.q] = c.d2;
      }
    }
    double sum = 0;
    for (const Corr& c : corr) sum += (double)c.d2;
    res->last_correspondences = (int64_t)corr.size();
    res->last_mse = corr.empty() ? 0.0 : sum / (double)corr.size();
    if (iteration_log) {
      iteration_log[2 * iterations] = (double)corr.size();
      iteration_log[2 * iterations + 1] = res->last_mse;
    }
    if (corr.size() < 3) {  // "Not enough correspondences found"
      state = LC3D_STATE_NO_CORRESPONDENCES;
      converged = false;
      break;
    }
    if (p->mode == LC3D_ICP_POINT_TO_PLANE) {
      bool ok;
      T = estimate_p2plane(X, tgt, corr, &ok);
    } else {
      T = estimate_svd(X, tgt, corr, umeyama_f32 != 0);
    }
    // transformCloud: incremental, in place, float32 (A.2)
    for (int64_t i = 0; i < n; ++i) {
      float* q = &X[3 * i];
      if (!finite3(q)) continue;
      float o[3];
      xform_point(T, q, o);
      q[0] = o[0];
      q[1] = o[1];
      q[2] = o[2];
      if (with_normals) {
        float* nn = &N[3 * i];
        if (!finite3(nn)) continue;
        xform_normal(T, nn, o);
        nn[0] = o[0];
        nn[1] = o[1];
        nn[2] = o[2];
      }
    }
    Tfinal = m4_mul(T, Tfinal);
    ++iterations;
    // DefaultConvergenceCriteria::hasConverged (A.3)
    state = LC3D_STATE_NOT_CONVERGED;
    if (iterations >= p->max_iterations) {
      state = LC3D_STATE_ITERATIONS;
      converged = true;
    } else {
      double cos_angle = 0.5 * (double)(T.m[0] + T.m[5] + T.m[10] - 1.0f);
      double tsq = (double)(T.m[3] * T.m[3]) + (double)(T.m[7] * T.m[7]) + (double)(T.m[11] * T.m[11]);
      if (cos_angle >= rot_thr && tsq <= transl_thr) {
        state = LC3D_STATE_TRANSFORM;
        converged = true;
      } else {
        double mse = res->last_mse;
        if (std::fabs(mse - prev_mse) < abs_mse) {
          state = LC3D_STATE_ABS_MSE;
          converged = true;
        } else if (std::fabs(mse - prev_mse) / prev_mse < rel_mse) {
          state = LC3D_STATE_REL_MSE;
          converged = true;
        } else {
          prev_mse = mse;
        }
      }
    }
  } while (!converged);

  std::memcpy(res->transformation, Tfinal.m, sizeof Tfinal.m);
  res->converged = converged ? 1 : 0;
  res->iterations = iterations;
  res->state = state;
  // output = transformCloud(*input_, final_transformation_) — one application (A.2)
  std::vector<float> Y;
  const bool need_Y = p->compute_fitness || (out && out->registered_xyz);
  if (need_Y) {
    Y.resize(3 * (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      const float* q = xyz_at(src, i);
      if (finite3(q))
        xform_point(Tfinal, q, &Y[3 * i]);
      else
        std::memcpy(&Y[3 * i], q, 12);
    }
  }
  if (out && out->registered_xyz) std::memcpy(out->registered_xyz, Y.data(), Y.size() * 4);
  if (out && out->registered_normal && with_normals) {
    for (int64_t i = 0; i < n; ++i) {
      const float* nn = nrm_at(src, i);
      if (finite3(nn))
        xform_normal(Tfinal, nn, &out->registered_normal[3 * i]);
      else
        std::memcpy(&out->registered_normal[3 * i], nn, 12);
    }
  }
  // getFitnessScore(max_range = DBL_MAX) (A.4): all source points, unbounded NN.
  if (p->compute_fitness) {
    double fs = 0;
    int64_t nr = 0;
    for (int64_t i = 0; i < n; ++i) {
      const float* q = &Y[3 * i];
      if (!finite3(q)) continue;
      int32_t idx;
      float d2;
      if (tree.knn(q, 1, &idx, &d2) == 0) continue;
      fs += (double)d2;
      ++nr;
    }
    res->fitness = nr > 0 ? fs / (double)nr : std::numeric_limits<double>::max();
  }
  return 0;
}

// One ICP-shaped iteration for timing the CPU baseline: kd-tree 1-NN over all source
// points + estimator, no loop.  Returns the number of correspondences.
int64_t orc_icp_one_iteration(const orc_kdtree* k, const lc3d_cloud* src, const lc3d_cloud* tgt,
                              double max_dist, int32_t mode, float out_T[16]) {
  const int64_t n = src->n;
  std::vector<float> X(3 * (size_t)n);
  for (int64_t i = 0; i < n; ++i) std::memcpy(&X[3 * i], xyz_at(src, i), 12);
  std::vector<Corr> corr;
  corr.reserve(n);
  const double max2 = max_dist * max_dist;
  for (int64_t i = 0; i < n; ++i) {
    int32_t idx;
    float d2;
    if (!finite3(&X[3 * i]) || k->t.knn(&X[3 * i], 1, &idx, &d2) == 0) continue;
    if ((double)d2 > max2) continue;
    corr.push_back({(int32_t)i, idx, d2});
  }
  M4 T = m4_identity();
  if (corr.size() >= 3) {
    bool ok;
    T = mode == LC3D_ICP_POINT_TO_PLANE ? estimate_p2plane(X, tgt, corr, &ok)
                                        : estimate_svd(X, tgt, corr, true);
    for (int64_t i = 0; i < n; ++i) {
      float o[3];
      xform_point(T, &X[3 * i], o);
      std::memcpy(&X[3 * i], o, 12);
    }
  }
  std::memcpy(out_T, T.m, sizeof T.m);
  return (int64_t)corr.size();
}

// pcl::compute3DCentroid (normal_estimation.cpp:101): float32 sequential sum / n.
void orc_centroid(const lc3d_cloud* c, float out[4]) {
  float s[3] = {0, 0, 0};
  int64_t cnt = 0;
  for (int64_t i = 0; i < c->n; ++i) {
    const float* p = xyz_at(c, i);
    if (!finite3(p)) continue;
    s[0] += p[0];
    s[1] += p[1];
    s[2] += p[2];
    ++cnt;
  }
  float fc = (float)cnt;
  out[0] = s[0] / fc;
  out[1] = s[1] / fc;
  out[2] = s[2] / fc;
  out[3] = 1.0f;
}

// pcl::NormalEstimation::compute as driven by normal_estimation.cpp:84-108.
// knn_idx (or NULL): precomputed n x k neighbour table to use instead of the kd-tree
// (lets tests feed both implementations the same neighbour sets).
int orc_normals(const lc3d_cloud* c, int32_t k, const float vp[3], const int32_t* knn_idx,
                float* out_normal, float* out_curv) {
  KdTree tree;
  if (!knn_idx) tree.build(c);
  std::vector<int32_t> idx(k);
  std::vector<float> d2(k);
  const float nanf_ = std::numeric_limits<float>::quiet_NaN();
  for (int64_t i = 0; i < c->n; ++i) {
    const float* p = xyz_at(c, i);
    float* on = out_normal + 3 * i;
    int cnt = 0;
    if (finite3(p)) {
      if (knn_idx) {
        for (int j = 0; j < k; ++j)
          if (knn_idx[i * k + j] >= 0) idx[cnt++] = knn_idx[i * k + j];
      } else {
        cnt = tree.knn(p, k, idx.data(), d2.data());
      }
    }
    if (cnt < 3) {
      on[0] = on[1] = on[2] = nanf_;
      out_curv[i] = nanf_;
      continue;
    }
    float a[9] = {0};
    for (int j = 0; j < cnt; ++j) {
      const float* s = xyz_at(c, idx[j]);
      a[0] += s[0] * s[0];
      a[1] += s[0] * s[1];
      a[2] += s[0] * s[2];
      a[3] += s[1] * s[1];
      a[4] += s[1] * s[2];
      a[5] += s[2] * s[2];
      a[6] += s[0];
      a[7] += s[1];
      a[8] += s[2];
    }
    float fn = (float)cnt;
    for (int j = 0; j < 9; ++j) a[j] /= fn;
    float C[9];
    C[0] = a[0] - a[6] * a[6];
    C[1] = a[1] - a[6] * a[7];
    C[2] = a[2] - a[6] * a[8];
    C[4] = a[3] - a[7] * a[7];
    C[5] = a[4] - a[7] * a[8];
    C[8] = a[5] - a[8] * a[8];
    C[3] = C[1];
    C[6] = C[2];
    C[7] = C[5];
    float ev, v[3];
    eigen33_smallest(C, &ev, v);
    float tr = C[0] + C[4] + C[8];
    out_curv[i] = tr != 0.0f ? std::fabs(ev / tr) : 0.0f;
    // flipNormalTowardsViewpoint
    float vx = vp[0] - p[0], vy = vp[1] - p[1], vz = vp[2] - p[2];
    float cs = vx * v[0] + vy * v[1] + vz * v[2];
    if (cs < 0) {
      v[0] *= -1;
      v[1] *= -1;
      v[2] *= -1;
    }
    on[0] = v[0];
    on[1] = v[1];
    on[2] = v[2];
  }
  return 0;
}

// pcl::StatisticalOutlierRemoval::applyFilterIndices as driven by
// outlier_removal.cpp:80-84 / :91-93.
int orc_sor(const lc3d_cloud* c, int32_t mean_k, double std_mul, int32_t negative,
            int32_t* out_kept, int64_t* out_count, float* out_mean_dist, double out_stats[3]) {
  KdTree tree;
  tree.build(c);
  const int64_t n = c->n;
  std::vector<float> dist(n, 0.0f);
  std::vector<char> valid(n, 0);
  std::vector<int32_t> idx(mean_k + 1);
  std::vector<float> d2(mean_k + 1);
  int64_t nvalid = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = xyz_at(c, i);
    if (!finite3(p)) continue;
    int cnt = tree.knn(p, mean_k + 1, idx.data(), d2.data());
    if (cnt == 0) continue;
    double s = 0;
    for (int j = 1; j < cnt; ++j) s += std::sqrt((double)d2[j]);
    dist[i] = (float)(s / (double)mean_k);
    valid[i] = 1;
    ++nvalid;
  }
  double sum = 0, sq = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (!valid[i]) continue;  // PCL adds the zeros of invalid points too; they contribute 0
    sum += (double)dist[i];
    sq += (double)(dist[i] * dist[i]);
  }
  double mean = sum / (double)nvalid;
  double var = (sq - sum * sum / (double)nvalid) / ((double)nvalid - 1.0);
  double sd = std::sqrt(var);
  double thr = mean + std_mul * sd;
  int64_t cnt = 0;
  for (int64_t i = 0; i < n; ++i) {
    // float vs double compare, as PCL; invalid points carry distance 0 (=> inliers)
    bool outlier = (double)dist[i] > thr;
    bool keep = negative ? outlier : !outlier;
    if (keep) out_kept[cnt++] = (int32_t)i;
  }
  *out_count = cnt;
  if (out_mean_dist) std::memcpy(out_mean_dist, dist.data(), n * 4);
  if (out_stats) {
    out_stats[0] = mean;
    out_stats[1] = sd;
    out_stats[2] = thr;
  }
  return 0;
}

// pcl::VoxelGrid::applyFilter as driven by cloud_downsampling.cpp:73-76.
// The reference uses a non-stable std::sort on the voxel index, so the within-voxel
// summation order is unspecified there; this restatement (and the CUDA path) use a
// stable sort, i.e. ascending input index inside a voxel.
int orc_voxel_grid(const lc3d_cloud* c, const float leaf[3], float* out_xyz, float* out_normal,
                   uint32_t* out_rgba, float* out_curv, int32_t* out_voxel_of_point,
                   int64_t* out_count) {
  const int64_t n = c->n;
  float inv[3] = {1.0f / leaf[0], 1.0f / leaf[1], 1.0f / leaf[2]};
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int64_t i = 0; i < n; ++i) {
    const float* p = xyz_at(c, i);
    if (!finite3(p)) continue;
    for (int d = 0; d < 3; ++d) {
      mn[d] = std::min(mn[d], p[d]);
      mx[d] = std::max(mx[d], p[d]);
    }
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv[0]) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv[1]) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv[2]) + 1;
  auto copy_through = [&]() {
    for (int64_t i = 0; i < n; ++i) {
      std::memcpy(out_xyz + 3 * i, xyz_at(c, i), 12);
      if (out_normal && c->normal) std::memcpy(out_normal + 3 * i, nrm_at(c, i), 12);
      if (out_rgba && c->rgba) out_rgba[i] = rgba_at(c, i);
      if (out_curv && c->curvature) out_curv[i] = curv_at(c, i);
      if (out_voxel_of_point) out_voxel_of_point[i] = (int32_t)i;
    }
    *out_count = n;
  };
  if (n == 0) {
    *out_count = 0;
    return 0;
  }
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
    copy_through();  // PCL: "Leaf size is too small ... Integer indices would overflow."
    return 1;
  }
  int minb[3], maxb[3], divb[3], mul[3];
  for (int d = 0; d < 3; ++d) {
    minb[d] = (int)std::floor(mn[d] * inv[d]);
    maxb[d] = (int)std::floor(mx[d] * inv[d]);
    divb[d] = maxb[d] - minb[d] + 1;
  }
  mul[0] = 1;
  mul[1] = divb[0];
  mul[2] = divb[0] * divb[1];
  struct KI {
    uint32_t key;
    int32_t idx;
  };
  std::vector<KI> v;
  v.reserve(n);
  for (int64_t i = 0; i < n; ++i) {
    const float* p = xyz_at(c, i);
    if (!finite3(p)) {
      if (out_voxel_of_point) out_voxel_of_point[i] = -1;
      continue;
    }
    int i0 = (int)(std::floor(p[0] * inv[0]) - (float)minb[0]);
    int i1 = (int)(std::floor(p[1] * inv[1]) - (float)minb[1]);
    int i2 = (int)(std::floor(p[2] * inv[2]) - (float)minb[2]);
    int idx = i0 * mul[0] + i1 * mul[1] + i2 * mul[2];
    v.push_back({(uint32_t)idx, (int32_t)i});
  }
  std::stable_sort(v.begin(), v.end(), [](const KI& a, const KI& b) { return a.key < b.key; });
  int64_t m = 0;
  size_t i = 0;
  while (i < v.size()) {
    size_t j = i;
    float sx = 0, sy = 0, sz = 0, snx = 0, sny = 0, snz = 0, scurv = 0, sr = 0, sg = 0, sb = 0, sa = 0;
    while (j < v.size() && v[j].key == v[i].key) {
      int32_t pi = v[j].idx;
      const float* p = xyz_at(c, pi);
      sx += p[0];
      sy += p[1];
      sz += p[2];
      if (c->normal) {
        const float* nn = nrm_at(c, pi);
        snx += nn[0];
        sny += nn[1];
        snz += nn[2];
      }
      if (c->curvature) scurv += curv_at(c, pi);
      if (c->rgba) {
        uint32_t rgba = rgba_at(c, pi);
        sr += (float)((rgba >> 16) & 0xff);
        sg += (float)((rgba >> 8) & 0xff);
        sb += (float)(rgba & 0xff);
        sa += (float)((rgba >> 24) & 0xff);
      }
      if (out_voxel_of_point) out_voxel_of_point[pi] = (int32_t)m;
      ++j;
    }
    float fn = (float)(j - i);
    out_xyz[3 * m + 0] = sx / fn;
    out_xyz[3 * m + 1] = sy / fn;
    out_xyz[3 * m + 2] = sz / fn;
    if (out_normal && c->normal) {  // CentroidPoint: summed normal, normalised (1.8+)
      float nrm2 = snx * snx + sny * sny + snz * snz;
      if (nrm2 > 0.0f) {
        float nrm = std::sqrt(nrm2);
        snx /= nrm;
        sny /= nrm;
        snz /= nrm;
      }
      out_normal[3 * m + 0] = snx;
      out_normal[3 * m + 1] = sny;
      out_normal[3 * m + 2] = snz;
    }
    if (out_curv && c->curvature) out_curv[m] = scurv / fn;
    if (out_rgba && c->rgba)
      out_rgba[m] = ((uint32_t)(sa / fn) << 24) | ((uint32_t)(sr / fn) << 16) |
                    ((uint32_t)(sg / fn) << 8) | (uint32_t)(sb / fn);
    ++m;
    i = j;
  }
  *out_count = m;
  return 0;
}

// pcl::transformPointCloudWithNormals (pcl_tools/transform.cpp:84-90).
void orc_transform(const lc3d_cloud* c, const float T16[16], float* out_xyz, float* out_normal) {
  M4 T;
  std::memcpy(T.m, T16, sizeof T.m);
  for (int64_t i = 0; i < c->n; ++i) {
    const float* p = xyz_at(c, i);
    if (finite3(p))
      xform_point(T, p, out_xyz + 3 * i);
    else
      std::memcpy(out_xyz + 3 * i, p, 12);
    if (out_normal && c->normal) {
      const float* nn = nrm_at(c, i);
      if (finite3(nn))
        xform_normal(T, nn, out_normal + 3 * i);
      else
        std::memcpy(out_normal + 3 * i, nn, 12);
    }
  }
}

// accumulate_clouds.cpp:100-111: for every target point t, pcl::CropBox (negative) drops the
// source points inside the axis-aligned box [t - r, t + r]; the box corners are computed as
// (float)((double)t.x -/+ r) and the test is min <= p <= max per component (inclusive), so a
// source point survives iff NO target point has it inside its box.  Non-finite source points
// are dropped by CropBox.  Literal O(N*M) restatement (the order of the survivors is the
// source order, as CropBox preserves it).
int orc_box_dedup(const lc3d_cloud* src, const lc3d_cloud* tgt, double radius, int32_t* out_kept,
                  int64_t* out_count) {
  std::vector<char> alive(src->n, 1);
  for (int64_t i = 0; i < src->n; ++i)
    if (!finite3(xyz_at(src, i))) alive[i] = 0;
  for (int64_t j = 0; j < tgt->n; ++j) {
    const float* t = xyz_at(tgt, j);
    const float lo[3] = {(float)((double)t[0] - radius), (float)((double)t[1] - radius), (float)((double)t[2] - radius)};
    const float hi[3] = {(float)((double)t[0] + radius), (float)((double)t[1] + radius), (float)((double)t[2] + radius)};
    for (int64_t i = 0; i < src->n; ++i) {
      if (!alive[i]) continue;
      const float* p = xyz_at(src, i);
      const bool outside = p[0] < lo[0] || p[1] < lo[1] || p[2] < lo[2] || p[0] > hi[0] || p[1] > hi[1] || p[2] > hi[2];
      if (!outside) alive[i] = 0;
    }
  }
  int64_t c = 0;
  for (int64_t i = 0; i < src->n; ++i)
    if (alive[i]) out_kept[c++] = (int32_t)i;
  *out_count = c;
  return 0;
}

// pcl::EuclideanClusterExtraction::extract (pcl_tools/cluster_extraction.cpp:94-101; PCL 1.8.1
// segmentation/impl/extract_clusters.hpp extractEuclideanClusters): breadth-first growth from
// every unprocessed point over radiusSearch(point, tolerance) neighbours, a queue is kept as a
// cluster iff min_size <= size <= max_size, each cluster's indices sorted ascending, clusters
// sorted by size descending (std::sort on reverse iterators: ties unspecified in PCL — here
// ties go to the cluster holding the lower point index).  Non-finite points are never part
// of a cluster (PCL asserts on them).  labels[i] = rank of the cluster of point i, or -1.
// sizes[0..min(count,cap)) = cluster sizes by rank.
int orc_euclidean_clusters(const lc3d_cloud* cloud, double tolerance, int64_t min_size, int64_t max_size,
                           int32_t* labels, int64_t* sizes, int64_t cap, int64_t* out_count) {
  const int64_t n = cloud->n;
  KdTree tree;
  tree.build(cloud);
  const float r2 = (float)(tolerance * tolerance);
  std::vector<char> processed(n, 0);
  std::vector<std::vector<int32_t>> clusters;
  std::vector<int32_t> nn, queue;
  for (int64_t i = 0; i < n; ++i) labels[i] = -1;
  for (int64_t i = 0; i < n; ++i) {
    if (processed[i]) continue;
    processed[i] = 1;
    if (!finite3(xyz_at(cloud, i))) continue;
    queue.clear();
    queue.push_back((int32_t)i);
    for (size_t h = 0; h < queue.size(); ++h) {
      tree.radius(xyz_at(cloud, queue[h]), r2, nn);
      for (int32_t j : nn) {
        if (processed[j]) continue;
        processed[j] = 1;
        queue.push_back(j);
      }
    }
    if ((int64_t)queue.size() >= min_size && (int64_t)queue.size() <= max_size) {
      std::sort(queue.begin(), queue.end());
      clusters.push_back(queue);
    }
  }
  std::sort(clusters.begin(), clusters.end(), [](const std::vector<int32_t>& a, const std::vector<int32_t>& b) {
    if (a.size() != b.size()) return a.size() > b.size();
    return a[0] < b[0];
  });
  for (size_t c = 0; c < clusters.size(); ++c) {
    for (int32_t j : clusters[c]) labels[j] = (int32_t)c;
    if ((int64_t)c < cap && sizes) sizes[c] = (int64_t)clusters[c].size();
  }
  *out_count = (int64_t)clusters.size();
  return 0;
}

// Exposed for tests of the oracle's own building blocks.
void orc_kabsch_rotation(const double S[9], double R[9]) { kabsch_rotation(S, R); }
void orc_eigen33(const float C[9], float* eigenvalue, float v[3]) { eigen33_smallest(C, eigenvalue, v); }

}  // extern "C"
