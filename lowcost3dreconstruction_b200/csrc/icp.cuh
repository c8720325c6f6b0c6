// icp.cuh — the ICP loop of pcl::IterativeClosestPoint::align as driven by
// pcl_tools/fine_registration.cpp:105-126, entirely on the device.
//
// Two kernels per iteration:
//   icp_iteration_kernel fuses
//     transformCloud (incremental float32 transform of the source, applied at the head of
//       the NEXT iteration because T_k is only known after the grid-wide reduce),
//     determineCorrespondences (exact 1-NN + max-distance gate, seeded by the previous match),
//     the estimator's sums (3x3 cross-covariance for Umeyama/SVD, or the 6x6 normal equations
//       of point-to-plane LLS) in fp64: one butterfly reduce-scatter per warp, one partial row
//       per warp, no block epilogue;
//   icp_solve_kernel reduces the warp rows in a fixed order (deterministic), the last block
//     solves (3x3 one-sided Jacobi SVD / 6x6 elimination, in registers), composes
//     final_transformation_ and runs DefaultConvergenceCriteria.
// The chain is launched with programmatic dependent launch; the host enqueues iterations in
// chunks and polls the device-side `done` flag asynchronously (no per-iteration round trip);
// once `done` is set the remaining launches return immediately.
#pragma once
#include "search.cuh"

namespace lc3d {

#ifndef LC3D_ICP_THREADS
#define LC3D_ICP_THREADS 128
#endif
constexpr int kIcpThreads = LC3D_ICP_THREADS;
constexpr int kNvP2P = 17;     // sum s(3) sum d(3) sum d s^T(9) sum d2(1) count(1)
constexpr int kNvP2Plane = 29; // JtJ upper(21) Jtr(6) sum d2(1) count(1)

struct IcpState {
  float T[16];       // transformation_ of the last solve (row-major)
  float Tfinal[16];  // final_transformation_
  double prev_mse;   // correspondences_prev_mse_
  double last_mse;
  double fitness_sum;
  long long fitness_cnt;
  long long last_corr;
  int iter;       // nr_iterations_
  int done;       // loop finished
  int converged;  // converged_
  int state;      // LC3D_STATE_*
  unsigned ticket;
  unsigned ticket2;
  int pad[2];
};

// LC3D_STATS=1 + LC3D_BLOCK_LOG=<file>: per-block (start ns, end ns, SM id) of every iteration kernel
__device__ unsigned long long* g_block_log = nullptr;

struct IcpConfig {
  float gate;         // largest float <= max_correspondence_distance^2
  float gate_ext;     // (gate_dist + margin)^2: searches run against this extended gate
  float gate_dist;    // sqrt(gate), rounded up
  float margin_slack; // slack learnt when nothing lies within the extended gate
  float slack_floor;  // slacks at or below this are treated as unknown (float safety)
  int max_iterations;
  double rot_thr;     // 1 - transformation_epsilon
  double transl_thr;  // transformation_epsilon (squared translation)
  double rel_mse;     // euclidean_fitness_epsilon
  double abs_mse;     // 1e-12
  int dump_iteration;
  int mode;
  SearchStats* stats;  // per-iteration search statistics (LC3D_STATS=1) or null
};

__global__ void icp_state_init(IcpState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int i = 0; i < 16; ++i) st->T[i] = st->Tfinal[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    st->prev_mse = 1.7976931348623157e308;  // DBL_MAX
    st->last_mse = 0.0;
    st->fitness_sum = 0.0;
    st->fitness_cnt = 0;
    st->last_corr = 0;
    st->iter = 0;
    st->done = 0;
    st->converged = 0;
    st->state = LC3D_STATE_NOT_CONVERGED;
    st->ticket = 0;
    st->ticket2 = 0;
    st->pad[0] = st->pad[1] = 0;
  }
}

// ---- small dense solvers (fp64, single thread) --------------------------------------

// Rotation of the Umeyama / Kabsch problem for cross-covariance S (row-major 3x3):
// S = U D V^T, R = U diag(1,1,det(U)det(V)) V^T.  One-sided (Hestenes) Jacobi on the
// columns of S; the third left vector is u1 x u2, which folds det(U) into the product.
// Written with fixed trip counts and selects instead of index arrays so that A, V, U, W stay
// in registers (the solve is one thread on the iteration's critical path).
__device__ __forceinline__ double sel3(int o, double a0, double a1, double a2) {
  return o == 0 ? a0 : (o == 1 ? a1 : a2);
}
__device__ __forceinline__ void kabsch_rotation_dev(const double* S, double* R) {
  double A[9], V[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    A[i] = S[i];
    V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 40; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;  // (0,1) (0,2) (1,2)
      double al = 0, be = 0, ga = 0;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        al += A[k * 3 + p] * A[k * 3 + p];
        be += A[k * 3 + q] * A[k * 3 + q];
        ga += A[k * 3 + p] * A[k * 3 + q];
      }
      if (ga == 0.0 || fabs(ga) <= 1e-16 * sqrt(al * be)) continue;
      rotated = true;
      double zeta = (be - al) / (2.0 * ga);
      double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        double ap = A[k * 3 + p], aq = A[k * 3 + q];
        A[k * 3 + p] = c * ap - s * aq;
        A[k * 3 + q] = s * ap + c * aq;
        double vp = V[k * 3 + p], vq = V[k * 3 + q];
        V[k * 3 + p] = c * vp - s * vq;
        V[k * 3 + q] = s * vp + c * vq;
      }
    }
    if (!rotated) break;
  }
  double sg0 = A[0] * A[0] + A[3] * A[3] + A[6] * A[6];
  double sg1 = A[1] * A[1] + A[4] * A[4] + A[7] * A[7];
  double sg2 = A[2] * A[2] + A[5] * A[5] + A[8] * A[8];
  int o0 = 0, o1 = 1, o2 = 2;  // descending singular values
  if (sel3(o0, sg0, sg1, sg2) < sel3(o1, sg0, sg1, sg2)) { int t = o0; o0 = o1; o1 = t; }
  if (sel3(o1, sg0, sg1, sg2) < sel3(o2, sg0, sg1, sg2)) { int t = o1; o1 = o2; o2 = t; }
  if (sel3(o0, sg0, sg1, sg2) < sel3(o1, sg0, sg1, sg2)) { int t = o0; o0 = o1; o1 = t; }
  double U[9], W[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    W[r * 3 + 0] = sel3(o0, V[r * 3], V[r * 3 + 1], V[r * 3 + 2]);
    W[r * 3 + 1] = sel3(o1, V[r * 3], V[r * 3 + 1], V[r * 3 + 2]);
    W[r * 3 + 2] = sel3(o2, V[r * 3], V[r * 3 + 1], V[r * 3 + 2]);
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int oc = c == 0 ? o0 : o1;
    const double nrm = sqrt(sel3(oc, sg0, sg1, sg2));
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double a = sel3(oc, A[r * 3], A[r * 3 + 1], A[r * 3 + 2]);
      U[r * 3 + c] = nrm > 0 ? a / nrm : (r == c ? 1.0 : 0.0);
    }
  }
  {
    double dot = U[0] * U[1] + U[3] * U[4] + U[6] * U[7], nrm = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      U[r * 3 + 1] -= dot * U[r * 3 + 0];
      nrm += U[r * 3 + 1] * U[r * 3 + 1];
    }
    nrm = sqrt(nrm);
    if (nrm > 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r) U[r * 3 + 1] /= nrm;
    }
    U[2] = U[3] * U[7] - U[6] * U[4];
    U[5] = U[6] * U[1] - U[0] * U[7];
    U[8] = U[0] * U[4] - U[3] * U[1];
  }
  double detW = W[0] * (W[4] * W[8] - W[5] * W[7]) - W[1] * (W[3] * W[8] - W[5] * W[6]) +
                W[2] * (W[3] * W[7] - W[4] * W[6]);
  double dv = detW < 0 ? -1.0 : 1.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R[i * 3 + j] = U[i * 3 + 0] * W[j * 3 + 0] + U[i * 3 + 1] * W[j * 3 + 1] +
                     dv * U[i * 3 + 2] * W[j * 3 + 2];
}

// 6x6 Gaussian elimination with partial pivoting, fully unrolled so that the matrix lives in
// registers (run by ONE thread on the iteration's critical path; with dynamic row indices the
// arrays went to local memory and the solve took ~7 us).  Pivoting by conditional row swaps:
// the pivot row is the same as with a single max search (ties aside), the other rows only
// change places, and each row's elimination arithmetic does not depend on its position.
__device__ __forceinline__ bool solve6_dev(double (&A)[36], double (&b)[6], double (&x)[6]) {
  // all loops have fixed trip counts (guards instead of variable bounds) so that nvcc unrolls
  // them completely and every A[...] index is a compile-time constant
  bool ok = true;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r > c) {
        const bool sw = fabs(A[r * 6 + c]) > fabs(A[c * 6 + c]);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          if (k >= c) {
            const double t = A[c * 6 + k], u = A[r * 6 + k];
            A[c * 6 + k] = sw ? u : t;
            A[r * 6 + k] = sw ? t : u;
          }
        }
        const double t = b[c], u = b[r];
        b[c] = sw ? u : t;
        b[r] = sw ? t : u;
      }
    }
    ok = ok && A[c * 6 + c] != 0.0;
    const double piv = ok ? A[c * 6 + c] : 1.0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r > c) {
        const double f = A[r * 6 + c] / piv;
#pragma unroll
        for (int k = 0; k < 6; ++k)
          if (k >= c) A[r * 6 + k] -= f * A[c * 6 + k];
        b[r] -= f * b[c];
      }
    }
  }
  if (!ok) return false;
#pragma unroll
  for (int rr = 0; rr < 6; ++rr) {
    const int r = 5 - rr;
    double s = b[r];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k > r) s -= A[r * 6 + k] * x[k];
    x[r] = s / A[r * 6 + r];
  }
  return true;
}

// The same elimination spread over the lanes of one warp: lane r < 6 owns row r of [A | b].
// Per column the pivot is the unused row with the largest |entry| (butterfly arg-max over the
// lanes), its row is broadcast, every other unused row eliminates in parallel; the back
// substitution walks the pivots in reverse.  Each row sees exactly the operations of the
// single-thread version above (same pivots, same order), so the results are identical; the
// dependent chain shrinks from ~600 to ~100 fp64 operations.  All 32 lanes must call; v = the
// reduced sums (upper triangle packed, then the right-hand side).  Returns false if singular.
__device__ __forceinline__ bool solve6_warp(const double* __restrict__ v, double (&x)[6]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int r = lane < 6 ? lane : 0;
  double a[6], b;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int i = r < k ? r : k, j = r < k ? k : r;  // symmetric fill from the packed upper triangle
    a[k] = v[i * 6 - (i * (i - 1)) / 2 + (j - i)];
  }
  b = v[21 + r];
  bool used = lane >= 6;  // lanes beyond the matrix never compete for a pivot
  bool ok = true;
  int prow[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    // arg-max of |a[c]| over the unused rows; ties -> lower lane
    double best = used ? -1.0 : fabs(a[c]);
    int who = lane;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(full, best, o);
      const int ow = __shfl_xor_sync(full, who, o);
      if (ob > best || (ob == best && ow < who)) {
        best = ob;
        who = ow;
      }
    }
    who = __shfl_sync(full, who, 0);  // lanes 0..7 agree; take lane 0's view
    prow[c] = who;
    double piv[6], pb;
#pragma unroll
    for (int k = 0; k < 6; ++k) piv[k] = __shfl_sync(full, a[k], who);
    pb = __shfl_sync(full, b, who);
    ok = ok && piv[c] != 0.0;
    const double pc = ok ? piv[c] : 1.0;
    if (lane == who) used = true;
    if (!used) {
      const double f = a[c] / pc;
#pragma unroll
      for (int k = 0; k < 6; ++k)
        if (k >= c) a[k] -= f * piv[k];
      b -= f * pb;
    }
  }
#pragma unroll
  for (int cc = 0; cc < 6; ++cc) {
    const int c = 5 - cc;
    double s = b;
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (k > c) s -= a[k] * x[k];
    const double xc = s / a[c];
    x[c] = __shfl_sync(full, xc, prow[c]);
  }
  return ok;
}

// Estimator + pose composition + DefaultConvergenceCriteria; run by ONE
// WARP of the last block (all 32 lanes call).  The 6x6 solve is spread over the lanes, the three
// sincos run on three lanes, lane L < 16 composes entry L of final_transformation_; the scalar
// bookkeeping is evaluated redundantly and written by lane 0.  v = the NV reduced sums.
template <int MODE>
__device__ void icp_solve_and_test(IcpState* st, const IcpConfig& cfg, const double* v) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const double cnt = v[NV - 1];
  const double mse = cnt > 0 ? v[NV - 2] / cnt : 0.0;
  // everything read from the state up front (lane 0 overwrites it below)
  const int iter = st->iter + 1;
  const double prev_mse = st->prev_mse;
  const float Fin = lane < 16 ? st->Tfinal[lane] : 0.0f;
  if (cnt < 3.0) {  // "Not enough correspondences found"
    if (lane == 0) {
      st->last_corr = (long long)cnt;
      st->last_mse = mse;
      st->state = LC3D_STATE_NO_CORRESPONDENCES;
      st->converged = 0;
      st->done = 1;
    }
    return;
  }
  float T[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  if (MODE == LC3D_ICP_POINT_TO_PLANE) {
    double x[6];
    if (solve6_warp(v, x)) {
      // lane 0,1,2: sincos of alpha, beta, gamma
      double sn = 0.0, cs = 1.0;
      if (lane < 3) sincos(x[lane], &sn, &cs);
      const double sa = __shfl_sync(full, sn, 0), ca = __shfl_sync(full, cs, 0);
      const double sb = __shfl_sync(full, sn, 1), cb = __shfl_sync(full, cs, 1);
      const double sg = __shfl_sync(full, sn, 2), cg = __shfl_sync(full, cs, 2);
      T[0] = (float)(cg * cb);
      T[1] = (float)(-sg * ca + cg * sb * sa);
      T[2] = (float)(sg * sa + cg * sb * ca);
      T[4] = (float)(sg * cb);
      T[5] = (float)(cg * ca + sg * sb * sa);
      T[6] = (float)(-cg * sa + sg * sb * ca);
      T[8] = (float)(-sb);
      T[9] = (float)(cb * sa);
      T[10] = (float)(cb * ca);
      T[3] = (float)x[3];
      T[7] = (float)x[4];
      T[11] = (float)x[5];
    }
  } else {
    double mu_s[3], mu_d[3], S[9], R[9];
    for (int k = 0; k < 3; ++k) {
      mu_s[k] = v[k] / cnt;
      mu_d[k] = v[3 + k] / cnt;
    }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) S[i * 3 + j] = v[6 + i * 3 + j] / cnt - mu_d[i] * mu_s[j];
    kabsch_rotation_dev(S, R);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T[i * 4 + j] = (float)R[i * 3 + j];
      T[i * 4 + 3] =
          (float)(mu_d[i] - (R[i * 3 + 0] * mu_s[0] + R[i * 3 + 1] * mu_s[1] + R[i * 3 + 2] * mu_s[2]));
    }
  }
  // final_transformation_ = transformation_ * final_transformation_ (Matrix4f, float32):
  // lane L < 16 computes entry (L / 4, L % 4)
  {
    const int i = (lane >> 2) & 3, j = lane & 3;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {  // row i of T without dynamic register indexing
      t0 = i == r ? T[r * 4 + 0] : t0;
      t1 = i == r ? T[r * 4 + 1] : t1;
      t2 = i == r ? T[r * 4 + 2] : t2;
      t3 = i == r ? T[r * 4 + 3] : t3;
    }
    float s = __fmul_rn(t0, __shfl_sync(full, Fin, 0 * 4 + j));
    s = __fadd_rn(s, __fmul_rn(t1, __shfl_sync(full, Fin, 1 * 4 + j)));
    s = __fadd_rn(s, __fmul_rn(t2, __shfl_sync(full, Fin, 2 * 4 + j)));
    s = __fadd_rn(s, __fmul_rn(t3, __shfl_sync(full, Fin, 3 * 4 + j)));
    if (lane < 16) {
      float tl = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) tl = lane == k ? T[k] : tl;
      st->T[lane] = tl;
      st->Tfinal[lane] = s;
    }
  }
  // DefaultConvergenceCriteria::hasConverged
  int state = LC3D_STATE_NOT_CONVERGED;
  bool keep_mse = false;
  if (iter >= cfg.max_iterations) {
    state = LC3D_STATE_ITERATIONS;
  } else {
    float tr = __fsub_rn(__fadd_rn(__fadd_rn(T[0], T[5]), T[10]), 1.0f);
    double cos_angle = 0.5 * (double)tr;
    double tsq = (double)__fmul_rn(T[3], T[3]) + (double)__fmul_rn(T[7], T[7]) +
                 (double)__fmul_rn(T[11], T[11]);
    if (cos_angle >= cfg.rot_thr && tsq <= cfg.transl_thr) {
      state = LC3D_STATE_TRANSFORM;
    } else {
      if (fabs(mse - prev_mse) < cfg.abs_mse)
        state = LC3D_STATE_ABS_MSE;
      else if (fabs(mse - prev_mse) / prev_mse < cfg.rel_mse)
        state = LC3D_STATE_REL_MSE;
      else
        keep_mse = true;
    }
  }
  if (lane == 0) {
    st->last_corr = (long long)cnt;
    st->last_mse = mse;
    st->iter = iter;
    if (keep_mse) st->prev_mse = mse;
    st->state = state;
    if (state != LC3D_STATE_NOT_CONVERGED) {
      st->converged = 1;
      st->done = 1;
    }
  }
}

// Deterministic grid-wide reduce of per-block partials by the last block, NV values.
// partials layout: [v * nblk + b].  Warp w reduces values w, w+8, ...; lanes stride over
// blocks in a fixed order, then a fixed-shape shuffle tree.
template <int NV>
__device__ __forceinline__ void last_block_reduce(const double* __restrict__ partials, int nblk,
                                                  double* out_smem) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
  for (int v = w; v < NV; v += nwarps) {
    const double* row = partials + (size_t)v * nblk;
    double s = 0.0;
#pragma unroll 8
    for (int b = lane; b < nblk; b += 32) s += __ldcg(row + b);  // L2 loads, independent: pipelined
    s = warp_sum(s);
    if (lane == 0) out_smem[v] = s;
  }
}

// Butterfly reduce-scatter of 32 per-lane values: after 5 exchange rounds lane L holds the
// warp total of value L.  31 shuffles instead of 32 x 5, fixed summation tree (deterministic).
// The values are produced by `val(i)` (i compile-time after unrolling) inside the first
// round, so only 16 doubles are ever live.
template <typename F>
__device__ __forceinline__ double warp_reduce_scatter32(F val, int lane) {
  const unsigned full = 0xffffffffu;
  double v[16];
  {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const double lo = val(i), hi = val(i + 16);
      const double send = up ? lo : hi;
      const double keep = up ? hi : lo;
      v[i] = keep + __shfl_xor_sync(full, send, 16);
    }
  }
#pragma unroll
  for (int h = 8; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const double send = up ? v[i] : v[i + h];
      const double keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, h);
    }
  }
  return v[0];
}

// estimator value i of one correspondence (compile-time i): point-to-plane layout
// [0..20] J^T J upper triangle, [21..26] J^T r, [27] d2, [28] 1; point-to-point layout
// [0..2] s, [3..5] d, [6..14] d s^T, [15] d2, [16] 1.
__device__ __forceinline__ double p2plane_value(int i, const double* J, double r, double d2, double one) {
  if (i < 21) {
    // (row, col) of the i-th upper-triangle entry; folds to constants once i is unrolled
    const int a = i < 6 ? 0 : i < 11 ? 1 : i < 15 ? 2 : i < 18 ? 3 : i < 20 ? 4 : 5;
    const int base = a == 0 ? 0 : a == 1 ? 6 : a == 2 ? 11 : a == 3 ? 15 : a == 4 ? 18 : 20;
    return J[a] * J[a + (i - base)];
  }
  if (i < 27) return J[i - 21] * r;
  if (i == 27) return d2;
  if (i == 28) return one;
  return 0.0;
}
__device__ __forceinline__ double p2p_value(int i, const double* sv, const double* dv, double d2, double one) {
  if (i < 3) return sv[i];
  if (i < 6) return dv[i - 3];
  if (i < 15) return dv[(i - 6) / 3] * sv[(i - 6) % 3];
  if (i == 15) return d2;
  if (i == 16) return one;
  return 0.0;
}

// The estimator sums of one warp (fp64): lane L returns the warp total of value L.  has: this lane
// holds a correspondence (q = transformed source point, j = sorted-target position of its match,
// d2 = its squared distance).  Shared by the fused iteration kernel and icp_estimate_kernel, so
// both produce bit-identical rows.
template <int MODE>
__device__ __forceinline__ double icp_estimator_acc(const GridDev& g, bool has, const float4 q, int j, float d2f,
                                                    int lane) {
  if (!__any_sync(0xffffffffu, has)) return 0.0;
  float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has) d = __ldg(&g.pts[j]);
  const double d2v = has ? (double)d2f : 0.0, one = has ? 1.0 : 0.0;
  if (MODE == LC3D_ICP_POINT_TO_PLANE) {
    double J[6] = {0, 0, 0, 0, 0, 0}, r = 0;
    if (has) {
      const float4 nn = __ldg(&g.nrm[j]);
      if (finite3(nn.x, nn.y, nn.z)) {
        // float32 products widened to double, as TransformationEstimationPointToPlaneLLS
        J[0] = (double)(nn.z * q.y - nn.y * q.z);
        J[1] = (double)(nn.x * q.z - nn.z * q.x);
        J[2] = (double)(nn.y * q.x - nn.x * q.y);
        J[3] = nn.x;
        J[4] = nn.y;
        J[5] = nn.z;
        r = (double)(nn.x * d.x + nn.y * d.y + nn.z * d.z - nn.x * q.x - nn.y * q.y - nn.z * q.z);
      }
    }
    return warp_reduce_scatter32([&](int i) { return p2plane_value(i, J, r, d2v, one); }, lane);
  }
  const double sv[3] = {has ? (double)q.x : 0.0, has ? (double)q.y : 0.0, has ? (double)q.z : 0.0};
  const double dv[3] = {d.x, d.y, d.z};
  return warp_reduce_scatter32([&](int i) { return p2p_value(i, sv, dv, d2v, one); }, lane);
}

// One partial row per BLOCK (value-major) without stalling the warps on a block barrier: every warp
// parks its row in shared memory; warps 1.. then ARRIVE at a named barrier and retire, warp 0 WAITS
// on it (a waiting warp takes no issue slot, and the block's registers are held until its slowest
// warp is done anyway), adds the rows in warp order (fixed order: deterministic) and writes the
// block's row.  The solve kernel reads 4x fewer rows.  (An earlier version let the LAST warp to
// arrive do the sum, ordered by __threadfence_block + a shared-memory ticket; same cost, but the
// named barrier is an ordering the hardware and compute-sanitizer both understand.)
#ifndef LC3D_ROW_TICKET
#define LC3D_ROW_TICKET 0
#endif
template <int NV>
__device__ __forceinline__ void icp_block_row(double acc, double (*s_rows)[32], unsigned* s_arrived,
                                              double* __restrict__ partials) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  s_rows[w][lane] = acc;
  __threadfence_block();
#if LC3D_ROW_TICKET
  __syncwarp();
  unsigned ticket = 0;
  if (lane == 0) ticket = atomicAdd(s_arrived, 1u);
  ticket = __shfl_sync(0xffffffffu, ticket, 0);
  if (ticket != (unsigned)(kIcpThreads / 32 - 1)) return;
  __syncwarp();
  __threadfence_block();
#else
  (void)s_arrived;
  if (w != 0) {
    asm volatile("bar.arrive 1, %0;" ::"n"(kIcpThreads) : "memory");
    return;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kIcpThreads) : "memory");
#endif
  if (lane < NV) {
    double s = 0.0;
#pragma unroll
    for (int ww = 0; ww < kIcpThreads / 32; ++ww) s += ((volatile double*)s_rows[ww])[lane];
    partials[(size_t)lane * gridDim.x + blockIdx.x] = s;
  }
}

// One ICP iteration = this kernel + icp_solve_kernel.  X: working copy of the source (float4,
// Morton order, w = original index), transformed in place.  One thread per source point; the
// 32 estimator sums of a warp are reduced with one butterfly reduce-scatter (lane L ends up
// with value L) and written as one partial row per warp.
//
// Per-query memory across iterations (temporal coherence):
//   Mj[i]  : sorted-target position of the previous match (seed: its distance is an exact
//            upper bound on the nearest-neighbour distance, which sizes the ball walk);
//   Bnd[i] : < 0  -> no target point within gate_distance + (-Bnd): while the accumulated
//            motion stays below that slack the query provably has no correspondence and
//            the search is skipped;  +inf -> nothing known.
// Searches run against an extended gate (r + margin)^2 so that rejected queries learn a
// slack; a correspondence is emitted iff d2 <= gate exactly as PCL does.
#ifndef LC3D_ICP_MINBLOCKS
#define LC3D_ICP_MINBLOCKS 8
#endif
// RINGS: centre-out ball walk (search.cuh), chosen by the host for the iterations right after the
// large first pose updates.
// SEARCH_ONLY: correspondences only (Mj, Bnd, the dump) — the estimator sums are left to
// icp_estimate_kernel.  The host-buffer path runs iteration 0 this way: the search needs no target
// normals, so their upload (PCIe) hides behind the most expensive search of the alignment.
template <int MODE, bool STATS, bool RINGS = false, bool SEARCH_ONLY = false>
__global__ void __launch_bounds__(kIcpThreads, LC3D_ICP_MINBLOCKS)
    icp_iteration_kernel(IcpState* __restrict__ st, const __grid_constant__ IcpConfig cfg,
                         const __grid_constant__ GridDev g,
                         float4* __restrict__ X, float* __restrict__ Bnd, int* __restrict__ Mj, int n,
                         double* __restrict__ partials, int32_t* __restrict__ dump_idx,
                         float* __restrict__ dump_d2) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  __shared__ float sT[16];
  __shared__ int s_flags[2];
  __shared__ double s_rows[kIcpThreads / 32][32];
  __shared__ unsigned s_arrived;
  pdl_wait();  // the previous solve kernel's pose / done flag
  if (threadIdx.x == 0) {
    s_flags[0] = st->done;
    s_flags[1] = st->iter;
    s_arrived = 0u;
  }
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
  __syncthreads();
  if (s_flags[0]) return;
  const int iter = s_flags[1];
  const int lane = threadIdx.x & 31;
  // the statistics plumbing is compiled out of the production instantiation: the search
  // kernel is register-bound and every live counter costs occupancy
  SearchStats* stats = (STATS && cfg.stats) ? cfg.stats + iter : nullptr;
  if (stats && threadIdx.x == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    atomicMin(&stats->c[11], t0);
    if (g_block_log) g_block_log[((size_t)iter * gridDim.x + blockIdx.x) * 3] = t0;
  }
  double acc = 0.0;  // lane L: total of estimator value L
  {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < n;
    float4 q = active ? X[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float bnd = (active && iter > 0) ? Bnd[i] : INFINITY;
    const int seed_j = (active && iter > 0) ? Mj[i] : -1;
    active = active && finite3(q.x, q.y, q.z);
    float delta = 0.0f;
    if (active && iter > 0) {  // transformCloud with the previous iteration's T
      const float x = xform_row(sT, 0, q.x, q.y, q.z);
      const float y = xform_row(sT, 1, q.x, q.y, q.z);
      const float z = xform_row(sT, 2, q.x, q.y, q.z);
      const float mx = x - q.x, my = y - q.y, mz = z - q.z;
      delta = sqrtf(mx * mx + my * my + mz * mz) * 1.00001f;
      q.x = x;
      q.y = y;
      q.z = z;
      X[i] = q;
    }
    bool skip = false;
    if (active && bnd < 0.0f) {
      const float slack = -bnd - delta;
      if (slack > cfg.slack_floor) {
        skip = true;
        bnd = -slack;
      }
    }
#ifndef LC3D_NO_OCC_REJECT
    if (active && !skip && seed_j < 0 && g.occ) {
      // no previous match: the dilated occupancy of the index may prove that nothing lies
      // within the gate (non-overlap regions of the first iterations) without any search
      const QueryCell qc = query_cell(g, q.x, q.y, q.z);
      if (occ_proves_empty(g, qc.ix, qc.iy, qc.iz)) {
        const float slack = ((float)g.occ_r - 0.01f) * g.c - cfg.gate_dist;
        if (slack > cfg.slack_floor) {
          skip = true;
          bnd = -slack;
        }
      }
    }
#endif
    const Best b = nn_search_seeded<false, RINGS>(g, active && !skip, q.x, q.y, q.z, cfg.gate_ext, seed_j, stats);
    const bool found = active && !skip && b.j >= 0;
    const bool has = found && b.d2 <= cfg.gate;
    if (active && !skip) {
      if (found) {
        if (has) {
          bnd = INFINITY;
        } else {  // exact nearest neighbour lies beyond the gate: nothing within d
          const float slack = sqrtf(b.d2) * 0.99999f - cfg.gate_dist;
          bnd = slack > cfg.slack_floor ? -slack : INFINITY;
        }
      } else {
        bnd = cfg.margin_slack > cfg.slack_floor ? -cfg.margin_slack : INFINITY;
      }
    }
    if (i < n) {
      Bnd[i] = bnd;
      Mj[i] = skip ? seed_j : (found ? b.j : -1);
    }
    if (dump_idx && iter == cfg.dump_iteration && i < n) {
      const int oi = __float_as_int(q.w);
      dump_idx[oi] = has ? b.oi : -1;
      dump_d2[oi] = has ? b.d2 : INFINITY;
    }
    if (!SEARCH_ONLY) acc = icp_estimator_acc<MODE>(g, has, q, b.j, b.d2, lane);
  }
  if (!SEARCH_ONLY) icp_block_row<NV>(acc, s_rows, &s_arrived, partials);
  if (STATS && g_block_log && lane == 0) {  // block end = its last warp's end
    unsigned long long t1;
    unsigned smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* rec = g_block_log + ((size_t)iter * gridDim.x + blockIdx.x) * 3;
    atomicMax(&rec[1], t1);
    rec[2] = smid;
  }
}

// The estimator half of an iteration whose search ran SEARCH_ONLY (iteration 0: no transform has
// been applied yet, X is the source as uploaded).  Same thread <-> point mapping, same arithmetic
// (the match distance is recomputed with the expression the search used) and the same row layout as
// the fused kernel: the partial rows are bit-identical.
template <int MODE>
__global__ void __launch_bounds__(kIcpThreads)
    icp_estimate_kernel(const IcpState* __restrict__ st, const __grid_constant__ IcpConfig cfg,
                        const __grid_constant__ GridDev g, const float4* __restrict__ X,
                        const int* __restrict__ Mj, int n, double* __restrict__ partials) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  __shared__ double s_rows[kIcpThreads / 32][32];
  __shared__ unsigned s_arrived;
  pdl_wait();
  if (threadIdx.x == 0) s_arrived = 0u;
  __syncthreads();
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  const float4 q = active ? X[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  const int j = active ? Mj[i] : -1;
  bool has = false;
  float d2 = 0.0f;
  if (j >= 0) {
    const float4 d = __ldg(&g.pts[j]);
    d2 = dist2_exact(q.x, q.y, q.z, d.x, d.y, d.z);
    has = d2 <= cfg.gate;
  }
  const double acc = icp_estimator_acc<MODE>(g, has, q, j, d2, lane);
  icp_block_row<NV>(acc, s_rows, &s_arrived, partials);
}

// Second kernel of an iteration: block v reduces estimator value v over all warp rows in a
// fixed order (deterministic), the last block to finish (atomic ticket) solves, composes the
// pose and runs the convergence test.  grid = NV blocks.
#ifndef LC3D_SOLVE_THREADS
#define LC3D_SOLVE_THREADS 512
#endif
constexpr int kSolveThreads = LC3D_SOLVE_THREADS;
template <int MODE>
__global__ void __launch_bounds__(kSolveThreads)
    icp_solve_kernel(IcpState* __restrict__ st, const IcpConfig cfg, const double* __restrict__ partials,
                     int nwarps, double* __restrict__ reduced) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  __shared__ double sm[kSolveThreads / 32];
  __shared__ unsigned s_ticket;
  pdl_wait();     // the search kernel's partial rows
  pdl_trigger();  // the next search kernel may stage its blocks behind this one
  if (st->done) return;
  SearchStats* stats = cfg.stats ? cfg.stats + st->iter : nullptr;  // LC3D_STATS=1 timeline
  unsigned long long tm0 = 0, tm1 = 0, tm2 = 0;
  if (stats) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm0));
  const int v = blockIdx.x;
  const double* row = partials + (size_t)v * nwarps;
  // fixed assignment and order (deterministic); 8 independent loads in flight per thread
  double s = 0.0;
  for (int base = threadIdx.x; base < nwarps; base += kSolveThreads * 8) {
    double t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + k * kSolveThreads;
      t[k] = i < nwarps ? __ldcg(row + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s += t[k];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < kSolveThreads / 32 ? sm[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) sm[0] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    reduced[v] = sm[0];
    __threadfence();
    s_ticket = atomicAdd(&st->ticket, 1u);
  }
  __syncthreads();
  if (s_ticket != (unsigned)(NV - 1)) return;
  if (threadIdx.x < 32) {  // one warp solves (icp_solve_and_test is warp-collective)
    __threadfence();
    const double mine = threadIdx.x < NV ? __ldcg(reduced + threadIdx.x) : 0.0;
    double red[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k] = __shfl_sync(0xffffffffu, mine, k);
    if (threadIdx.x == 0) st->ticket = 0;
    if (stats) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm1));
    icp_solve_and_test<MODE>(st, cfg, red);
    if (stats) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm2));
      if (threadIdx.x == 0) {
        stats->c[8] = tm0 - stats->c[11];  // search kernel start -> start of the last solve block
        stats->c[9] = tm1 - tm0;           // row reduce + ticket
        stats->c[10] = tm2 - tm1;          // solve + pose composition + criteria
      }
    }
  }
}

// ---- getFitnessScore (fine_registration.cpp:126) --------------------------
// transformPointCloud(*input_, tmp, final_transformation_) then the unbounded 1-NN of every
// point; fitness = sum d2 / count.  src0: ORIGINAL source (cell-sorted order).  Three kernels:
//   icp_fitness_kernel        one thread per point, seeded by the last iteration's matches;
//                             resolves every point whose search ball is a few cells wide and
//                             queues the rest (points far off the target: ~4 % at the bench);
//   icp_fitness_hard_kernel   one WARP per queued point (coarse-cell ring search), grid-stride
//                             over the queue, so the expensive queries are spread over the
//                             whole GPU instead of serialising the few warps that own them
//                             (measured before the split: SMs active 29 % of the kernel);
//   icp_fitness_reduce_kernel fixed-order sum of the per-point squared distances.
// d2_all[i]: squared distance, -1 = no neighbour / not a finite point.
struct FitQueue {
  int* idx;         // queued point (position in src0)
  int* seed;        // best candidate so far (sorted-target position) or -1
  unsigned* count;  // queue length (reset by the reduce kernel)
};

__global__ void __launch_bounds__(kIcpThreads, LC3D_ICP_MINBLOCKS)
    icp_fitness_kernel(const IcpState* __restrict__ st, const __grid_constant__ GridDev g,
                       const float4* __restrict__ src0, const int* __restrict__ Mj, int mj_stride, int n,
                       float* __restrict__ d2_all, const FitQueue fq, SearchStats* stats) {
  __shared__ float sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = st->Tfinal[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = i < n;
  float4 q = active ? src0[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  active = active && finite3(q.x, q.y, q.z);
  const float x = xform_row(sT, 0, q.x, q.y, q.z);
  const float y = xform_row(sT, 1, q.x, q.y, q.z);
  const float z = xform_row(sT, 2, q.x, q.y, q.z);
  // seeded by the last iteration's matches (the final pose differs from the incremental one
  // only by float rounding), unbounded: every source point counts
  const int seed_j = (active && Mj) ? Mj[(size_t)i * mj_stride] : -1;
  bool deferred = false;
  const Best b = nn_search_seeded<true>(g, active, x, y, z, INFINITY, seed_j, stats, &deferred);
  const unsigned dm = __ballot_sync(0xffffffffu, deferred);
  if (dm) {  // warp-aggregated append
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(fq.count, (unsigned)__popc(dm));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (deferred) {
      const unsigned pos = base + __popc(dm & ((1u << lane) - 1u));
      fq.idx[pos] = i;
      fq.seed[pos] = b.j;
    }
  }
  if (i < n) d2_all[i] = (active && !deferred && b.j >= 0) ? b.d2 : -1.0f;
}

#ifndef LC3D_HARD_THREADS
#define LC3D_HARD_THREADS 128
#endif
constexpr int kHardThreads = LC3D_HARD_THREADS;
__global__ void __launch_bounds__(kHardThreads)
    icp_fitness_hard_kernel(const IcpState* __restrict__ st, const __grid_constant__ GridDev g,
                            const float4* __restrict__ src0, float* __restrict__ d2_all,
                            const FitQueue fq) {
  __shared__ float sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = st->Tfinal[threadIdx.x];
  __syncthreads();
  const unsigned count = *fq.count;
  const int lane = threadIdx.x & 31;
  // contiguous queue chunks per warp: consecutive entries were appended together by one warp
  // of the search kernel (Morton neighbours), so each result seeds the next query
  const unsigned nw = gridDim.x * (blockDim.x >> 5);
  const unsigned chunk = (count + nw - 1) / nw;
  const unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const unsigned p0 = w * chunk, p1 = min(count, p0 + chunk);
  int last_j = -1;
  for (unsigned pos = p0; pos < p1; ++pos) {
    const int i = fq.idx[pos], sj = fq.seed[pos];
    const float4 q = src0[i];
    const float x = xform_row(sT, 0, q.x, q.y, q.z);
    const float y = xform_row(sT, 1, q.x, q.y, q.z);
    const float z = xform_row(sT, 2, q.x, q.y, q.z);
    Best b;
    b.d2 = INFINITY;
    b.j = -1;
    b.oi = 0x7fffffff;
    if (sj >= 0) consider(__ldg(&g.pts[sj]), sj, x, y, z, b);
    if (last_j >= 0) consider(__ldg(&g.pts[last_j]), last_j, x, y, z, b);
    if (!nn_ball_warp(g, x, y, z, b)) nn_phase2_warp(g, x, y, z, b);
    last_j = b.j;
    if (lane == 0) d2_all[i] = b.j >= 0 ? b.d2 : -1.0f;
  }
}

constexpr int kFitReduceBlocks = 148;
__global__ void __launch_bounds__(256)
    icp_fitness_reduce_kernel(IcpState* __restrict__ st, const float* __restrict__ d2_all, int n,
                              double* __restrict__ partials, unsigned* __restrict__ queue_count) {
  __shared__ double ssum[256], scnt[256];
  __shared__ double red[2];
  __shared__ unsigned s_ticket;
  const int nblk = gridDim.x;
  const int chunk = (n + nblk - 1) / nblk;
  const int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  double s = 0.0, c = 0.0;
  for (int i = lo + threadIdx.x; i < hi; i += 256) {
    const float d2 = __ldcg(d2_all + i);
    if (d2 >= 0.0f) {
      s += (double)d2;
      c += 1.0;
    }
  }
  ssum[threadIdx.x] = s;
  scnt[threadIdx.x] = c;
  __syncthreads();
#pragma unroll
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
      scnt[threadIdx.x] += scnt[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = ssum[0];
    partials[nblk + blockIdx.x] = scnt[0];
    __threadfence();
    s_ticket = atomicAdd(&st->ticket2, 1u);
  }
  __syncthreads();
  if (s_ticket != (unsigned)(nblk - 1)) return;
  __threadfence();
  last_block_reduce<2>(partials, nblk, red);
  __syncthreads();
  if (threadIdx.x == 0) {
    st->ticket2 = 0;
    *queue_count = 0;
    st->fitness_sum = red[0];
    st->fitness_cnt = (long long)red[1];
  }
}

// ---- one pair sharded by SOURCE points over the GPUs of a box (second mode) --------
// Every rank holds the whole target + index and a slice of the source.  Per iteration each
// rank's search kernel writes its partial estimator rows into an exchange buffer that the other
// ranks have mapped (CUDA IPC, peer access over NVLink / NVSwitch) and publishes them with an
// iteration stamp; the solve kernel of EVERY rank then reads all ranks' rows straight from
// peer memory, adds them in rank order and solves — bit-identical pose and convergence decision on
// all ranks, no separate all-reduce, no broadcast.  Rows are double-buffered by iteration parity:
// a rank can only overwrite buffer p two iterations later, after every peer has published the
// iteration in between, i.e. finished reading.
constexpr int kShardMaxWorld = 8;
struct ShardHeader {
  unsigned stamp;   // iterations published so far (k + 1 after the rows of iteration k are complete)
  unsigned nblk;    // partial rows per value of this rank
  unsigned pad[2];
};
struct ShardView {
  int world, rank;
  unsigned base;                           // epoch stamp offset of this alignment (same on all ranks)
  const ShardHeader* hdr[kShardMaxWorld];  // peers' headers (own entry included)
  const double* rows[kShardMaxWorld];      // peers' row buffers: 2 parities x 32 values x row_stride doubles
  long long row_stride;                    // row capacity per value (same on all ranks); rows are packed with pitch nblk
};

// stream-ordered after the search kernel: its rows are complete and visible
__global__ void shard_publish_kernel(const IcpState* __restrict__ st, ShardHeader* hdr, unsigned nblk, unsigned base) {
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x != 0 || st->done) return;
  hdr->nblk = nblk;
  __threadfence_system();
  *(volatile unsigned*)&hdr->stamp = base + (unsigned)st->iter + 1u;
  __threadfence_system();
}

template <int MODE>
__global__ void __launch_bounds__(kSolveThreads)
    icp_solve_sharded_kernel(IcpState* __restrict__ st, const IcpConfig cfg, const ShardView sv,
                             double* __restrict__ reduced) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  __shared__ double sm[kSolveThreads / 32];
  __shared__ unsigned s_ticket;
  __shared__ int s_fail;
  pdl_wait();
  pdl_trigger();
  if (st->done) return;
  const int iter = st->iter;
  const int v = blockIdx.x;
  // wait until every rank has published this iteration's rows (system-scope polling over peer memory)
  if (threadIdx.x == 0) {
    s_fail = 0;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (int r = 0; r < sv.world; ++r) {
      while (*(volatile const unsigned*)&sv.hdr[r]->stamp < sv.base + (unsigned)iter + 1u) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) {  // 20 s: a peer died
          s_fail = 1;
          break;
        }
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  const bool fail = s_fail != 0;
  double total = 0.0;
  if (!fail) {
    for (int r = 0; r < sv.world; ++r) {  // rank order: the same arithmetic on every rank
      const int nrows = (int)*(volatile const unsigned*)&sv.hdr[r]->nblk;
      const double* row = sv.rows[r] + (size_t)(iter & 1) * 32 * sv.row_stride + (size_t)v * nrows;
      double s = 0.0;
      for (int i = threadIdx.x; i < nrows; i += kSolveThreads) s += *(volatile const double*)(row + i);
      s = warp_sum(s);
      if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
      __syncthreads();
      if (threadIdx.x < 32) {
        double t = threadIdx.x < kSolveThreads / 32 ? sm[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) total += t;
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    reduced[v] = total;
    __threadfence();
    s_ticket = atomicAdd(&st->ticket, 1u);
  }
  __syncthreads();
  if (s_ticket != (unsigned)(NV - 1)) return;
  if (threadIdx.x < 32) {
    __threadfence();
    const double mine = threadIdx.x < NV ? __ldcg(reduced + threadIdx.x) : 0.0;
    double red[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) red[k] = __shfl_sync(0xffffffffu, mine, k);
    if (threadIdx.x == 0) st->ticket = 0;
    if (fail) {
      if (threadIdx.x == 0) {
        st->pad[0] = 1;  // exchange timed out
        st->converged = 0;
        st->done = 1;
      }
    } else {
      icp_solve_and_test<MODE>(st, cfg, red);
    }
  }
}

// registered = transformCloud(*input_, final_transformation_) (fine_registration.cpp:121),
// also serves lc3d_transform (pcl_tools/transform.cpp:84-90).  Input order, packed xyz out.
__global__ void __launch_bounds__(256)
    transform_kernel(const float* __restrict__ T16, const float4* __restrict__ xyz,
                     const float4* __restrict__ nrm, int n, float* __restrict__ out_xyz,
                     float* __restrict__ out_nrm) {
  __shared__ float sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = T16[threadIdx.x];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = xyz[i];
  float ox = p.x, oy = p.y, oz = p.z;
  if (finite3(p.x, p.y, p.z)) {
    ox = xform_row(sT, 0, p.x, p.y, p.z);
    oy = xform_row(sT, 1, p.x, p.y, p.z);
    oz = xform_row(sT, 2, p.x, p.y, p.z);
  }
  out_xyz[3 * (size_t)i + 0] = ox;
  out_xyz[3 * (size_t)i + 1] = oy;
  out_xyz[3 * (size_t)i + 2] = oz;
  if (nrm && out_nrm) {
    float4 m = nrm[i];
    float nx = m.x, ny = m.y, nz = m.z;
    if (finite3(m.x, m.y, m.z)) {
      nx = rot_row(sT, 0, m.x, m.y, m.z);
      ny = rot_row(sT, 1, m.x, m.y, m.z);
      nz = rot_row(sT, 2, m.x, m.y, m.z);
    }
    out_nrm[3 * (size_t)i + 0] = nx;
    out_nrm[3 * (size_t)i + 1] = ny;
    out_nrm[3 * (size_t)i + 2] = nz;
  }
}

// Plain batched 1-NN (lc3d_nn): queries in cell-sorted order, results scattered back to
// the original query order.
__global__ void __launch_bounds__(256)
    nn_kernel(const __grid_constant__ GridDev g, const float4* __restrict__ q_sorted, int n, float gate,
              int32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = i < n;
  float4 q = active ? q_sorted[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  active = active && finite3(q.x, q.y, q.z);
  Best b = nn_search(g, active, q.x, q.y, q.z, gate);
  if (i < n) {
    const int oi = __float_as_int(q.w);
    const bool has = active && b.j >= 0;
    out_idx[oi] = has ? b.oi : -1;
    out_d2[oi] = has ? b.d2 : INFINITY;
  }
}

}  // namespace lc3d
