// knn.cuh — exact k-nearest-neighbour search on the uniform grid, one warp per query, and
// the two consumers the reference feeds with it:
//   * pcl::StatisticalOutlierRemoval (outlier_removal.cpp:80-84): mean distance
//     to the k nearest neighbours,
//   * pcl::NormalEstimation (normal_estimation.cpp:84-108): float32 single-pass
//     covariance -> closed-form smallest eigenpair (pcl::eigen33) -> viewpoint flip.
//
// Search: the warp takes the (2K+1)^3 block of cells around the query and its GUARANTEED ball
// (radius = distance from the query to the nearest block face that still has grid behind it:
// every indexed point inside that ball lies inside the block).  The (2K+1)^2 cell rows are
// located by the lanes in parallel (one row per lane: slab test against the ball, x-run clipped
// to the ball, the two cell_start loads), a warp scan turns the run lengths into offsets, and
// the candidates are then evaluated 32 at a time whatever row they come from (each lane finds
// its run with a 5-step shuffle search).  Only points INSIDE the guaranteed ball are kept
// (ballot/popc append to a per-warp shared-memory buffer): if there are at least k of them, the
// k smallest keys (d2 bits << 32 | original index) are the exact k nearest neighbours and ONE
// bitonic sort of the buffer yields them in ascending (d2, index) order — the order PCL's
// consumers sum in.  Otherwise the block grows (by the density it has just seen) and the search
// starts over.  The index cell edge is chosen from k (knn_cell_factor) so that K = 2 holds
// ~1.4 k points in the ball: 25 rows = one lane-parallel step, one sort, no second ring.
#pragma once
#include "search.cuh"

namespace lc3d {

constexpr int kKnnWarps = 4;  // warps (= queries in flight) per block
constexpr unsigned long long kMaxKey = 0xffffffffffffffffull;

template <int N>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long* buf, int lane) {
  for (int kk = 2; kk <= N; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int t = lane; t < N / 2; t += 32) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        const bool asc = (i & kk) == 0;
        const unsigned long long a = buf[i], b = buf[l];
        if ((a > b) == asc) {
          buf[i] = b;
          buf[l] = a;
        }
      }
      __syncwarp();
    }
  }
}

// The same network with the keys in registers: element i = r * 32 + lane lives in register r of its
// lane, so compare-exchange distances below 32 are one __shfl_xor per key and distances of 32 and
// more pair two registers of the same lane.  Same instruction count as the shared-memory version
// but half its dependent latency per step (no load -> compare -> store -> __syncwarp round trip), and
// the R keys of a lane are independent work — the k-NN kernels wait on this sort.
// buf[0..cnt) in, buf[0..32R) sorted ascending out (padded with kMaxKey).
template <int R>
__device__ __forceinline__ void warp_bitonic_sort_regs(unsigned long long* buf, int cnt, int lane) {
  const unsigned full = 0xffffffffu;
  unsigned long long v[R];
#pragma unroll
  for (int r = 0; r < R; ++r) v[r] = (r * 32 + lane < cnt) ? buf[r * 32 + lane] : 0xffffffffffffffffull;
#pragma unroll
  for (int kk = 2; kk <= 32 * R; kk <<= 1) {
#pragma unroll
    for (int j = kk >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int jr = j >> 5;
          if ((r & jr) == 0) {
            const bool asc = ((r * 32) & kk) == 0;  // kk >= 64 here: the lane bits do not matter
            const unsigned long long a = v[r], b = v[r ^ jr];
            const bool sw = (a > b) == asc;
            v[r] = sw ? b : a;
            v[r ^ jr] = sw ? a : b;
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const bool asc = ((r * 32 + lane) & kk) == 0;
          const unsigned long long a = v[r];
          const unsigned long long b = __shfl_xor_sync(full, a, j);
          const bool take_min = lower == asc;
          v[r] = ((a < b) == take_min) ? a : b;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) buf[r * 32 + lane] = v[r];
  __syncwarp();
}

template <int CAP>
struct KnnState {
  unsigned long long* buf;  // CAP keys, per warp, shared memory
  int cnt;                  // valid keys in buf (warp-uniform)
  unsigned long long thresh;  // keys >= thresh cannot enter the top-k (warp-uniform)
  int k;
};

// Sorts the buffer ascending (padded to the next power of two >= cnt, at least 32) and keeps
// the k smallest keys.
template <int CAP>
__device__ __forceinline__ void knn_sort_truncate(KnnState<CAP>& s, int lane) {
  const int n = s.cnt <= 32 ? 32 : s.cnt <= 64 ? 64 : s.cnt <= 128 ? 128 : s.cnt <= 256 ? 256 : 512;
#ifndef LC3D_KNN_SMEM_SORT
  __syncwarp();  // the appended keys are visible to every lane
  if (n == 32) {
    warp_bitonic_sort_regs<1>(s.buf, s.cnt, lane);
  } else if (n == 64) {
    warp_bitonic_sort_regs<(CAP >= 64 ? 2 : 1)>(s.buf, s.cnt, lane);
  } else if (n == 128) {
    warp_bitonic_sort_regs<(CAP >= 128 ? 4 : 1)>(s.buf, s.cnt, lane);
  } else
#endif
  {
    for (int t = s.cnt + lane; t < n; t += 32) s.buf[t] = kMaxKey;
    __syncwarp();
    if (n == 32) warp_bitonic_sort<32>(s.buf, lane);
    else if (n == 64) warp_bitonic_sort<(CAP >= 64 ? 64 : CAP)>(s.buf, lane);
    else if (n == 128) warp_bitonic_sort<(CAP >= 128 ? 128 : CAP)>(s.buf, lane);
    else if (n == 256) warp_bitonic_sort<(CAP >= 256 ? 256 : CAP)>(s.buf, lane);
    else warp_bitonic_sort<CAP>(s.buf, lane);
  }
  if (s.cnt > s.k) s.cnt = s.k;
  s.thresh = s.cnt == s.k ? s.buf[s.k - 1] : kMaxKey;
}

// Exact kNN of (qx,qy,qz); on return s.buf[0..s.cnt) holds the neighbours ascending.
// Warp-uniform.  k <= CAP - 32.  kfirst: first block half-width (cells), from the index density.
// mask (nullable): only points whose ORIGINAL index has mask[index] != 0 exist for this search.
template <int CAP>
__device__ void knn_search_warp(const GridDev& g, float qx, float qy, float qz, KnnState<CAP>& s,
                                int lane, int kfirst, const uint32_t* __restrict__ mask = nullptr) {
  const unsigned full = 0xffffffffu;
  const unsigned lt = (1u << lane) - 1u;
  s.cnt = 0;
  s.thresh = kMaxKey;
  if (g.n == 0) return;
  // the block logic below assumes cubic cells: k-NN indices are built with xsub = 1
  const QueryCell qc = query_cell(g, qx, qy, qz);
  const float c2 = g.c * g.c * 0.9999f;
  // first block that touches the grid (queries may lie outside it)
  int K = kfirst;
  K = max(K, max(-qc.ix, qc.ix - (g.dx - 1)));
  K = max(K, max(-qc.iy, qc.iy - (g.dy - 1)));
  K = max(K, max(-qc.iz, qc.iz - (g.dz - 1)));
  for (;;) {
    // guaranteed radius of block K (faces with no grid behind them bound nothing)
    float gmin = 1.0e30f;
    bool open = false;
    if (qc.ix - K > 0) { gmin = fminf(gmin, qc.fx - (float)(qc.ix - K)); open = true; }
    if (qc.ix + K + 1 < g.dx) { gmin = fminf(gmin, (float)(qc.ix + K + 1) - qc.fx); open = true; }
    if (qc.iy - K > 0) { gmin = fminf(gmin, qc.fy - (float)(qc.iy - K)); open = true; }
    if (qc.iy + K + 1 < g.dy) { gmin = fminf(gmin, (float)(qc.iy + K + 1) - qc.fy); open = true; }
    if (qc.iz - K > 0) { gmin = fminf(gmin, qc.fz - (float)(qc.iz - K)); open = true; }
    if (qc.iz + K + 1 < g.dz) { gmin = fminf(gmin, (float)(qc.iz + K + 1) - qc.fz); open = true; }
    gmin = fmaxf(gmin - kCellSlack, 0.0f);
    const float rg2 = open ? gmin * gmin * c2 : INFINITY;  // squared guaranteed radius
    const float rgc2 = open ? gmin * gmin : INFINITY;      // the same in cells^2
    const int side = 2 * K + 1, total = side * side;
    const float inv_side = 1.0f / (float)side;
    for (int t0 = 0; t0 < total; t0 += 32) {
      // ---- one (y,z) row per lane: its point run inside the guaranteed ball
      const int t = t0 + lane;
      uint32_t rs = 0, rl = 0;
      if (t < total) {
        int rz = (int)(((float)t + 0.5f) * inv_side);
        rz -= (rz * side > t);  // exact for any side the int range allows
        rz += ((rz + 1) * side <= t);
        const int yy = qc.iy + (t - rz * side) - K, zz = qc.iz + rz - K;
        if ((unsigned)yy < (unsigned)g.dy && (unsigned)zz < (unsigned)g.dz) {
          const float gy = slab_gap(qc.fy, yy, yy), gz = slab_gap(qc.fz, zz, zz);
          const float rem = rgc2 - (gy * gy + gz * gz);
          if (rem >= 0.0f) {
            int xa = max(qc.ix - K, 0), xb = min(qc.ix + K, g.dx - 1);
            if (open) {  // clip the run to the x-extent of the ball
              const float wx = sqrtf(rem) + 2.0f * kCellSlack;
              xa = max(xa, (int)floorf(qc.fx - wx));
              xb = min(xb, (int)floorf(qc.fx + wx));
            }
            if (xa <= xb) {
              const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
              rs = __ldg(row + xa);
              rl = __ldg(row + xb + 1) - rs;
            }
          }
        }
      }
      // ---- run lengths -> offsets (inclusive warp scan)
      uint32_t inc = rl;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(full, inc, o);
        if (lane >= o) inc += v;
      }
      const uint32_t off = inc - rl;
      const uint32_t T = __shfl_sync(full, inc, 31);
      // ---- candidates, 32 at a time whatever their row
      for (uint32_t base = 0; base < T; base += 32) {
        const uint32_t c = base + lane;
        int r = 0;  // the last run whose offset is <= c: the non-empty run that holds candidate c
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int pr = r + step;
          const uint32_t v = __shfl_sync(full, off, pr & 31);
          if (pr < 32 && v <= c) r = pr;
        }
        const uint32_t rs_r = __shfl_sync(full, rs, r), off_r = __shfl_sync(full, off, r);
        bool pass = false;
        unsigned long long key = 0;
        if (c < T) {
          const float4 p = __ldg(&g.pts[rs_r + (c - off_r)]);
          const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
          key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)__float_as_int(p.w);
          pass = d2 <= rg2 && key < s.thresh;
          if (mask && pass) pass = __ldg(&mask[__float_as_int(p.w)]) != 0u;
        }
        const unsigned m = __ballot_sync(full, pass);
        if (pass) s.buf[s.cnt + __popc(m & lt)] = key;
        s.cnt += __popc(m);
        __syncwarp();
        if (s.cnt > CAP - 32) knn_sort_truncate<CAP>(s, lane);  // crowded cells: keep the best k so far
      }
    }
    const int found = s.cnt;  // points inside the guaranteed ball (>= k if the buffer was truncated)
    knn_sort_truncate<CAP>(s, lane);
    if (!open) break;              // the block covers the whole grid
    if (s.cnt == s.k) break;       // k points inside the guaranteed ball: exact
    // too few: grow the block by the density just seen (at least one cell) and start over
    const float grow = sqrtf(1.3f * (float)s.k / (float)max(found, 1));
    K = max(K + 1, min(2 * K + 1, (int)ceilf((float)K * grow)));
    s.cnt = 0;
    s.thresh = kMaxKey;
  }
}

// ---- pcl::eigen33 smallest eigenpair, float32, closed form.  This TU is
// compiled with -fmad=false so the float expressions below evaluate as written.
__device__ __forceinline__ void roots2_dev(float b, float c, float* r) {
  r[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}
__device__ __forceinline__ void swapf(float& a, float& b) {
  float t = a;
  a = b;
  b = t;
}
__device__ void roots3_dev(const float* m, float* r) {
  float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] - m[4] * m[2] * m[2] -
             m[8] * m[1] * m[1];
  float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  float c2 = m[0] + m[4] + m[8];
  if (fabsf(c0) < 1.1920929e-07f) {  // FLT_EPSILON
    roots2_dev(c2, c1, r);
    return;
  }
  const float inv3 = (float)(1.0 / 3.0);
  const float sqrt3 = sqrtf(3.0f);
  float c2_3 = c2 * inv3;
  float a_3 = (c1 - c2 * c2_3) * inv3;
  if (a_3 > 0.0f) a_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
  float q = half_b * half_b + a_3 * a_3 * a_3;
  if (q > 0.0f) q = 0.0f;
  float rho = sqrtf(-a_3);
  float theta = atan2f(sqrtf(-q), half_b) * inv3;
  float ct = cosf(theta), st = sinf(theta);
  r[0] = c2_3 + 2.0f * rho * ct;
  r[1] = c2_3 - rho * (ct + sqrt3 * st);
  r[2] = c2_3 - rho * (ct - sqrt3 * st);
  if (r[0] >= r[1]) swapf(r[0], r[1]);
  if (r[1] >= r[2]) {
    swapf(r[1], r[2]);
    if (r[0] >= r[1]) swapf(r[0], r[1]);
  }
  if (r[0] <= 0.0f) roots2_dev(c2, c1, r);
}
__device__ __forceinline__ void cross3_dev(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ void eigen33_smallest_dev(const float* C, float* eigenvalue, float* v) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(C[i]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;  // FLT_MIN
  float m[9];
  for (int i = 0; i < 9; ++i) m[i] = C[i] / scale;
  float r[3];
  roots3_dev(m, r);
  *eigenvalue = r[0] * scale;
  m[0] -= r[0];
  m[4] -= r[0];
  m[8] -= r[0];
  float v1[3], v2[3], v3[3];
  cross3_dev(&m[0], &m[3], v1);
  cross3_dev(&m[0], &m[6], v2);
  cross3_dev(&m[3], &m[6], v3);
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
  float s = sqrtf(len);
  for (int k = 0; k < 3; ++k) v[k] = best[k] / s;
}

enum KnnConsumer { kConsumeIndices = 0, kConsumeSor = 1, kConsumeNormals = 2 };

struct KnnArgs {
  const float4* queries;  // cell-sorted order, w = original query index
  int nq;
  int k;
  const float4* xyz_in;  // cloud in input order (normals consumer gathers neighbours here)
  // outputs (indexed by original query index)
  int32_t* out_idx;   // nq x k   (kConsumeIndices)
  float* out_d2;      // nq x k
  float* out_mean;    // nq       (kConsumeSor): mean distance, 0 for invalid
  uint8_t* out_valid; // nq       (kConsumeSor)
  float* out_normal;  // nq x 3   (kConsumeNormals)
  float* out_curv;    // nq
  float vpx, vpy, vpz;
  int kfirst;  // first block half-width in cells (knn_kfirst)
  // kConsumeSor, optional: the neighbour lists themselves (nq x k original indices, ascending
  // (d2, index), -1 = none), for a later pass to reuse
  int32_t* out_lists;
  // kConsumeNormals on a SUBSET of the indexed cloud (lc3d_prepare_view: normals of the points SOR
  // kept, on SOR's own index and neighbour lists): keep[i] != 0 marks the points that exist,
  // remap[i] = output slot of query i, lists / list_k = neighbour lists of a k-NN pass over the whole
  // cloud with list_k >= k.  The k nearest KEPT neighbours of a kept query are the first k kept
  // entries of its list whenever the list holds that many (every kept point inside the list's
  // radius is on the list, in the same (d2, index) order); otherwise the query is searched on the
  // index with the mask.  All null / 0 for the plain pass.
  const uint32_t* keep;
  const uint32_t* remap;
  const int32_t* lists;
  int list_k;
};

template <int CAP, int CONSUMER>
// (56 registers, 9 blocks per SM; forcing 12 or 16 blocks or allowing 80 registers changes the SOR /
// normals stages by less than the run-to-run spread: the warps wait on the shared-memory sort)
__global__ void __launch_bounds__(kKnnWarps * 32) knn_kernel(const __grid_constant__ GridDev g, const __grid_constant__ KnnArgs a) {
  __shared__ unsigned long long sbuf[kKnnWarps][CAP];
  __shared__ float snb[CONSUMER == kConsumeNormals ? kKnnWarps : 1][CONSUMER == kConsumeNormals ? CAP : 1][3];
  __shared__ double sdist[CONSUMER == kConsumeSor ? kKnnWarps : 1][CONSUMER == kConsumeSor ? CAP : 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int qi = blockIdx.x * kKnnWarps + w;
  if (qi >= a.nq) return;
  const float4 q = a.queries[qi];
  const int oq = __float_as_int(q.w);
  KnnState<CAP> s;
  s.buf = sbuf[w];
  s.k = a.k;
  s.cnt = 0;
  s.thresh = kMaxKey;
  const bool ok = finite3(q.x, q.y, q.z);
  int oslot = oq;  // output slot
  if (CONSUMER == kConsumeNormals && a.keep) {
    if (__ldg(&a.keep[oq]) == 0u) return;  // the query itself was filtered out (warp-uniform)
    oslot = (int)__ldg(&a.remap[oq]);
    // the first k kept entries of the query's neighbour list
    const unsigned lt = (1u << lane) - 1u;
    int cnt = 0;
    for (int base = 0; base < a.list_k && cnt < a.k; base += 32) {
      const int t = base + lane;
      const int idx = t < a.list_k ? __ldg(&a.lists[(size_t)oq * a.list_k + t]) : -1;
      const bool kp = idx >= 0 && __ldg(&a.keep[idx]) != 0u;
      const unsigned m = __ballot_sync(0xffffffffu, kp);
      const int pos = cnt + __popc(m & lt);
      if (kp && pos < a.k) s.buf[pos] = (unsigned long long)(unsigned)idx;  // the consumer reads the index only
      cnt += __popc(m);
    }
    __syncwarp();
    if (cnt >= a.k) {
      s.cnt = a.k;
    } else if (ok) {  // too many of the listed neighbours were filtered out: exact masked search
      knn_search_warp<CAP>(g, q.x, q.y, q.z, s, lane, a.kfirst, a.keep);
    }
  } else if (ok) {
    knn_search_warp<CAP>(g, q.x, q.y, q.z, s, lane, a.kfirst);
  }
  __syncwarp();
  if (CONSUMER == kConsumeIndices) {
    for (int t = lane; t < a.k; t += 32) {
      const bool has = t < s.cnt;
      const unsigned long long key = has ? s.buf[t] : 0ull;
      a.out_idx[(size_t)oq * a.k + t] = has ? (int)(unsigned)(key & 0xffffffffu) : -1;
      a.out_d2[(size_t)oq * a.k + t] = has ? __uint_as_float((unsigned)(key >> 32)) : INFINITY;
    }
  } else if (CONSUMER == kConsumeSor) {
    // dist_sum = sum_{j=1..k-1} sqrt(d2_j) in double, ascending order (index 0 = the query
    // itself); distances[i] = (float)(dist_sum / mean_k), mean_k = k - 1.
    for (int t = lane; t < s.cnt; t += 32)
      sdist[w][t] = sqrt((double)__uint_as_float((unsigned)(s.buf[t] >> 32)));
    __syncwarp();
    if (lane == 0) {
      float md = 0.0f;
      uint8_t valid = 0;
      if (ok && s.cnt > 0) {
        double sum = 0.0;
        for (int t = 1; t < s.cnt; ++t) sum += sdist[w][t];
        md = (float)(sum / (double)(a.k - 1));
        valid = 1;
      }
      a.out_mean[oq] = md;
      a.out_valid[oq] = valid;
    }
    if (a.out_lists)
      for (int t = lane; t < a.k; t += 32)
        a.out_lists[(size_t)oq * a.k + t] = t < s.cnt ? (int)(unsigned)(s.buf[t] & 0xffffffffu) : -1;
  } else {
    // gather the neighbours' coordinates in parallel, then accumulate the 9 float32 sums of
    // computeMeanAndCovarianceMatrix sequentially in neighbour order (lane a owns sum a)
    for (int t = lane; t < s.cnt; t += 32) {
      const float4 p = __ldg(&a.xyz_in[(unsigned)(s.buf[t] & 0xffffffffu)]);
      snb[w][t][0] = p.x;
      snb[w][t][1] = p.y;
      snb[w][t][2] = p.z;
    }
    __syncwarp();
    float acc = 0.0f;
    if (lane < 9) {
      // accu layout: xx xy xz yy yz zz x y z
      const int ia = lane < 3 ? 0 : lane < 5 ? 1 : lane == 5 ? 2 : lane - 6;
      const int ib = lane < 3 ? lane : lane < 5 ? lane - 2 : lane == 5 ? 2 : -1;
      for (int t = 0; t < s.cnt; ++t) {
        const float va = snb[w][t][ia];
        const float term = ib >= 0 ? va * snb[w][t][ib] : va;
        acc += term;
      }
      acc /= (float)s.cnt;
    }
    float accu[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) accu[i] = __shfl_sync(0xffffffffu, acc, i);
    if (lane == 0) {
      float nx, ny, nz, curv;
      if (!ok || s.cnt < 3) {
        nx = ny = nz = curv = __int_as_float(0x7fc00000);
      } else {
        float C[9];
        C[0] = accu[0] - accu[6] * accu[6];
        C[1] = accu[1] - accu[6] * accu[7];
        C[2] = accu[2] - accu[6] * accu[8];
        C[4] = accu[3] - accu[7] * accu[7];
        C[5] = accu[4] - accu[7] * accu[8];
        C[8] = accu[5] - accu[8] * accu[8];
        C[3] = C[1];
        C[6] = C[2];
        C[7] = C[5];
        float ev, v[3];
        eigen33_smallest_dev(C, &ev, v);
        const float tr = C[0] + C[4] + C[8];
        curv = tr != 0.0f ? fabsf(ev / tr) : 0.0f;
        const float vx = a.vpx - q.x, vy = a.vpy - q.y, vz = a.vpz - q.z;
        const float cs = vx * v[0] + vy * v[1] + vz * v[2];
        if (cs < 0.0f) {
          v[0] *= -1.0f;
          v[1] *= -1.0f;
          v[2] *= -1.0f;
        }
        nx = v[0];
        ny = v[1];
        nz = v[2];
      }
      a.out_normal[3 * (size_t)oslot + 0] = nx;
      a.out_normal[3 * (size_t)oslot + 1] = ny;
      a.out_normal[3 * (size_t)oslot + 2] = nz;
      a.out_curv[oslot] = curv;
    }
  }
}

template <int CONSUMER>
inline void knn_launch(lc3d_ctx* ctx, const GridDev& g, const KnnArgs& a) {
  if (a.nq == 0) return;
  const int grid = div_up(a.nq, kKnnWarps);
  if (a.k <= 32)
    LC3D_LAUNCH(ctx, (knn_kernel<64, CONSUMER>), grid, kKnnWarps * 32, 0, g, a);
  else if (a.k <= 96)
    LC3D_LAUNCH(ctx, (knn_kernel<128, CONSUMER>), grid, kKnnWarps * 32, 0, g, a);
  else if (a.k <= 224)
    LC3D_LAUNCH(ctx, (knn_kernel<256, CONSUMER>), grid, kKnnWarps * 32, 0, g, a);
  else if (a.k <= 480)
    LC3D_LAUNCH(ctx, (knn_kernel<512, CONSUMER>), grid, kKnnWarps * 32, 0, g, a);
  else
    throw CudaError{"k too large (max 480)"};
}

// ---- SOR statistics + classification ------------------------------------
struct SorStats {
  double sum, sq_sum;
  long long valid;
  double mean, stddev, thr;
  unsigned ticket;
  unsigned pad;
};

__global__ void __launch_bounds__(256)
    sor_stats_kernel(const float* __restrict__ dist, const uint8_t* __restrict__ valid, int n,
                     double std_mul, double* __restrict__ partials, SorStats* __restrict__ st) {
  __shared__ double ws[8][3];
  __shared__ double red[3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double s = 0, sq = 0, c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (valid[i]) {
      const float d = dist[i];
      s += (double)d;
      sq += (double)(d * d);  // float32 product, as PCL
      c += 1.0;
    }
  }
  s = warp_sum(s);
  sq = warp_sum(sq);
  c = warp_sum(c);
  if (lane == 0) {
    ws[w][0] = s;
    ws[w][1] = sq;
    ws[w][2] = c;
  }
  __syncthreads();
  const int nblk = gridDim.x;
  if (threadIdx.x < 3) {
    double t = 0;
    for (int ww = 0; ww < 8; ++ww) t += ws[ww][threadIdx.x];
    partials[(size_t)threadIdx.x * nblk + blockIdx.x] = t;
  }
  __shared__ unsigned s_ticket;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_ticket = atomicAdd(&st->ticket, 1u);
  }
  __syncthreads();
  if (s_ticket != (unsigned)(nblk - 1)) return;
  __threadfence();
  {
    for (int v = w; v < 3; v += 8) {
      const double* row = partials + (size_t)v * nblk;
      double t = 0.0;
      for (int b = lane; b < nblk; b += 32) t += __ldcg(row + b);
      t = warp_sum(t);
      if (lane == 0) red[v] = t;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    st->ticket = 0;
    const double sum = red[0], sqs = red[1], nv = red[2];
    st->sum = sum;
    st->sq_sum = sqs;
    st->valid = (long long)nv;
    const double mean = sum / nv;
    const double var = (sqs - sum * sum / nv) / (nv - 1.0);
    const double sd = sqrt(var);
    st->mean = mean;
    st->stddev = sd;
    st->thr = mean + std_mul * sd;
  }
}

__global__ void __launch_bounds__(256)
    sor_flag_kernel(const float* __restrict__ dist, int n, const SorStats* __restrict__ st, int negative,
                    uint32_t* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool outlier = (double)dist[i] > st->thr;  // float vs double compare, as PCL
  flags[i] = (negative ? outlier : !outlier) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
    compact_indices_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, int n,
                           int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flags[i]) out[pos[i]] = i;
}

// pcl::compute3DCentroid: float32 sum in input order / count.  The sum is sequential by
// definition (bit-parity with the reference's float32 accumulation), but the loads are not: one
// warp fetches 32 points at a time (one coalesced 512-byte request) and every lane then adds them
// one by one in index order, so the result is the sequential sum while the memory latency is
// paid once per 32 points.  Launch with ONE warp.
__global__ void __launch_bounds__(32) centroid_kernel(const float4* __restrict__ xyz, int n, float* __restrict__ out) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int cnt = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
    bool ok = false;
    if (i < n) {
      p = xyz[i];
      ok = finite3(p.x, p.y, p.z);
    }
    unsigned m = __ballot_sync(full, ok);
    cnt += __popc(m);
    while (m) {  // ascending index order, non-finite points skipped as PCL does
      const int k = __ffs(m) - 1;
      m &= m - 1;
      sx += __shfl_sync(full, p.x, k);
      sy += __shfl_sync(full, p.y, k);
      sz += __shfl_sync(full, p.z, k);
    }
  }
  if (lane == 0) {
    const float fc = (float)cnt;
    out[0] = sx / fc;
    out[1] = sy / fc;
    out[2] = sz / fc;
    out[3] = 1.0f;
  }
}

}  // namespace lc3d
