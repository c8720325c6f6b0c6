// icp2.cuh — the ICP iteration kernel (second generation).
//
// One launch = transformCloud + determineCorrespondences + the estimator's sums of ONE
// iteration of pcl::IterativeClosestPoint::align (pcl_tools/fine_registration.cpp:121;
// SURVEY A.2-A.5); icp_solve_kernel (icp.cuh) follows and closes the iteration.
//
// A block owns a tile of 256 consecutive (Morton-ordered) source points and works in three
// phases:
//   A  one thread per point: incremental float32 transform, then the TRIANGLE-INEQUALITY
//      test.  Every point remembers its match j and a proven lower bound L on its distance
//      to every OTHER target point.  After the point has moved by delta the bound is
//      L - delta; while dist(q, j) < L - sum(delta) the match is still the unique exact
//      nearest neighbour and no search is needed (points without a target inside the gate
//      keep a proven-empty radius the same way).  Survivors are compacted, in order, into a
//      shared-memory queue.
//   B  the queue is searched by 8-lane groups (32 groups per block, contiguous chunks of
//      the queue, so the warps stay dense however few points survive): the lanes of a group
//      take the cell rows of the search ball centre-out from a precomputed offset table,
//      then walk each non-empty row's point run together (coalesced 128-byte steps),
//      tracking the best and the runner-up; the ball shrinks to the best distance (+ a
//      margin that buys the next bound) after every round of 8 rows.  A chunk's previous
//      result seeds the next query (Morton neighbours).  Exact by construction: everything
//      inside the final ball has been examined.
//   C  one thread per point again: estimator terms staged in shared memory, lane v of each
//      warp accumulates estimator value v over the warp's 32 points in a fixed order
//      (fp64 FMAs of exactly representable products), one partial row per block.
// Results (correspondence sets, d2) are bit-identical to the exhaustive search whatever the
// seeds or the skip decisions were: both only decide how much work is spent.
#pragma once
#include "icp.cuh"

namespace lc3d {

constexpr int kI2Threads = 256;
constexpr int kI2Warps = kI2Threads / 32;
constexpr int kI2Groups = kI2Threads / 8;
constexpr int kUW = 9;        // staged values per point: 7 estimator terms, d2, 1
constexpr int kTabW = 16;     // the row-offset table covers |dy|,|dz| <= kTabW
constexpr int kTabN = (2 * kTabW + 1) * (2 * kTabW + 1);
constexpr int kTabPad = (kTabN + 7) & ~7;  // padded with never-reached entries (lb = +inf)

#ifndef LC3D_I2_MINBLOCKS
#define LC3D_I2_MINBLOCKS 4
#endif

// Row-offset table: the (dy,dz) cell-row offsets of a search window ordered by a lower
// bound `lb` (cells) on the distance from the query to the row, assuming the query lies in
// the positive half of its cell along both axes (the kernel mirrors the offsets otherwise):
// offset d > 0 -> lb = d - 1, d < 0 -> lb = |d| - 0.5, d = 0 -> 0.  Entry = (packed offsets,
// bits of max(lb - slack, 0)^2 summed over the two axes).
inline void build_row_table(std::vector<int2>& tab) {
  struct E {
    int dy, dz;
    float lb2;
  };
  std::vector<E> es;
  auto lb1 = [](int d) {
    float v = d > 0 ? (float)(d - 1) : d < 0 ? (float)(-d) - 0.5f : 0.0f;
    v -= 0.01f;
    return v > 0.f ? v : 0.f;
  };
  for (int dz = -kTabW; dz <= kTabW; ++dz)
    for (int dy = -kTabW; dy <= kTabW; ++dy) {
      const float a = lb1(dy), b = lb1(dz);
      es.push_back(E{dy, dz, a * a + b * b});
    }
  std::stable_sort(es.begin(), es.end(), [](const E& a, const E& b) {
    if (a.lb2 != b.lb2) return a.lb2 < b.lb2;
    return a.dy * a.dy + a.dz * a.dz < b.dy * b.dy + b.dz * b.dz;
  });
  tab.assign(kTabPad, make_int2(128 | (128 << 8), 0x7f800000));
  for (size_t k = 0; k < es.size(); ++k) {
    int bits;
    std::memcpy(&bits, &es[k].lb2, 4);
    tab[k] = make_int2((es[k].dy + 128) | ((es[k].dz + 128) << 8), bits);
  }
}

constexpr unsigned long long kKeyNone = ((unsigned long long)0x7f800000u << 32) | 0x7fffffffu;  // (+inf, max index)

__device__ __forceinline__ float key_d2(unsigned long long key) { return __uint_as_float((unsigned)(key >> 32)); }

// candidate evaluation with runner-up tracking: best = lexicographic (d2, original index)
// minimum, sec = smallest d2 among everything else seen
__device__ __forceinline__ void consider2(const float4 p, int j, float qx, float qy, float qz,
                                          unsigned long long& bkey, int& bj, float& sec) {
  const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
  const unsigned long long key = pack_key(d2, __float_as_int(p.w));
  const bool better = key < bkey;
  sec = fminf(sec, better ? key_d2(bkey) : d2);
  bkey = better ? key : bkey;
  bj = better ? j : bj;
}

// Are all 8^3 super-cells touching the ball (radius Rc cells) around the query empty?
// Executed by an 8-lane group (lanes of other groups may pass different queries).
__device__ __forceinline__ bool coarse_ball_empty(const GridDev& g, const QueryCell& qc, bool want, float Rc,
                                                  int sub, int gbase) {
  if (!want) Rc = 0.0f;
  const float fxc = qc.fx * g.inv_xs;
  const int x0 = max((int)floorf(fxc - Rc) >> kCoarseShift, 0), x1 = min((int)floorf(fxc + Rc) >> kCoarseShift, g.cdx - 1);
  const int y0 = max((int)floorf(qc.fy - Rc) >> kCoarseShift, 0), y1 = min((int)floorf(qc.fy + Rc) >> kCoarseShift, g.cdy - 1);
  const int z0 = max((int)floorf(qc.fz - Rc) >> kCoarseShift, 0), z1 = min((int)floorf(qc.fz + Rc) >> kCoarseShift, g.cdz - 1);
  const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
  const int total = (want && nx > 0 && ny > 0 && nz > 0) ? nx * ny * nz : 0;
  bool any = false;
  for (int t = sub; t < total; t += 8) {
    const int cx = x0 + t % nx, cy = y0 + (t / nx) % ny, cz = z0 + t / (nx * ny);
    any = any || __ldg(&g.coarse_cnt[(cz * g.cdy + cy) * g.cdx + cx]) != 0u;
  }
  __syncwarp();
  return ((__ballot_sync(0xffffffffu, any) >> gbase) & 0xffu) == 0u;
}

struct CoopResult {
  unsigned long long key;  // best (d2, original index) or kKeyNone
  int j;                   // its position in the sorted target, -1 = none
  float sec;               // smallest squared distance among the other examined points
  float R;                 // final culling radius: every point within R has been examined
};

// 8-lane cooperative exact search: all points within the (shrinking) ball of radius R around
// the query are examined.  valid / q / R / mu are uniform within a group; control flow is
// warp-uniform (groups without work idle).  n_rounds / n_steps: optional statistics.
template <bool STATS>
__device__ __forceinline__ CoopResult coop_search(const GridDev& g, const int2* __restrict__ rowtab, bool valid,
                                                  float qx, float qy, float qz, float R, float mu,
                                                  unsigned& n_rounds, unsigned& n_steps) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, sub = lane & 7, gbase = lane & ~7;
  const QueryCell qc = query_cell(g, qx, qy, qz);
  const int sy = (qc.fy - (float)qc.iy) >= 0.5f ? 1 : -1;
  const int sz = (qc.fz - (float)qc.iz) >= 0.5f ? 1 : -1;
  const float inv_c2 = 1.0f / (g.c * g.c * 0.9999f);
  unsigned long long bkey = kKeyNone;
  int bj = -1;
  float sec = INFINITY;
  float Rc2 = R * R * inv_c2;  // culling radius in cells^2 (inflated: conservative)
  for (int t0 = 0; t0 < kTabPad; t0 += 8) {
    const bool more = valid && __int_as_float(__ldg(&rowtab[t0]).y) <= Rc2;
    if (!__any_sync(full, more)) break;
    uint32_t s = 0, e = 0;
    if (more) {
      const int2 ent = __ldg(&rowtab[t0 + sub]);
      if (__int_as_float(ent.y) <= Rc2) {
        const int yy = qc.iy + sy * ((ent.x & 0xff) - 128), zz = qc.iz + sz * (((ent.x >> 8) & 0xff) - 128);
        if ((unsigned)yy < (unsigned)g.dy && (unsigned)zz < (unsigned)g.dz) {
          const float gy = slab_gap(qc.fy, yy, yy), gz = slab_gap(qc.fz, zz, zz);
          const float rem = Rc2 - (gy * gy + gz * gz);
          if (rem >= 0.0f) {
            const float wx = (sqrtf(rem) + 2.0f * kCellSlack) * (float)g.xs;  // x-subcells
            const int xa = max((int)floorf(qc.fx - wx), 0), xb = min((int)floorf(qc.fx + wx), g.dx - 1);
            if (xa <= xb) {
              const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
              s = __ldg(row + xa);
              e = __ldg(row + xb + 1);
            }
          }
        }
      }
    }
    if (STATS) n_rounds += 1;
    // the group walks its non-empty rows one after the other, 8 points per step
    unsigned mg = (__ballot_sync(full, e > s) >> gbase) & 0xffu;
    while (__any_sync(full, mg != 0u)) {
      const int l = mg ? __ffs(mg) - 1 : 0;
      const uint32_t ss = __shfl_sync(full, s, gbase + l);
      uint32_t ee = __shfl_sync(full, e, gbase + l);
      if (!mg) ee = ss;
      mg &= mg - 1u;
      for (uint32_t j = ss + sub; j < ee; j += 8) {
        consider2(__ldg(&g.pts[j]), (int)j, qx, qy, qz, bkey, bj, sec);
        if (STATS) n_steps += 1;
      }
    }
    // shrink the ball to the best distance (+ margin) seen by the group so far
    float bd = key_d2(bkey);
    bd = fminf(bd, __shfl_xor_sync(full, bd, 4));
    bd = fminf(bd, __shfl_xor_sync(full, bd, 2));
    bd = fminf(bd, __shfl_xor_sync(full, bd, 1));
    const float rn = sqrtf(bd) * 1.00001f + mu;  // +inf while nothing has been found
    if (rn < R) {
      R = rn;
      Rc2 = R * R * inv_c2;
    }
  }
  // merge the 8 lanes' (best, runner-up) pairs
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    const unsigned long long ok = __shfl_xor_sync(full, bkey, o);
    const int oj = __shfl_xor_sync(full, bj, o);
    const float os = __shfl_xor_sync(full, sec, o);
    const bool take = ok < bkey;
    const unsigned long long loser = take ? bkey : ok;
    sec = fminf(fminf(sec, os), ok == bkey ? INFINITY : key_d2(loser));
    bkey = take ? ok : bkey;
    bj = take ? oj : bj;
  }
  CoopResult r;
  r.key = bkey;
  r.j = bj;
  r.sec = sec;
  r.R = R;
  return r;
}

// index pair (ia, ib) of the staged values whose product lane v accumulates
template <int MODE>
__device__ __forceinline__ void estimator_pair(int v, int& ia, int& ib) {
  ia = ib = 8;  // 1 * 1: the correspondence count
  if (MODE == LC3D_ICP_POINT_TO_PLANE) {
    // staged: J[0..5], r, d2, 1.  [0..20] J^T J upper triangle, [21..26] J^T r, [27] d2, [28] 1
    if (v < 21) {
      const int a = v < 6 ? 0 : v < 11 ? 1 : v < 15 ? 2 : v < 18 ? 3 : v < 20 ? 4 : 5;
      const int base = a == 0 ? 0 : a == 1 ? 6 : a == 2 ? 11 : a == 3 ? 15 : a == 4 ? 18 : 20;
      ia = a;
      ib = a + (v - base);
    } else if (v < 27) {
      ia = v - 21;
      ib = 6;
    } else if (v == 27) {
      ia = 7;
    }
  } else {
    // staged: s[0..2], d[0..2], -, d2, 1.  [0..2] s, [3..5] d, [6..14] d s^T, [15] d2, [16] 1
    if (v < 6) {
      ia = v;
    } else if (v < 15) {
      ia = 3 + (v - 6) / 3;
      ib = (v - 6) % 3;
    } else if (v == 15) {
      ia = 7;
    }
  }
}

template <int MODE, bool STATS>
__global__ void __launch_bounds__(kI2Threads, LC3D_I2_MINBLOCKS)
    icp_iter2_kernel(IcpState* __restrict__ st, const __grid_constant__ IcpConfig cfg,
                     const __grid_constant__ GridDev g, float4* __restrict__ X, int2* __restrict__ MB, int n,
                     double* __restrict__ partials, int32_t* __restrict__ dump_idx, float* __restrict__ dump_d2,
                     const int2* __restrict__ rowtab) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  const unsigned full = 0xffffffffu;
  __shared__ float sT[16];
  __shared__ int s_flags[2];
  __shared__ float s_qx[kI2Threads], s_qy[kI2Threads], s_qz[kI2Threads];
  __shared__ float s_d2[kI2Threads], s_L[kI2Threads], s_mu[kI2Threads];
  __shared__ int s_j[kI2Threads];
  __shared__ unsigned short s_queue[kI2Threads], s_queue2[kI2Threads];
  __shared__ int s_wcnt[kI2Warps];
  __shared__ int s_cnt2;
  __shared__ double s_U[kI2Warps][32][kUW];
  __shared__ double s_part[kI2Warps][32];
  pdl_wait();  // the previous solve kernel's pose / done flag
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) {
    s_flags[0] = st->done;
    s_flags[1] = st->iter;
    s_cnt2 = 0;
  }
  if (tid < 16) sT[tid] = st->T[tid];
  __syncthreads();
  if (s_flags[0]) return;
  const int iter = s_flags[1];
  SearchStats* stats = (STATS && cfg.stats) ? cfg.stats + iter : nullptr;
  if (stats && tid == 0) {
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    atomicMin(&stats->c[11], t0);
  }
  // ---- phase A: transform + triangle-inequality test -----------------------------------
  const int i = blockIdx.x * kI2Threads + tid;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    bool active = i < n;
    if (active) q = X[i];
    int2 mb = make_int2(-1, 0);
    if (active && iter > 0) mb = MB[i];
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
    int mj = active ? mb.x : -1;
    const float Lb = fmaxf(__int_as_float(mb.y) - delta, 0.0f);
    float d2m = INFINITY;
    bool need = active && g.n > 0;
    if (need) {
      if (mj >= 0) {
        const float4 p = __ldg(&g.pts[mj]);
        d2m = dist2_exact(q.x, q.y, q.z, p.x, p.y, p.z);
        if (sqrtf(d2m) * 1.00002f < Lb) need = false;  // still the unique nearest neighbour
      } else if (Lb * 0.9999f > cfg.gate_dist) {
        need = false;  // still nothing within the gate
      }
    }
    s_qx[tid] = q.x;
    s_qy[tid] = q.y;
    s_qz[tid] = q.z;
    s_j[tid] = mj;
    s_d2[tid] = d2m;
    s_L[tid] = Lb;
    s_mu[tid] = iter == 0 ? cfg.mu0 : fminf(fmaxf(cfg.mu_kappa * delta, cfg.mu_min), cfg.mu_max);
    // ordered compaction of the survivors into the block queue
    const unsigned nm = __ballot_sync(full, need);
    if (lane == 0) s_wcnt[w] = __popc(nm);
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int ww = 0; ww < kI2Warps; ++ww) base += ww < w ? s_wcnt[ww] : 0;
    if (need) s_queue[base + __popc(nm & ((1u << lane) - 1u))] = (unsigned short)tid;
  }
  int S = 0;
#pragma unroll
  for (int ww = 0; ww < kI2Warps; ++ww) S += s_wcnt[ww];
  __syncthreads();
  // ---- phase B: cooperative search of the queue ----------------------------------------
  if (S > 0) {
    const int grp = lane >> 3, sub = lane & 7, gbase = lane & ~7;
    const int G = w * 4 + grp;
    const int C = (S + kI2Groups - 1) / kI2Groups;  // queue entries per group (contiguous chunk)
    int last_j = -1;
    unsigned n_rounds = 0, n_steps = 0, n_fallback = 0;
    for (int k = 0; k < C; ++k) {
      const int e = G * C + k;
      const bool valid = e < S;
      const int slot = valid ? s_queue[e] : 0;
      const float qx = s_qx[slot], qy = s_qy[slot], qz = s_qz[slot], mu = s_mu[slot];
      const int sj = s_j[slot];
      float sd2 = s_d2[slot];  // distance to the previous match (inf if none)
      if (valid && last_j >= 0 && last_j != sj) {  // second seed: the chunk's previous result
        const float4 p = __ldg(&g.pts[last_j]);
        sd2 = fminf(sd2, dist2_exact(qx, qy, qz, p.x, p.y, p.z));
      }
      float R = cfg.r_cap;  // +inf without a gate
      const bool seeded = sd2 < INFINITY;
      if (seeded) R = fminf(R, sqrtf(sd2) * 1.00001f + mu);
      const float Rc = R * g.inv_c * 1.0001f + 0.01f;  // cells
      const bool coop = valid && Rc <= cfg.tab_wmax;
      bool search = coop;
      if (__any_sync(full, coop && !seeded)) {
        // nothing known: is there anything at all within reach?  (non-overlap regions)
        const QueryCell qc = query_cell(g, qx, qy, qz);
        const bool empty = coarse_ball_empty(g, qc, coop && !seeded, Rc, sub, gbase);
        if (coop && !seeded && empty) search = false;
      }
      const CoopResult r = coop_search<STATS>(g, rowtab, search, qx, qy, qz, R, mu, n_rounds, n_steps);
      if (valid && sub == 0) {
        if (coop) {
          // a best beyond the final ball is not proven nearest: report "nothing within R"
          const float bd2 = key_d2(r.key);
          const bool found = search && r.j >= 0 && sqrtf(bd2) * 1.00001f <= r.R;
          s_j[slot] = found ? r.j : -1;
          s_d2[slot] = found ? bd2 : INFINITY;
          s_L[slot] = (found ? fminf(sqrtf(r.sec), r.R) : r.R) * 0.9999f;
        } else {
          s_queue2[atomicAdd(&s_cnt2, 1)] = (unsigned short)slot;  // ball too wide for the row table
        }
      }
      if (STATS && valid && !coop && sub == 0) n_fallback += 1;
      if (coop && search && r.j >= 0) last_j = r.j;
    }
    if (STATS && stats) {
      for (int o = 16; o > 0; o >>= 1) {
        n_rounds = max(n_rounds, __shfl_xor_sync(full, n_rounds, o));
        n_steps = max(n_steps, __shfl_xor_sync(full, n_steps, o));
        n_fallback += __shfl_xor_sync(full, n_fallback, o);
      }
      if (lane == 0) {
        stat_add(stats, 5, n_fallback);
        stat_add(stats, 6, n_steps);
        stat_add(stats, 7, n_rounds);
      }
      if (tid == 0) {
        stat_add(stats, 4, (unsigned long long)S);
        stat_add(stats, 0, (unsigned long long)min(kI2Threads, n - blockIdx.x * kI2Threads));
      }
    }
    __syncthreads();
    // rare: balls wider than the row table (huge gates, no gate): warp-cooperative ring search
    const int S2 = s_cnt2;
    for (int e = w; e < S2; e += kI2Warps) {
      const int slot = s_queue2[e];
      const float qx = s_qx[slot], qy = s_qy[slot], qz = s_qz[slot];
      Best b;
      b.d2 = cfg.gate_ext;
      b.j = -1;
      b.oi = 0x7fffffff;
      const int sj = s_j[slot];
      if (sj >= 0) consider(__ldg(&g.pts[sj]), sj, qx, qy, qz, b);
      nn_phase2_warp(g, qx, qy, qz, b);
      if (lane == 0) {
        s_j[slot] = b.j;
        s_d2[slot] = b.j >= 0 ? b.d2 : INFINITY;
        // nothing learnt about the runner-up; without a match nothing lies within the extended gate
        s_L[slot] = b.j >= 0 ? 0.0f : cfg.r_cap * 0.9999f;
      }
    }
    __syncthreads();
  }
  // ---- phase C: state write-back + estimator sums --------------------------------------
  const int j = s_j[tid];
  const float d2 = s_d2[tid];
  if (i < n) MB[i] = make_int2(j, __float_as_int(s_L[tid]));
  const bool has = j >= 0 && d2 <= cfg.gate;
  float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has) d = __ldg(&g.pts[j]);
  if (dump_idx && iter == cfg.dump_iteration && i < n) {
    const int oi = __float_as_int(q.w);
    dump_idx[oi] = has ? __float_as_int(d.w) : -1;
    dump_d2[oi] = has ? d2 : INFINITY;
  }
  double acc = 0.0;
  const unsigned hm = __ballot_sync(full, has);
  if (hm) {
    double* u = &s_U[w][lane][0];
    if (MODE == LC3D_ICP_POINT_TO_PLANE) {
      float J[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, r = 0.f;
      if (has) {
        const float4 nn = __ldg(&g.nrm[j]);
        if (finite3(nn.x, nn.y, nn.z)) {
          // float32 products widened to double, as TransformationEstimationPointToPlaneLLS
          J[0] = nn.z * q.y - nn.y * q.z;
          J[1] = nn.x * q.z - nn.z * q.x;
          J[2] = nn.y * q.x - nn.x * q.y;
          J[3] = nn.x;
          J[4] = nn.y;
          J[5] = nn.z;
          r = nn.x * d.x + nn.y * d.y + nn.z * d.z - nn.x * q.x - nn.y * q.y - nn.z * q.z;
        }
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) u[k] = (double)J[k];
      u[6] = (double)r;
    } else {
      u[0] = has ? (double)q.x : 0.0;
      u[1] = has ? (double)q.y : 0.0;
      u[2] = has ? (double)q.z : 0.0;
      u[3] = (double)d.x;
      u[4] = (double)d.y;
      u[5] = (double)d.z;
      u[6] = 0.0;
    }
    u[7] = has ? (double)d2 : 0.0;
    u[8] = has ? 1.0 : 0.0;
    __syncwarp();
    int ia, ib;
    estimator_pair<MODE>(lane, ia, ib);
    const double* ua = &s_U[w][0][ia];
    const double* ub = &s_U[w][0][ib];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc = __fma_rn(ua[k * kUW], ub[k * kUW], acc);
  }
  s_part[w][lane] = acc;
  __syncthreads();
  if (w == 0 && lane < NV) {
    double s = 0.0;
#pragma unroll
    for (int ww = 0; ww < kI2Warps; ++ww) s += s_part[ww][lane];
    partials[(size_t)lane * gridDim.x + blockIdx.x] = s;
  }
}

}  // namespace lc3d
