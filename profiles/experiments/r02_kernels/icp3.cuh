// icp3.cuh — the ICP iteration kernel (third generation): multi-pass compacted search.
//
// One launch = transformCloud + determineCorrespondences + the estimator's sums of ONE
// iteration of pcl::IterativeClosestPoint::align (pcl_tools/fine_registration.cpp:121;
// SURVEY A.2-A.5); icp_solve_kernel (icp.cuh) follows and closes the iteration.
//
// A block owns a tile of 256 consecutive (Morton-ordered) source points:
//   A   one thread per point: incremental float32 transform; the previous match is
//       re-evaluated (an exact upper bound on the nearest-neighbour distance); points
//       whose cached candidate list is still provably complete (triangle inequality, see
//       below) or that provably have nothing within the gate are finished here.  The rest
//       is compacted into a shared-memory queue.
//   B   the queue is searched in PASSES of growing reach, one thread per queued point,
//       with the unresolved points re-compacted between passes so that warps stay dense and
//       the cheap majority never waits for the expensive few:
//         pass 0   the 2x2 cell rows nearest to the point, +-1 cell along x
//         pass 1   the 4x4 nearest rows, +-2 cells along x
//         pass 2   every row that intersects the search ball (after an occupancy test that
//                  rejects points with nothing in reach), then, for balls wider than the
//                  row table, the warp-cooperative ring search.
//       All passes walk the same precomputed centre-out row table (icp2.cuh) and stop at
//       the first row whose lower bound exceeds the running best; a pass resolves a point
//       when the rows and the x-range it examined cover the ball of the best distance —
//       exact by construction.
//   C   one thread per point: estimator terms staged in shared memory, lane v of each warp
//       accumulates estimator value v over the warp's 32 points in a fixed order (fp64
//       FMAs of exactly representable products), one partial row per block.
// Candidate lists (temporal coherence): a search may also record every target point within
// radius d_best + mu of the point (at most kListK entries) together with the proven radius
// Lb of that record.  After the point has moved by delta, every target point within
// Lb - sum(delta) is still in the list, so while the nearest LIST entry is closer than that
// bound it is the exact nearest neighbour and no search is needed.
#pragma once
#include "icp2.cuh"

namespace lc3d {

#ifndef LC3D_I3_THREADS
#define LC3D_I3_THREADS 256
#endif
constexpr int kI3Threads = LC3D_I3_THREADS;
constexpr int kI3Warps = kI3Threads / 32;
#ifndef LC3D_LIST_K
#define LC3D_LIST_K 8
#endif
constexpr int kListK = LC3D_LIST_K;  // candidate-list capacity per source point

#ifndef LC3D_I3_MINBLOCKS
#define LC3D_I3_MINBLOCKS 4
#endif

// Thread-per-query walk over the centre-out row table: rows [0, t_end) (stopping at the first
// row whose lower bound exceeds the running best), x-range = the ball clipped to +-xw cells
// around the query's cell.  b: running best, b.d2 = bound (gate or seed distance) on entry.
// Returns true when the examined region covers the ball of the final best distance, i.e. b is
// the exact nearest neighbour (or, with b.j < 0, nothing lies within sqrt(b.d2)).
// All 32 lanes must call (warp-uniform row loop, lanes mask themselves out).
// COLLECT: additionally record every examined point with d2 <= collect_r2 into the caller's
// list (lst[k * lst_stride], at most kListK), counting in *lst_n; *lst_drop receives the
// smallest squared distance among in-radius points that did not fit.
template <bool COLLECT>
__device__ __forceinline__ bool walk_rows(const GridDev& g, const int2* __restrict__ rowtab, bool act, float qx,
                                          float qy, float qz, int t_end, int xw, Best& b, float collect_r2,
                                          int* lst, int lst_stride, int* lst_n, float* lst_drop,
                                          unsigned* n_rows, unsigned* n_cand) {
  const unsigned full = 0xffffffffu;
  const QueryCell qc = query_cell(g, qx, qy, qz);
  const int sy = (qc.fy - (float)qc.iy) >= 0.5f ? 1 : -1;
  const int sz = (qc.fz - (float)qc.iz) >= 0.5f ? 1 : -1;
  const float inv_c2 = 1.0f / (g.c * g.c * 0.9999f);
  const int xlo = max(qc.ix - xw * g.xs, 0), xhi = min(qc.ix + xw * g.xs + g.xs - 1, g.dx - 1);
  int cnt = 0;
  float drop = INFINITY;
  int t = 0;
  float bc2 = 0.0f;
  for (; t < t_end; ++t) {
    const int2 ent = __ldg(&rowtab[t]);  // warp-uniform address
    bc2 = (COLLECT ? collect_r2 : b.d2) * inv_c2;  // ball radius^2 in cells (inflated: conservative)
    const bool go = act && __int_as_float(ent.y) <= bc2;
    if (!__any_sync(full, go)) break;
    if (go) {
      const int yy = qc.iy + sy * ((ent.x & 0xff) - 128), zz = qc.iz + sz * (((ent.x >> 8) & 0xff) - 128);
      if ((unsigned)yy < (unsigned)g.dy && (unsigned)zz < (unsigned)g.dz) {
        const float gy = slab_gap(qc.fy, yy, yy), gz = slab_gap(qc.fz, zz, zz);
        const float rem = bc2 - (gy * gy + gz * gz);
        if (rem >= 0.0f) {
          const float wx = (sqrtf(rem) + 2.0f * kCellSlack) * (float)g.xs;  // x-subcells
          const int xa = max((int)floorf(qc.fx - wx), xlo), xb = min((int)floorf(qc.fx + wx), xhi);
          if (xa <= xb) {
            const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
            const uint32_t s = __ldg(row + xa), e = __ldg(row + xb + 1);
            if (n_rows) {
              *n_rows += 1;
              *n_cand += e - s;
            }
            for (uint32_t j = s; j < e; ++j) {
              const float4 p = __ldg(&g.pts[j]);
              if (COLLECT) {
                const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
                if (d2 <= collect_r2) {
                  if (cnt < kListK)
                    lst[cnt * lst_stride] = (int)j;
                  else
                    drop = fminf(drop, d2);
                  ++cnt;
                }
              } else {
                consider(p, (int)j, qx, qy, qz, b);
              }
            }
          }
        }
      }
    }
  }
  if (COLLECT) {
    *lst_n = min(cnt, kListK);
    *lst_drop = drop;
  }
  // rows: every row not examined has a lower bound beyond the final ball
  const float fin2 = (COLLECT ? collect_r2 : b.d2) * inv_c2;
  const bool rows_ok = t < kTabPad && __int_as_float(__ldg(&rowtab[min(t, kTabPad - 1)]).y) > fin2;
  // x: the final ball fits into the +-xw window (or the window ends at the grid border)
  const float wf = (sqrtf(fin2) + 2.0f * kCellSlack) * (float)g.xs;
  const int need_lo = (int)floorf(qc.fx - wf), need_hi = (int)floorf(qc.fx + wf);
  const bool x_ok = (need_lo >= qc.ix - xw * g.xs || xlo == 0) && (need_hi <= qc.ix + xw * g.xs + g.xs - 1 || xhi == g.dx - 1);
  return rows_ok && x_ok;
}

// thread-per-query: are all super-cells touching the ball (radius Rc cells) empty?
__device__ __forceinline__ bool coarse_ball_empty_thread(const GridDev& g, float qx, float qy, float qz, float Rc) {
  const QueryCell qc = query_cell(g, qx, qy, qz);
  const float fxc = qc.fx * g.inv_xs;
  const int x0 = max((int)floorf(fxc - Rc) >> kCoarseShift, 0), x1 = min((int)floorf(fxc + Rc) >> kCoarseShift, g.cdx - 1);
  const int y0 = max((int)floorf(qc.fy - Rc) >> kCoarseShift, 0), y1 = min((int)floorf(qc.fy + Rc) >> kCoarseShift, g.cdy - 1);
  const int z0 = max((int)floorf(qc.fz - Rc) >> kCoarseShift, 0), z1 = min((int)floorf(qc.fz + Rc) >> kCoarseShift, g.cdz - 1);
  bool any = false;
  for (int cz = z0; cz <= z1; ++cz)
    for (int cy = y0; cy <= y1; ++cy)
      for (int cx = x0; cx <= x1; ++cx) any = any || __ldg(&g.coarse_cnt[(cz * g.cdy + cy) * g.cdx + cx]) != 0u;
  return !any;
}

struct Icp3Lists {
  int* lst;            // [kListK][n] sorted-target positions of the cached candidates
  unsigned char* cnt;  // [n] entries in use
};

template <int MODE, bool STATS>
__global__ void __launch_bounds__(kI3Threads, LC3D_I3_MINBLOCKS)
    icp_iter3_kernel(IcpState* __restrict__ st, const __grid_constant__ IcpConfig cfg,
                     const __grid_constant__ GridDev g, float4* __restrict__ X, int2* __restrict__ MB,
                     const Icp3Lists lists, int n, double* __restrict__ partials, int32_t* __restrict__ dump_idx,
                     float* __restrict__ dump_d2, const int2* __restrict__ rowtab) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  const unsigned full = 0xffffffffu;
  __shared__ float sT[16];
  __shared__ int s_flags[2];
  __shared__ float s_qx[kI3Threads], s_qy[kI3Threads], s_qz[kI3Threads];
  __shared__ float s_d2[kI3Threads], s_L[kI3Threads], s_mu[kI3Threads];
  __shared__ int s_j[kI3Threads], s_oi[kI3Threads];
  __shared__ unsigned char s_nlst[kI3Threads];  // list entries valid after this iteration
  __shared__ unsigned short s_q0[kI3Threads], s_q1[kI3Threads], s_q2[kI3Threads], s_q3[kI3Threads], s_ql[kI3Threads];
  __shared__ int s_wcnt[kI3Warps];
  __shared__ int s_n1, s_n2, s_n3, s_nl;
  __shared__ double s_U[kI3Warps][32][kUW];
  __shared__ double s_part[kI3Warps][32];
  pdl_wait();  // the previous solve kernel's pose / done flag
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) {
    s_flags[0] = st->done;
    s_flags[1] = st->iter;
    s_n1 = s_n2 = s_n3 = s_nl = 0;
  }
  if (tid < 16) sT[tid] = st->T[tid];
  __syncthreads();
  if (s_flags[0]) return;
  const int iter = s_flags[1];
  SearchStats* stats = (STATS && cfg.stats) ? cfg.stats + iter : nullptr;
  // ---- phase A ---------------------------------------------------------------------------
  const int i = blockIdx.x * kI3Threads + tid;
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
    const int mj = active ? mb.x : -1;
    const float Lb = fmaxf(__int_as_float(mb.y) - delta, 0.0f);
    Best b;
    b.d2 = cfg.gate_ext;
    b.j = -1;
    b.oi = 0x7fffffff;
    bool need = active && g.n > 0;
    int nl = 0;
    if (need) {
      nl = (lists.cnt && iter > 0) ? (int)lists.cnt[i] : 0;
      if (nl > 0) {
        // cached candidates: every target point within Lb of the point is among them.  Loads are
        // issued in batches of 8 (indices, then points) so that a list costs two memory round
        // trips per batch instead of two per entry.
        for (int k0 = 0; k0 < nl; k0 += 8) {
          int jj[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) jj[k] = k0 + k < nl ? lists.lst[(size_t)(k0 + k) * n + i] : -1;
          float4 pp[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) pp[k] = jj[k] >= 0 ? __ldg(&g.pts[jj[k]]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (jj[k] >= 0) consider(pp[k], jj[k], q.x, q.y, q.z, b);
        }
        if (b.j >= 0 && sqrtf(b.d2) * 1.00002f < Lb) need = false;  // exact nearest neighbour
      } else if (mj >= 0) {
        consider(__ldg(&g.pts[mj]), mj, q.x, q.y, q.z, b);  // previous match: an upper bound
      } else if (Lb * 0.9999f > cfg.gate_dist) {
        need = false;  // still nothing within the gate
      }
    }
    float Lkeep = Lb;
    if (need && b.j < 0) {
      // no candidate at all: the dilated occupancy may prove that nothing lies within the gate
      const QueryCell qc = query_cell(g, q.x, q.y, q.z);
      if (occ_proves_empty(g, qc.ix, qc.iy, qc.iz)) {
        need = false;
        Lkeep = ((float)g.occ_r - 0.01f) * g.c;
      }
    }
    s_qx[tid] = q.x;
    s_qy[tid] = q.y;
    s_qz[tid] = q.z;
    s_j[tid] = b.j;
    s_d2[tid] = b.d2;
    s_oi[tid] = b.oi;
    s_L[tid] = need ? 0.0f : Lkeep;
    s_nlst[tid] = (unsigned char)(need ? 0 : nl);
    // list margin: motion still to come is a small multiple of the last motion
    const float mu = cfg.mu_kappa * delta;
    s_mu[tid] = (iter > 0 && mu <= cfg.mu_max) ? fmaxf(mu, cfg.mu_min) : -1.0f;  // < 0: no list
    const unsigned nm = __ballot_sync(full, need);
    if (lane == 0) s_wcnt[w] = __popc(nm);
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int ww = 0; ww < kI3Warps; ++ww) base += ww < w ? s_wcnt[ww] : 0;
    if (need) s_q0[base + __popc(nm & ((1u << lane) - 1u))] = (unsigned short)tid;
  }
  int S = 0;
#pragma unroll
  for (int ww = 0; ww < kI3Warps; ++ww) S += s_wcnt[ww];
  __syncthreads();
  // ---- phase B: passes of growing reach ----------------------------------------------------
  if (S > 0) {
    unsigned n_rows = 0, n_cand = 0;
    // append `slot` to a shared queue, one atomic per warp
    auto push = [&](bool p, unsigned short* qu, int* cnt, int slot) {
      const unsigned m = __ballot_sync(full, p);
      if (m) {
        int b0 = 0;
        if (lane == __ffs(m) - 1) b0 = atomicAdd(cnt, __popc(m));
        b0 = __shfl_sync(full, b0, __ffs(m) - 1);
        if (p) qu[b0 + __popc(m & ((1u << lane) - 1u))] = (unsigned short)slot;
      }
    };
    auto load_slot = [&](int slot, float& qx, float& qy, float& qz, Best& b) {
      qx = s_qx[slot];
      qy = s_qy[slot];
      qz = s_qz[slot];
      b.d2 = s_d2[slot];
      b.j = s_j[slot];
      b.oi = s_oi[slot];
    };
    auto store_slot = [&](int slot, const Best& b) {
      s_d2[slot] = b.d2;
      s_j[slot] = b.j;
      s_oi[slot] = b.oi;
    };
    // resolved: final state of the slot.  A best beyond the gate_ext bound never enters b.
    auto resolve = [&](int slot, const Best& b) {
      store_slot(slot, b);
      // nothing found: nothing lies within r_cap.  found: nothing learnt about the others
      s_L[slot] = b.j >= 0 ? 0.0f : cfg.r_cap * 0.9999f;
    };
    // pass 0: 2x2 rows, +-1 cell
    if (w * 32 < S) {
      const bool act = tid < S;
      const int slot = act ? s_q0[tid] : 0;
      float qx, qy, qz;
      Best b;
      load_slot(slot, qx, qy, qz, b);
      const bool ok = walk_rows<false>(g, rowtab, act, qx, qy, qz, 4, 1, b, 0.f, nullptr, 0, nullptr, nullptr,
                                       STATS ? &n_rows : nullptr, STATS ? &n_cand : nullptr);
      if (act) {
        if (ok) resolve(slot, b); else store_slot(slot, b);
      }
      push(act && !ok, s_q1, &s_n1, slot);
      push(act && ok && b.j >= 0 && s_mu[slot] >= 0.0f, s_ql, &s_nl, slot);
    }
    __syncthreads();
    const int S1 = s_n1;
    // pass 1: 4x4 rows, +-2 cells
    if (w * 32 < S1) {
      const bool act = tid < S1;
      const int slot = act ? s_q1[tid] : 0;
      float qx, qy, qz;
      Best b;
      load_slot(slot, qx, qy, qz, b);
      const bool ok = walk_rows<false>(g, rowtab, act, qx, qy, qz, 16, 2, b, 0.f, nullptr, 0, nullptr, nullptr,
                                       STATS ? &n_rows : nullptr, STATS ? &n_cand : nullptr);
      if (act) {
        if (ok) resolve(slot, b); else store_slot(slot, b);
      }
      push(act && !ok, s_q2, &s_n2, slot);
      push(act && ok && b.j >= 0 && s_mu[slot] >= 0.0f, s_ql, &s_nl, slot);
    }
    __syncthreads();
    const int S2 = s_n2;
#ifdef LC3D_I3_COOP_PASS2
    // pass 2, cooperative variant: one WARP per point
    for (int e = w; e < S2; e += kI3Warps) {
      const int slot = s_q2[e];
      float qx, qy, qz;
      Best b;
      load_slot(slot, qx, qy, qz, b);
      if (!nn_ball_warp(g, qx, qy, qz, b)) nn_phase2_warp(g, qx, qy, qz, b);
      if (lane == 0) resolve(slot, b);
      if (lane == 0 && b.j >= 0 && s_mu[slot] >= 0.0f) s_ql[atomicAdd(&s_nl, 1)] = (unsigned short)slot;
    }
    const int S3 = 0;
    __syncthreads();
#else
    // pass 2: every row of the ball (neighbouring points have similar balls, so the lanes of a
    // warp stay roughly in step); balls wider than the row table go to the ring search
    if (w * 32 < S2) {
      const bool act = tid < S2;
      const int slot = act ? s_q2[tid] : 0;
      float qx, qy, qz;
      Best b;
      load_slot(slot, qx, qy, qz, b);
      const float Rc = sqrtf(b.d2) * g.inv_c * 1.0001f + 0.01f;  // ball radius in cells
      const bool walk = act && Rc <= cfg.tab_wmax;
      bool ok = false;
      if (__any_sync(full, walk)) {
        const bool okw = walk_rows<false>(g, rowtab, walk, qx, qy, qz, kTabN, 1 << 20, b, 0.f, nullptr, 0, nullptr,
                                          nullptr, STATS ? &n_rows : nullptr, STATS ? &n_cand : nullptr);
        if (walk) ok = okw;
      }
      if (act) {
        if (ok) resolve(slot, b); else store_slot(slot, b);
      }
      push(act && !ok, s_q3, &s_n3, slot);
      push(act && ok && b.j >= 0 && s_mu[slot] >= 0.0f, s_ql, &s_nl, slot);
    }
    __syncthreads();
    const int S3 = s_n3;
    for (int e = w; e < S3; e += kI3Warps) {
      const int slot = s_q3[e];
      float qx, qy, qz;
      Best b;
      load_slot(slot, qx, qy, qz, b);
      nn_phase2_warp(g, qx, qy, qz, b);
      if (lane == 0) resolve(slot, b);
    }
#endif
    // candidate lists of the points resolved above (those with a motion small enough to pay)
    const int SL = s_nl;
    if (lists.cnt && w * 32 < SL) {
      const bool act = tid < SL;
      const int slot = act ? s_ql[tid] : 0;
      float qx, qy, qz;
      Best b;
      load_slot(slot, qx, qy, qz, b);
      const float rl = fminf(sqrtf(b.d2) * 1.00001f + fmaxf(s_mu[slot], 0.0f), cfg.r_cap);
      const int gi = blockIdx.x * kI3Threads + slot;
      int cnt = 0;
      float drop = INFINITY;
      const bool ok = walk_rows<true>(g, rowtab, act && rl * g.inv_c * 1.0001f + 0.01f <= cfg.tab_wmax, qx, qy, qz,
                                      kTabN, 1 << 20, b, rl * rl, lists.lst + gi, n, &cnt, &drop, nullptr, nullptr);
      if (act && ok) {
        s_nlst[slot] = (unsigned char)cnt;
        s_L[slot] = fminf(rl, sqrtf(drop)) * 0.9999f;
      }
    }
    if (STATS && stats) {
      for (int o = 16; o > 0; o >>= 1) {
        n_rows = max(n_rows, __shfl_xor_sync(full, n_rows, o));
        n_cand = max(n_cand, __shfl_xor_sync(full, n_cand, o));
      }
      if (lane == 0) {
        stat_add(stats, 6, n_cand);
        stat_add(stats, 7, n_rows);
      }
      if (tid == 0) {
        stat_add(stats, 0, (unsigned long long)min(kI3Threads, n - blockIdx.x * kI3Threads));
        stat_add(stats, 1, (unsigned long long)S);
        stat_add(stats, 2, (unsigned long long)S1);
        stat_add(stats, 3, (unsigned long long)S2);
        stat_add(stats, 5, (unsigned long long)S3);
        stat_add(stats, 4, (unsigned long long)SL);
      }
    }
    __syncthreads();
  }
  // ---- phase C: state write-back + estimator sums ------------------------------------------
  const int j = s_j[tid];
  const float d2 = s_d2[tid];
  if (i < n) {
    MB[i] = make_int2(j, __float_as_int(s_L[tid]));
    if (lists.cnt) lists.cnt[i] = s_nlst[tid];  // a list is only valid together with the bound written with it
  }
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
    for (int ww = 0; ww < kI3Warps; ++ww) s += s_part[ww][lane];
    partials[(size_t)lane * gridDim.x + blockIdx.x] = s;
  }
}

}  // namespace lc3d
