// search.cuh — exact nearest-neighbour search on the uniform grid (device functions).
//
// Replaces the FLANN kd-tree queries PCL issues: exact 1-NN under the
// float32 squared distance ((dx^2)+dy^2)+dz^2 evaluated without FMA, optional gate
// d2 <= thr, ties resolved to the lower ORIGINAL point index.
//
// Two phases per query:
//   phase 1 (one thread per query): the 3x3x3 cell block around the query = 9 contiguous
//     point runs (x-fastest cell order), centre row first, rows culled by their y/z
//     slab distance against the running best.  Proves the result exact whenever the
//     best distance is within the block's guaranteed radius — the common case once
//     ICP is near convergence.
//   phase 2 (warp-cooperative): queries that phase 1 could not prove are handed, one at
//     a time, to the whole warp: rings of 8^3-cell super-cells are enumerated in
//     parallel, empty / too-far super-cells rejected on the small occupancy table, the
//     64 cell rows of each surviving super-cell are split over the lanes, and the lane
//     bests are min-reduced with __shfl.  Handles large distances, the unbounded search
//     of getFitnessScore, and queries outside the grid.
#pragma once
#include "grid.cuh"

namespace lc3d {

struct Best {
  float d2;  // running best squared distance (initialised to the gate or +inf)
  int j;     // position in the sorted target arrays, -1 = none
  int oi;    // original index of that target point (tie-break key)
};

// (d2, original index) packed so that one unsigned 64-bit compare orders candidates by
// distance, ties by lower index (d2 >= 0, so the float bit pattern is monotonic).
__device__ __forceinline__ unsigned long long pack_key(float d2, int oi) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned)oi;
}

__device__ __forceinline__ void consider(const float4 p, int j, float qx, float qy, float qz,
                                         Best& b) {
  const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
  const int oi = __float_as_int(p.w);
  const bool better = pack_key(d2, oi) < pack_key(b.d2, b.oi);
  b.d2 = better ? d2 : b.d2;
  b.j = better ? j : b.j;
  b.oi = better ? oi : b.oi;
}

__device__ __forceinline__ void scan_run(const float4* __restrict__ pts, uint32_t s, uint32_t e,
                                         float qx, float qy, float qz, Best& b) {
  for (uint32_t j = s; j < e; ++j) consider(__ldg(&pts[j]), (int)j, qx, qy, qz, b);
}

// Distance (in cells, >= 0) from coordinate f to the slab [lo, hi] of cell indices.
__device__ __forceinline__ float slab_gap(float f, int lo, int hi) {
  float a = (float)lo - f, b = f - (float)(hi + 1);
  return fmaxf(fmaxf(a, b) - kCellSlack, 0.0f);
}

struct QueryCell {
  float fx, fy, fz;
  int ix, iy, iz;
};
__device__ __forceinline__ QueryCell query_cell(const GridDev& g, float qx, float qy, float qz) {
  QueryCell c;
  c.fx = cell_coord(qx, g.ox, g.inv_cx);  // in x-SUBcells
  c.fy = cell_coord(qy, g.oy, g.inv_c);
  c.fz = cell_coord(qz, g.oz, g.inv_c);
  // clamp so that int conversion and later arithmetic cannot overflow
  c.fx = fminf(fmaxf(c.fx, -1.0e6f), 1.0e6f);
  c.fy = fminf(fmaxf(c.fy, -1.0e6f), 1.0e6f);
  c.fz = fminf(fmaxf(c.fz, -1.0e6f), 1.0e6f);
  c.ix = (int)floorf(c.fx);
  c.iy = (int)floorf(c.fy);
  c.iz = (int)floorf(c.fz);
  return c;
}

// Phase 1.  Returns true when the result in `b` is proven exact.
__device__ __forceinline__ bool nn_phase1(const GridDev& g, const QueryCell& qc, float qx, float qy,
                                          float qz, Best& b) {
  const float c2 = g.c * g.c * 0.9999f;  // conservative scale for cell-unit bounds
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    // centre row first, then the 4 edge-adjacent rows, then the 4 diagonal rows
    const int oy = (t == 1 || t == 5 || t == 7) ? -1 : (t == 2 || t == 6 || t == 8) ? 1 : 0;
    const int oz = (t == 3 || t == 5 || t == 6) ? -1 : (t == 4 || t == 7 || t == 8) ? 1 : 0;
    const int yy = qc.iy + oy, zz = qc.iz + oz;
    if ((unsigned)yy >= (unsigned)g.dy || (unsigned)zz >= (unsigned)g.dz) continue;
    if (t > 0) {
      float gy = oy ? slab_gap(qc.fy, yy, yy) : 0.0f;
      float gz = oz ? slab_gap(qc.fz, zz, zz) : 0.0f;
      if ((gy * gy + gz * gz) * c2 > b.d2) continue;
    }
    const int x0 = max(qc.ix - g.xs, 0), x1 = min(qc.ix + g.xs, g.dx - 1);
    if (x0 > x1) continue;
    const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
    const uint32_t s = __ldg(row + x0), e = __ldg(row + x1 + 1);
    scan_run(g.pts, s, e, qx, qy, qz, b);
  }
  // Everything inside the 3x3x3 block was scanned or safely culled; any other point is
  // at least `gmin` cells away (faces of the block that lie outside the grid bound nothing).
  float gmin = 1.0e30f;
  if (qc.ix - g.xs > 0) gmin = fminf(gmin, (qc.fx - (float)(qc.ix - g.xs)) * g.inv_xs);
  if (qc.ix + g.xs + 1 < g.dx) gmin = fminf(gmin, ((float)(qc.ix + g.xs + 1) - qc.fx) * g.inv_xs);
  if (qc.iy - 1 > 0) gmin = fminf(gmin, qc.fy - (float)(qc.iy - 1));
  if (qc.iy + 2 < g.dy) gmin = fminf(gmin, (float)(qc.iy + 2) - qc.fy);
  if (qc.iz - 1 > 0) gmin = fminf(gmin, qc.fz - (float)(qc.iz - 1));
  if (qc.iz + 2 < g.dz) gmin = fminf(gmin, (float)(qc.iz + 2) - qc.fz);
  gmin = fmaxf(gmin - kCellSlack, 0.0f);
  return b.d2 <= gmin * gmin * c2;
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Phase 2: all 32 lanes cooperate on ONE query (qx,qy,qz and b uniform across the warp
// on entry and on exit).
__device__ __noinline__ void nn_phase2_warp(const GridDev& g, float qx, float qy, float qz,
                                            Best& b) {
  const int lane = threadIdx.x & 31;
  const QueryCell qc = query_cell(g, qx, qy, qz);
  const float c2 = g.c * g.c * 0.9999f;
  const float inv_c2 = 1.0f / c2;
  const int xsh = kCoarseShift + g.xs_shift, kCoarseX = 1 << xsh;  // super-cell edge in x-subcells
  const int cx = qc.ix >> xsh, cy = qc.iy >> kCoarseShift, cz = qc.iz >> kCoarseShift;
  Best lb = b;        // lane-local best
  float bd = b.d2;    // warp-uniform pruning bound (min over lanes so far)
  // first ring that can touch the grid
  int R = 0;
  R = max(R, max(-cx, cx - (g.cdx - 1)));
  R = max(R, max(-cy, cy - (g.cdy - 1)));
  R = max(R, max(-cz, cz - (g.cdz - 1)));
  for (;; ++R) {
    const int x0 = max(cx - R, 0), x1 = min(cx + R, g.cdx - 1);
    const int y0 = max(cy - R, 0), y1 = min(cy + R, g.cdy - 1);
    const int z0 = max(cz - R, 0), z1 = min(cz + R, g.cdz - 1);
    const int nx = x1 - x0 + 1, ny = y1 - y0 + 1, nz = z1 - z0 + 1;
    const int total = (nx > 0 && ny > 0 && nz > 0) ? nx * ny * nz : 0;
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      bool work = false;
      int ccx = 0, ccy = 0, ccz = 0;
      if (t < total) {
        ccx = x0 + t % nx;
        ccy = y0 + (t / nx) % ny;
        ccz = z0 + t / (nx * ny);
        const int ring = max(max(abs(ccx - cx), abs(ccy - cy)), abs(ccz - cz));
        if (ring == R) {
          float gx = slab_gap(qc.fx, ccx * kCoarseX, ccx * kCoarseX + kCoarseX - 1) * g.inv_xs;
          float gy = slab_gap(qc.fy, ccy * kCoarse, ccy * kCoarse + kCoarse - 1);
          float gz = slab_gap(qc.fz, ccz * kCoarse, ccz * kCoarse + kCoarse - 1);
          if ((gx * gx + gy * gy + gz * gz) * c2 <= bd)
            work = __ldg(&g.coarse_cnt[(ccz * g.cdy + ccy) * g.cdx + ccx]) != 0u;
        }
      }
      unsigned m = __ballot_sync(0xffffffffu, work);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int bx = __shfl_sync(0xffffffffu, ccx, src) * kCoarseX;
        const int by = __shfl_sync(0xffffffffu, ccy, src) * kCoarse;
        const int bz = __shfl_sync(0xffffffffu, ccz, src) * kCoarse;
        const float bdc = bd * inv_c2;  // bound in cells^2 (may be +inf)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = lane + rr * 32;
          const int yy = by + (r & 7), zz = bz + (r >> 3);
          if (yy < g.dy && zz < g.dz) {
            const float gy = slab_gap(qc.fy, yy, yy), gz = slab_gap(qc.fz, zz, zz);
            const float rem = bdc - (gy * gy + gz * gz);
            if (rem >= 0.0f) {
              const float w = (sqrtf(rem) + 2.0f * kCellSlack) * (float)g.xs;  // x-subcells
              int xa = bx, xb = min(bx + kCoarseX - 1, g.dx - 1);
              // clip the row to the x-extent of the search ball (w may be +inf)
              const float fl = qc.fx - w, fh = qc.fx + w;
              if (fl > (float)xa) xa = (int)floorf(fl);
              if (fh < (float)xb) xb = (int)floorf(fh);
              if (xa <= xb) {
                const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
                const uint32_t s = __ldg(row + xa), e = __ldg(row + xb + 1);
                scan_run(g.pts, s, e, qx, qy, qz, lb);
              }
            }
          }
        }
        bd = fminf(bd, warp_min_f(lb.d2));
      }
    }
    // termination: distance (cells) from the query to the faces of the scanned cube of
    // super-cells that still have grid behind them
    float gmin = 1.0e30f;
    bool open = false;
    if ((cx - R) > 0) { gmin = fminf(gmin, (qc.fx - (float)((cx - R) * kCoarseX)) * g.inv_xs); open = true; }
    if ((cx + R + 1) < g.cdx) { gmin = fminf(gmin, ((float)((cx + R + 1) * kCoarseX) - qc.fx) * g.inv_xs); open = true; }
    if ((cy - R) > 0) { gmin = fminf(gmin, qc.fy - (float)((cy - R) * kCoarse)); open = true; }
    if ((cy + R + 1) < g.cdy) { gmin = fminf(gmin, (float)((cy + R + 1) * kCoarse) - qc.fy); open = true; }
    if ((cz - R) > 0) { gmin = fminf(gmin, qc.fz - (float)((cz - R) * kCoarse)); open = true; }
    if ((cz + R + 1) < g.cdz) { gmin = fminf(gmin, (float)((cz + R + 1) * kCoarse) - qc.fz); open = true; }
    if (!open) break;
    gmin = fmaxf(gmin - kCellSlack, 0.0f);
    if (bd <= gmin * gmin * c2) break;
  }
  // lexicographic (d2, original index) min over the lanes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float od = __shfl_xor_sync(0xffffffffu, lb.d2, o);
    int oj = __shfl_xor_sync(0xffffffffu, lb.j, o);
    int oo = __shfl_xor_sync(0xffffffffu, lb.oi, o);
    if (od < lb.d2 || (od == lb.d2 && oo < lb.oi)) {
      lb.d2 = od;
      lb.j = oj;
      lb.oi = oo;
    }
  }
  b = lb;
}

// Optional search statistics (LC3D_STATS=1): [0] queries searched, [1] seeded by the previous
// match, [2] seeded/proven by the 3x3x3 probe, [3] seeded by a warp neighbour, [4] ball-walked,
// [5] warp-cooperative fallbacks, [6] candidate loop iterations (warp-level), [7] rows (warp).
struct SearchStats {
  unsigned long long c[12];  // [8] sum of per-warp cycles, [9] max per-warp cycles, [10] max batch cycles
};
__device__ __forceinline__ void stat_add(SearchStats* st, int k, unsigned long long v) {
  if (st && v) atomicAdd(&st->c[k], v);
}

// ---- ball walk -------------------------------------------------------------------------
// Given a valid candidate (b.d2 = squared distance to some target point, or the gate), the
// exact nearest neighbour lies in the ball of that radius around the query.  Each lane walks
// the cell rows that intersect ITS ball — rows culled by their y/z slab distance against the
// shrinking bound, the x-run of each row clipped to the ball — so the result is exact by
// construction.  The row loops run over the warp-wide window (lockstep), lanes mask
// themselves out of rows outside their own ball.
constexpr int kBallMaxW = 12;  // window half-width (cells) beyond which phase 2 takes over

__device__ __forceinline__ int ball_halfwidth(const GridDev& g, float d2) {
  const float w = sqrtf(d2) * g.inv_c * 1.0001f + 2.0f * kCellSlack;
  return w < 1.0e6f ? (int)ceilf(w) : 0x3fffffff;
}

// Warp-cooperative ball walk for ONE query (inputs and result uniform across the warp): the
// (2W+1)^2 cell rows that can intersect the ball of the incoming candidate are dealt out to the
// lanes 32 at a time, each lane scans the clipped run of its row, and the pruning bound is
// refreshed (warp min) between steps.  For a query a few centimetres off the target this is a
// handful of steps, where the super-cell ring search pays for whole 8^3 blocks.  Returns false
// (nothing done) when there is no finite candidate or the window is too wide.
constexpr int kCoopMaxW = 16;
__device__ __forceinline__ bool nn_ball_warp(const GridDev& g, float qx, float qy, float qz, Best& b) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int W = ball_halfwidth(g, b.d2);
  if (b.j < 0 || W > kCoopMaxW) return false;
  const QueryCell qc = query_cell(g, qx, qy, qz);
  const float inv_c2 = 1.0f / (g.c * g.c * 0.9999f);
  const int side = 2 * W + 1, total = side * side;
  const float inv_side = 1.0f / (float)side;
  Best lb = b;
  float bd = b.d2;
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    if (t < total) {
      const int rz = (int)(((float)t + 0.5f) * inv_side);  // exact: t < 33^2
      const int dy = t - rz * side - W, dz = rz - W;
      const int yy = qc.iy + dy, zz = qc.iz + dz;
      if ((unsigned)yy < (unsigned)g.dy && (unsigned)zz < (unsigned)g.dz) {
        const float gy = slab_gap(qc.fy, yy, yy), gz = slab_gap(qc.fz, zz, zz);
        const float rem = bd * inv_c2 - (gy * gy + gz * gz);
        if (rem >= 0.0f) {
          const float wx = (sqrtf(rem) + 2.0f * kCellSlack) * (float)g.xs;  // x-subcells
          const int xa = max((int)floorf(qc.fx - wx), 0), xb = min((int)floorf(qc.fx + wx), g.dx - 1);
          if (xa <= xb) {
            const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
            scan_run(g.pts, __ldg(row + xa), __ldg(row + xb + 1), qx, qy, qz, lb);
          }
        }
      }
    }
    bd = fminf(bd, warp_min_f(lb.d2));
  }
  // lexicographic (d2, original index) min over the lanes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(full, lb.d2, o);
    const int oj = __shfl_xor_sync(full, lb.j, o);
    const int oo = __shfl_xor_sync(full, lb.oi, o);
    if (od < lb.d2 || (od == lb.d2 && oo < lb.oi)) {
      lb.d2 = od;
      lb.j = oj;
      lb.oi = oo;
    }
  }
  b = lb;
  return true;
}

// Walks the cell rows that intersect the lane's ball: rows culled by their y/z slab distance
// against the shrinking bound, the x-run of each row clipped to the ball.  The row loops run
// over the warp-wide window (lockstep); lanes mask themselves out of rows outside their ball.
__device__ __forceinline__ void ball_walk(const GridDev& g, bool act, const QueryCell& qc, float qx,
                                          float qy, float qz, Best& b, unsigned* n_cand,
                                          unsigned* n_rows) {
  const unsigned full = 0xffffffffu;
  const float inv_c2 = 1.0f / (g.c * g.c * 0.9999f);
  int W = act ? ball_halfwidth(g, b.d2) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) W = max(W, __shfl_xor_sync(full, W, o));
  for (int dz = -W; dz <= W; ++dz) {
    const int zz = qc.iz + dz;
    const float gz = slab_gap(qc.fz, zz, zz);
    const float gz2 = gz * gz;
    const bool zok = act && (unsigned)zz < (unsigned)g.dz;
    for (int dy = -W; dy <= W; ++dy) {
      const int yy = qc.iy + dy;
      uint32_t s = 0, e = 0;
      if (zok && (unsigned)yy < (unsigned)g.dy) {
        const float gy = slab_gap(qc.fy, yy, yy);
        const float rem = b.d2 * inv_c2 - (gy * gy + gz2);
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
      if (n_rows) {
        *n_rows += 1;
        *n_cand += e - s;
      }
      for (uint32_t j = s; j < e; ++j) consider(__ldg(&g.pts[j]), (int)j, qx, qy, qz, b);
    }
  }
}

// Centre-out variant of the ball walk: the rows are visited ring by ring (Chebyshev rings of
// (y,z) cells around the query's own cell).  The centre rows shrink the bound first, so that the
// outer rings of a wide ball (stale seeds right after a large pose update) are culled by the slab
// test instead of being scanned with the initial radius, and the loop stops as soon as no lane's
// ball reaches the next ring (warp vote).  Pays off while the balls are several cells wide
// (iterations 1-2 of a 5-degree pair: 152 -> 132 us and 120 -> 91 us); in the converged regime
// (one-cell balls) the ring bookkeeping costs ~5 us per iteration, so the caller picks the walk
// per iteration (profiles/r02_summary.md).
__device__ __forceinline__ void ball_walk_rings(const GridDev& g, bool act, const QueryCell& qc, float qx,
                                                float qy, float qz, Best& b, unsigned* n_cand,
                                                unsigned* n_rows) {
  const unsigned full = 0xffffffffu;
  const float inv_c2 = 1.0f / (g.c * g.c * 0.9999f);
  int W = act ? ball_halfwidth(g, b.d2) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) W = max(W, __shfl_xor_sync(full, W, o));
  // distance (cells) from the query to the nearest face of its own (y,z) cell: every row of
  // ring r is at least r - 1 + m cells away
  const float m = fminf(fminf(qc.fy - (float)qc.iy, (float)(qc.iy + 1) - qc.fy),
                        fminf(qc.fz - (float)qc.iz, (float)(qc.iz + 1) - qc.fz));
  for (int r = 0; r <= W; ++r) {
    if (r >= 1) {
      const float gr = fmaxf((float)(r - 1) + m - kCellSlack, 0.0f);
      if (!__any_sync(full, act && gr * gr <= b.d2 * inv_c2)) break;
    }
    const int nside = r == 0 ? 1 : 4, len = r == 0 ? 1 : 2 * r;
    for (int side = 0; side < nside; ++side) {
      for (int k = 0; k < len; ++k) {
        // four sides of 2r cells each, walked around the ring
        int dy, dz;
        if (side == 0) { dy = k - r; dz = -r; }
        else if (side == 1) { dy = r; dz = k - r; }
        else if (side == 2) { dy = r - k; dz = r; }
        else { dy = -r; dz = r - k; }
        if (r == 0) dy = dz = 0;
        const int yy = qc.iy + dy, zz = qc.iz + dz;
        uint32_t s = 0, e = 0;
        if (act && (unsigned)zz < (unsigned)g.dz && (unsigned)yy < (unsigned)g.dy) {
          const float gy = slab_gap(qc.fy, yy, yy), gz = slab_gap(qc.fz, zz, zz);
          const float rem = b.d2 * inv_c2 - (gy * gy + gz * gz);
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
        if (n_rows) {
          *n_rows += 1;
          *n_cand += e - s;
        }
        for (uint32_t j = s; j < e; ++j) consider(__ldg(&g.pts[j]), (int)j, qx, qy, qz, b);
      }
    }
  }
}

// Exact gated 1-NN for one query per thread, seeded.  All 32 lanes must call.
// seed_j: position (sorted target order) of a plausible neighbour, e.g. the previous
// iteration's match, or -1.  gate may be +inf (unbounded): lanes whose ball is too large for
// the walk (or that found no candidate at all) go through the warp-cooperative ring search.
// DEFER: lanes that the ball walk cannot resolve cheaply (window wider than kDeferW, or no
// candidate at all) are NOT searched here: *deferred is set and the returned Best holds the
// best candidate so far (a bound for whoever finishes the job).  A far query costs tens of
// thousands of instructions; left to its own thread it serialises its whole warp and a
// handful of such warps becomes the tail of the kernel — the caller queues these queries
// for a warp-per-query pass spread over the whole GPU instead.
#ifndef LC3D_DEFER_W
#define LC3D_DEFER_W 2
#endif
#ifndef LC3D_COOP_W
#define LC3D_COOP_W 2
#endif
#ifndef LC3D_COOP_LANES
#define LC3D_COOP_LANES 0
#endif
constexpr int kCoopW = LC3D_COOP_W;          // balls wider than this many cells count as wide
// at most this many wide lanes per warp go cooperative.  0 = off, the measured optimum on the bench
// pair (same-box A/B, profiles/r02_summary.md: 0 -> 0.73 ms loop, 3 -> 0.75, 6 -> 0.80, 12 -> 0.90)
constexpr int kCoopLanes = LC3D_COOP_LANES;
constexpr int kDeferW = LC3D_DEFER_W;
template <bool DEFER = false, bool RINGS = false>
__device__ __forceinline__ Best nn_search_seeded(const GridDev& g, bool active, float qx, float qy,
                                                 float qz, float gate, int seed_j, SearchStats* stats,
                                                 bool* deferred = nullptr) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  Best b;
  b.d2 = gate;
  b.j = -1;
  b.oi = 0x7fffffff;
  bool need = active && g.n > 0;
  const QueryCell qc = query_cell(g, qx, qy, qz);
  // 1. seed from the previous match
  bool seeded = false;
  if (need && seed_j >= 0) {
    consider(__ldg(&g.pts[seed_j]), seed_j, qx, qy, qz, b);
    seeded = b.j >= 0;
  }
  const unsigned n_prev = stats ? __popc(__ballot_sync(full, seeded)) : 0;
  // 2. unseeded lanes probe their 3x3x3 block (this may already prove the result)
  bool probed = false;
  if (need && !seeded) {
    if (nn_phase1(g, qc, qx, qy, qz, b)) need = false;
    probed = b.j >= 0;
  }
  // 3. lanes that still have no candidate borrow a warp neighbour's match as a seed
  {
    const unsigned have = __ballot_sync(full, b.j >= 0);
    const bool want = need && b.j < 0;
    if (have && __any_sync(full, want)) {
      // nearest lane (in lane index = Morton order) that has a candidate
      const unsigned lower = have & ((1u << lane) - 1u), upper = have & ~((2u << lane) - 1u);
      int src = lane;
      if (lower && upper) {
        const int lo = 31 - __clz(lower), hi = __ffs(upper) - 1;
        src = (lane - lo <= hi - lane) ? lo : hi;
      } else if (lower) {
        src = 31 - __clz(lower);
      } else if (upper) {
        src = __ffs(upper) - 1;
      }
      const int sj = __shfl_sync(full, b.j, src);
      if (want && src != lane && sj >= 0) consider(__ldg(&g.pts[sj]), sj, qx, qy, qz, b);
    }
  }
  const bool borrowed = need && !seeded && !probed && b.j >= 0;
  // 4. ball walk for every lane whose ball is small enough, warp-cooperative rings otherwise
  const int Wl = need ? ball_halfwidth(g, b.d2) : 0;
  // A FEW lanes with wide balls would drag the whole warp through their windows in lockstep
  // (thousands of instructions at 1/32 utilisation): when there are at most kCoopLanes of them
  // they are left out of the per-lane walk and handled one at a time by all 32 lanes below.
  // Many wide lanes (incoherent early iterations, non-overlap regions) stay in the per-lane
  // walk, where they keep each other company.
  int wmax = DEFER ? kDeferW : kBallMaxW;
  if (!DEFER && kCoopLanes > 0) {
    const unsigned wide = __ballot_sync(full, need && Wl > kCoopW && b.j >= 0);
    if (wide && __popc(wide) <= kCoopLanes) wmax = kCoopW;
  }
  const bool walk = need && (Wl <= wmax || (b.j < 0 && Wl <= kBallMaxW));
  unsigned n_cand = 0, n_rows = 0;
  if (__any_sync(full, walk)) {
    if (RINGS)
      ball_walk_rings(g, walk, qc, qx, qy, qz, b, stats ? &n_cand : nullptr, stats ? &n_rows : nullptr);
    else
      ball_walk(g, walk, qc, qx, qy, qz, b, stats ? &n_cand : nullptr, stats ? &n_rows : nullptr);
  }
  if (walk) need = false;
  if (DEFER) *deferred = need;
  unsigned todo = DEFER ? 0u : __ballot_sync(full, need);
  if (stats) {
    unsigned mc = n_cand, mr = n_rows;
    for (int o = 16; o > 0; o >>= 1) {
      mc = max(mc, __shfl_xor_sync(full, mc, o));
      mr = max(mr, __shfl_xor_sync(full, mr, o));
    }
    const unsigned na = __popc(__ballot_sync(full, active && g.n > 0));
    const unsigned npb = __popc(__ballot_sync(full, probed)), nbr = __popc(__ballot_sync(full, borrowed));
    const unsigned nw = __popc(__ballot_sync(full, walk));
    const unsigned n_left = __popc(__ballot_sync(full, need));  // ring searches / deferred
    if (lane == 0) {
      stat_add(stats, 0, na);
      stat_add(stats, 1, n_prev);
      stat_add(stats, 2, npb);
      stat_add(stats, 3, nbr);
      stat_add(stats, 4, nw);
      stat_add(stats, 5, n_left);
      stat_add(stats, 6, mc);
      stat_add(stats, 7, mr);
    }
  }
  int last_j = -1;  // most recent ring-search result: a good seed for the next lane (Morton
                    // neighbours), it bounds that lane's search ball from the start
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    if (lane == src && last_j >= 0) consider(__ldg(&g.pts[last_j]), last_j, qx, qy, qz, b);
    Best wb;
    wb.d2 = __shfl_sync(full, b.d2, src);
    wb.j = __shfl_sync(full, b.j, src);
    wb.oi = __shfl_sync(full, b.oi, src);
    const float wx = __shfl_sync(full, qx, src);
    const float wy = __shfl_sync(full, qy, src);
    const float wz = __shfl_sync(full, qz, src);
    if (!nn_ball_warp(g, wx, wy, wz, wb)) nn_phase2_warp(g, wx, wy, wz, wb);
    if (lane == src) b = wb;
    if (wb.j >= 0) last_j = wb.j;
  }
  return b;
}

// Full exact 1-NN for one query per thread.  MUST be called by all 32 lanes of every
// warp (inactive lanes pass active=false).  gate: accept only d2 <= gate (+inf = none).
__device__ __forceinline__ Best nn_search(const GridDev& g, bool active, float qx, float qy,
                                          float qz, float gate) {
  Best b;
  b.d2 = gate;
  b.j = -1;
  b.oi = 0x7fffffff;
  bool need2 = false;
  if (active && g.n > 0) {
    const QueryCell qc = query_cell(g, qx, qy, qz);
    need2 = !nn_phase1(g, qc, qx, qy, qz, b);
  }
  unsigned todo = __ballot_sync(0xffffffffu, need2);
  const int lane = threadIdx.x & 31;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    Best wb;
    wb.d2 = __shfl_sync(0xffffffffu, b.d2, src);
    wb.j = __shfl_sync(0xffffffffu, b.j, src);
    wb.oi = __shfl_sync(0xffffffffu, b.oi, src);
    const float wx = __shfl_sync(0xffffffffu, qx, src);
    const float wy = __shfl_sync(0xffffffffu, qy, src);
    const float wz = __shfl_sync(0xffffffffu, qz, src);
    nn_phase2_warp(g, wx, wy, wz, wb);
    if (lane == src) b = wb;
  }
  return b;
}

}  // namespace lc3d
