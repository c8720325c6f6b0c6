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

__device__ __forceinline__ void consider(const float4 p, int j, float qx, float qy, float qz,
                                         Best& b) {
  float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
  int oi = __float_as_int(p.w);
  if (d2 < b.d2 || (d2 == b.d2 && oi < b.oi)) {
    b.d2 = d2;
    b.j = j;
    b.oi = oi;
  }
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
  c.fx = cell_coord(qx, g.ox, g.inv_c);
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
    const int x0 = max(qc.ix - 1, 0), x1 = min(qc.ix + 1, g.dx - 1);
    if (x0 > x1) continue;
    const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
    const uint32_t s = __ldg(row + x0), e = __ldg(row + x1 + 1);
    scan_run(g.pts, s, e, qx, qy, qz, b);
  }
  // Everything inside the 3x3x3 block was scanned or safely culled; any other point is
  // at least `gmin` cells away (faces of the block that lie outside the grid bound nothing).
  float gmin = 1.0e30f;
  if (qc.ix - 1 > 0) gmin = fminf(gmin, qc.fx - (float)(qc.ix - 1));
  if (qc.ix + 2 < g.dx) gmin = fminf(gmin, (float)(qc.ix + 2) - qc.fx);
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
  const int cx = qc.ix >> kCoarseShift, cy = qc.iy >> kCoarseShift, cz = qc.iz >> kCoarseShift;
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
          float gx = slab_gap(qc.fx, ccx * kCoarse, ccx * kCoarse + kCoarse - 1);
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
        const int bx = __shfl_sync(0xffffffffu, ccx, src) * kCoarse;
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
              const float w = sqrtf(rem) + 2.0f * kCellSlack;
              int xa = bx, xb = min(bx + kCoarse - 1, g.dx - 1);
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
    if ((cx - R) > 0) { gmin = fminf(gmin, qc.fx - (float)((cx - R) * kCoarse)); open = true; }
    if ((cx + R + 1) < g.cdx) { gmin = fminf(gmin, (float)((cx + R + 1) * kCoarse) - qc.fx); open = true; }
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
