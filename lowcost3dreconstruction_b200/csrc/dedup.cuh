// dedup.cuh — the de-duplication step of accumulate_clouds (pcl_tools/accumulate_clouds.cpp:
// 100-111): the reference filters the whole source cloud with one
// pcl::CropBox per TARGET point (O(N*M)); here every source point asks the grid index of the
// target whether any target point's +/- radius box contains it (O(N), early exit on the first
// hit).  The box test is evaluated exactly as CropBox does: corners (float)((double)t -/+ r),
// inclusive bounds.
#pragma once
#include "search.cuh"

namespace lc3d {

__global__ void __launch_bounds__(256)
    box_dedup_kernel(const GridDev g, const float4* __restrict__ src, int n, double radius,
                     uint32_t* __restrict__ keep_flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 s = src[i];
  if (!finite3(s.x, s.y, s.z)) {  // CropBox drops non-finite points
    keep_flags[i] = 0u;
    return;
  }
  bool hit = false;
  if (g.n > 0) {
    // search window: every cell that can hold a target point whose box reaches s (the margin
    // covers the float rounding of the box corners; the exact test below decides)
    const float rm = (float)radius * 1.0001f + 1e-6f * g.c;
    const float wc = rm * g.inv_c + 2.0f * kCellSlack;
    const QueryCell qc = query_cell(g, s.x, s.y, s.z);
    const float wcx = wc * (float)g.xs;  // x-subcells
    const int x0 = max((int)floorf(qc.fx - wcx), 0), x1 = min((int)floorf(qc.fx + wcx), g.dx - 1);
    const int y0 = max((int)floorf(qc.fy - wc), 0), y1 = min((int)floorf(qc.fy + wc), g.dy - 1);
    const int z0 = max((int)floorf(qc.fz - wc), 0), z1 = min((int)floorf(qc.fz + wc), g.dz - 1);
    if (x0 <= x1 && y0 <= y1 && z0 <= z1) {
      auto scan_row = [&](int yy, int zz) {
        const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
        const uint32_t a = __ldg(row + x0), b = __ldg(row + x1 + 1);
        for (uint32_t j = a; j < b; ++j) {
          const float4 t = __ldg(&g.pts[j]);
          const float lx = __double2float_rn((double)t.x - radius), hx = __double2float_rn((double)t.x + radius);
          const float ly = __double2float_rn((double)t.y - radius), hy = __double2float_rn((double)t.y + radius);
          const float lz = __double2float_rn((double)t.z - radius), hz = __double2float_rn((double)t.z + radius);
          const bool outside = s.x < lx || s.y < ly || s.z < lz || s.x > hx || s.y > hy || s.z > hz;
          if (!outside) return true;
        }
        return false;
      };
      // the query's own row first (overlapping surfaces hit there), then the rest of the window
      const int cy = min(max(qc.iy, y0), y1), cz = min(max(qc.iz, z0), z1);
      hit = scan_row(cy, cz);
      for (int zz = z0; zz <= z1 && !hit; ++zz)
        for (int yy = y0; yy <= y1 && !hit; ++yy)
          if (yy != cy || zz != cz) hit = scan_row(yy, zz);
    }
  }
  keep_flags[i] = hit ? 0u : 1u;
}

}  // namespace lc3d
