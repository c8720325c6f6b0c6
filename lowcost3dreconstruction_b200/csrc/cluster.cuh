// cluster.cuh — Euclidean cluster extraction (pcl_tools/cluster_extraction.cpp:88-101;
// pcl::EuclideanClusterExtraction -> extractEuclideanClusters, PCL 1.8.1
// segmentation/impl/extract_clusters.hpp).
//
// PCL grows clusters breadth-first over radiusSearch(point, tolerance) neighbourhoods: the
// clusters are exactly the connected components of the graph with an edge wherever
// d2 < (float)(tolerance^2) (FLANN's RadiusResultSet is strict), d2 the float32 squared distance
// ((dx^2)+dy^2)+dz^2.  Here: lock-free union-find over the points in grid order.  Every point
// walks the cell rows that meet its tolerance ball (x-runs clipped to the ball), looks only at
// EARLIER points (each edge once), skips candidates already known to share its root (one 4-byte
// load instead of a distance) and hooks the larger root under the smaller one with atomicCAS.
// The partition does not depend on the scheduling; clusters are then named by their smallest
// original point index, which makes the output deterministic.
#pragma once
#include "search.cuh"

namespace lc3d {

__device__ __forceinline__ int cc_find(int* parent, int a) {
  volatile int* vp = parent;
  int cur = a, next;
  while ((next = vp[cur]) != cur) {
    const int nn = vp[next];
    if (nn != next) vp[cur] = nn;  // path halving (roots only ever move to smaller ids)
    cur = nn;
  }
  return cur;
}

// Returns the common root after the union.
__device__ __forceinline__ int cc_union(int* parent, int ra, int b) {
  int rb = cc_find(parent, b);
  while (ra != rb) {
    if (ra < rb) {
      const int t = ra;
      ra = rb;
      rb = t;
    }
    const int old = atomicCAS(&parent[ra], ra, rb);  // hook the larger root under the smaller
    if (old == ra) return rb;
    ra = cc_find(parent, old);
    rb = cc_find(parent, rb);
  }
  return ra;
}

__global__ void __launch_bounds__(256) cc_init_kernel(int* __restrict__ parent, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) parent[j] = j;
}

// One thread per indexed point (grid order).  r2 = (float)(tolerance^2); r = a float >= tolerance.
__global__ void __launch_bounds__(128)
    cc_hook_kernel(const __grid_constant__ GridDev g, float r, float r2, int* __restrict__ parent) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n) return;
  const float4 p = __ldg(&g.pts[j]);
  const QueryCell qc = query_cell(g, p.x, p.y, p.z);
  const float rc = r * g.inv_c * 1.0001f + 2.0f * kCellSlack;  // tolerance in cells
  const int W = (int)ceilf(rc);
  const float rc2 = rc * rc;
  int root = cc_find(parent, j);
  for (int dz = -W; dz <= W; ++dz) {
    const int zz = qc.iz + dz;
    if ((unsigned)zz >= (unsigned)g.dz) continue;
    const float gz = slab_gap(qc.fz, zz, zz);
    for (int dy = -W; dy <= W; ++dy) {
      const int yy = qc.iy + dy;
      if ((unsigned)yy >= (unsigned)g.dy) continue;
      const float gy = slab_gap(qc.fy, yy, yy);
      const float rem = rc2 - (gy * gy + gz * gz);
      if (rem < 0.0f) continue;
      const float wx = (sqrtf(rem) + 2.0f * kCellSlack) * (float)g.xs;  // x-subcells
      const int xa = max((int)floorf(qc.fx - wx), 0), xb = min((int)floorf(qc.fx + wx), g.dx - 1);
      if (xa > xb) continue;
      const uint32_t* row = g.cell_start + (size_t)(zz * g.dy + yy) * g.dx;
      const uint32_t s = __ldg(row + xa);
      const uint32_t e = min(__ldg(row + xb + 1), (uint32_t)j);  // earlier points only
      for (uint32_t k = s; k < e; ++k) {
        if (((volatile int*)parent)[k] == root) continue;  // already together
        const float4 t = __ldg(&g.pts[k]);
        if (dist2_exact(p.x, p.y, p.z, t.x, t.y, t.z) < r2) root = cc_union(parent, root, (int)k);
      }
    }
  }
}

// root_of[j] = representative; size / smallest original index per representative.  The result
// goes to its own array: writing it into parent[] would race with the path-halving stores of
// concurrent finds, which may overwrite it with a non-root ancestor.
__global__ void __launch_bounds__(256)
    cc_flatten_kernel(const GridDev g, int* __restrict__ parent, int* __restrict__ root_of,
                      uint32_t* __restrict__ size, uint32_t* __restrict__ min_orig) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n) return;
  const int r = cc_find(parent, j);
  root_of[j] = r;
  atomicAdd(&size[r], 1u);
  atomicMin(&min_orig[r], (uint32_t)__float_as_int(g.pts[j].w));
}

// flags[j] = 1 for representatives whose component size lies in [min_size, max_size].
__global__ void __launch_bounds__(256)
    cc_select_kernel(const int* __restrict__ root_of, const uint32_t* __restrict__ size, int n,
                     uint32_t min_size, uint32_t max_size, uint32_t* __restrict__ flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  flags[j] = (root_of[j] == j && size[j] >= min_size && size[j] <= max_size) ? 1u : 0u;
}

// Compact (size, min original index, representative) of the selected components.
__global__ void __launch_bounds__(256)
    cc_gather_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos,
                     const uint32_t* __restrict__ size, const uint32_t* __restrict__ min_orig, int n,
                     uint32_t* __restrict__ out3) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !flags[j]) return;
  const uint32_t o = pos[j];
  out3[3 * (size_t)o + 0] = size[j];
  out3[3 * (size_t)o + 1] = min_orig[j];
  out3[3 * (size_t)o + 2] = (uint32_t)j;
}

// labels[original index] = rank of the point's cluster (rank_of[compact position]) or -1.
__global__ void __launch_bounds__(256)
    cc_label_kernel(const GridDev g, const int* __restrict__ root_of, const uint32_t* __restrict__ flags,
                    const uint32_t* __restrict__ pos, const int32_t* __restrict__ rank_of,
                    int32_t* __restrict__ labels) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g.n) return;
  const int r = root_of[j];
  labels[__float_as_int(g.pts[j].w)] = flags[r] ? rank_of[pos[r]] : -1;
}

}  // namespace lc3d
