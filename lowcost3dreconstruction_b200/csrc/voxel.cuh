// voxel.cuh — pcl::VoxelGrid::applyFilter (cloud_downsampling.cpp:73-76):
// voxel index keys -> stable radix sort -> run heads -> scan -> per-voxel sums in float32,
// point by point in ascending input index (CentroidPoint semantics); long runs are summed by a
// whole warp in the same order.
#pragma once
#include "grid.cuh"

namespace lc3d {

struct VoxelParams {
  float inv[3];
  int minb[3];
  int mul[3];
};

__global__ void __launch_bounds__(256)
    voxel_keys_kernel(const float4* __restrict__ xyz, int n, VoxelParams vp, uint32_t* __restrict__ keys,
                      uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = xyz[i];
  uint32_t key = 0xffffffffu;
  if (finite3(p.x, p.y, p.z)) {
    // ijk = (int)(floor(x * inv) - (float)min_b), float32
    const int i0 = (int)(floorf(p.x * vp.inv[0]) - (float)vp.minb[0]);
    const int i1 = (int)(floorf(p.y * vp.inv[1]) - (float)vp.minb[1]);
    const int i2 = (int)(floorf(p.z * vp.inv[2]) - (float)vp.minb[2]);
    key = (uint32_t)(i0 * vp.mul[0] + i1 * vp.mul[1] + i2 * vp.mul[2]);
  }
  keys[i] = key;
  vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
    voxel_heads_kernel(const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ flags) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const uint32_t k = keys[j];
  flags[j] = (k != 0xffffffffu && (j == 0 || keys[j - 1] != k)) ? 1u : 0u;
}

struct VoxelIO {
  const float4* xyz;
  const float4* nrm;      // or null
  const uint32_t* rgba;   // or null
  const float* curv;      // or null
  float* out_xyz;         // m x 3
  float* out_nrm;         // m x 3 or null
  uint32_t* out_rgba;     // or null
  float* out_curv;        // or null
  int32_t* voxel_of_point;  // n or null
};

// Sums of one voxel (CentroidPoint): float32, accumulated point by point in ascending input index.
struct VoxelAcc {
  float sx = 0, sy = 0, sz = 0, snx = 0, sny = 0, snz = 0, sc = 0, sr = 0, sg = 0, sb = 0, sa = 0;
  int cnt = 0;
};
__device__ __forceinline__ void voxel_write(const VoxelIO& io, uint32_t m, const VoxelAcc& a) {
  const float fn = (float)a.cnt;
  io.out_xyz[3 * (size_t)m + 0] = a.sx / fn;
  io.out_xyz[3 * (size_t)m + 1] = a.sy / fn;
  io.out_xyz[3 * (size_t)m + 2] = a.sz / fn;
  if (io.out_nrm && io.nrm) {  // CentroidPoint: summed normal, normalised
    float snx = a.snx, sny = a.sny, snz = a.snz;
    const float n2 = snx * snx + sny * sny + snz * snz;
    if (n2 > 0.0f) {
      const float nn = sqrtf(n2);
      snx /= nn;
      sny /= nn;
      snz /= nn;
    }
    io.out_nrm[3 * (size_t)m + 0] = snx;
    io.out_nrm[3 * (size_t)m + 1] = sny;
    io.out_nrm[3 * (size_t)m + 2] = snz;
  }
  if (io.out_curv && io.curv) io.out_curv[m] = a.sc / fn;
  if (io.out_rgba && io.rgba)
    io.out_rgba[m] = ((uint32_t)(a.sa / fn) << 24) | ((uint32_t)(a.sr / fn) << 16) |
                     ((uint32_t)(a.sg / fn) << 8) | (uint32_t)(a.sb / fn);
}

// One thread per sorted position; the thread at the head of a voxel's run sums it.  Runs of up
// to kVoxelSerial points (the usual case: a few points per voxel) are summed by that thread alone.
// Longer runs — a large leaf such as the CLI default of 1.0 packs the whole cloud into a handful
// of voxels — are handed to the whole warp: 32 points are loaded at a time (coalesced index and
// gather loads in flight together) and then added ONE BY ONE in run order, every lane keeping
// the same running sums, so the float32 result is bit-identical to the sequential sum while the
// memory latency of a 100k-point run is paid 32 points at a time instead of once per point.
constexpr int kVoxelSerial = 32;
__global__ void __launch_bounds__(128)
    voxel_centroid_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                          const uint32_t* __restrict__ flags, const uint32_t* __restrict__ rank, int n,
                          VoxelIO io) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool head = j < n && flags[j];
  uint32_t key = 0, m = 0;
  bool big = false;
  if (head) {
    key = keys[j];
    m = rank[j];
    VoxelAcc a;
    int t = j;
    for (; t < n && keys[t] == key && a.cnt < kVoxelSerial; ++t) {
      const uint32_t pi = vals[t];
      const float4 p = io.xyz[pi];
      a.sx += p.x;
      a.sy += p.y;
      a.sz += p.z;
      if (io.nrm) {
        const float4 q = io.nrm[pi];
        a.snx += q.x;
        a.sny += q.y;
        a.snz += q.z;
      }
      if (io.curv) a.sc += io.curv[pi];
      if (io.rgba) {
        const uint32_t c = io.rgba[pi];
        a.sr += (float)((c >> 16) & 0xff);
        a.sg += (float)((c >> 8) & 0xff);
        a.sb += (float)(c & 0xff);
        a.sa += (float)((c >> 24) & 0xff);
      }
      if (io.voxel_of_point) io.voxel_of_point[pi] = (int32_t)m;
      ++a.cnt;
    }
    big = t < n && keys[t] == key;  // the run goes on: redo it with the whole warp
    if (!big) voxel_write(io, m, a);
  }
  unsigned bm = __ballot_sync(full, big);
  while (bm) {
    const int src = __ffs(bm) - 1;
    bm &= bm - 1;
    const int j0 = __shfl_sync(full, j, src);
    const uint32_t k0 = __shfl_sync(full, key, src), m0 = __shfl_sync(full, m, src);
    VoxelAcc a;
    for (int t0 = j0;; t0 += 32) {
      const int t = t0 + lane;
      const bool valid = t < n && keys[t] == k0;
      float4 p = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
      float cv = 0.f;
      uint32_t c = 0;
      if (valid) {
        const uint32_t pi = vals[t];
        p = io.xyz[pi];
        if (io.nrm) q = io.nrm[pi];
        if (io.curv) cv = io.curv[pi];
        if (io.rgba) c = io.rgba[pi];
        if (io.voxel_of_point) io.voxel_of_point[pi] = (int32_t)m0;
      }
      const int nv = __popc(__ballot_sync(full, valid));  // the valid lanes are a prefix of the warp
      for (int k = 0; k < nv; ++k) {
        a.sx += __shfl_sync(full, p.x, k);
        a.sy += __shfl_sync(full, p.y, k);
        a.sz += __shfl_sync(full, p.z, k);
        if (io.nrm) {
          a.snx += __shfl_sync(full, q.x, k);
          a.sny += __shfl_sync(full, q.y, k);
          a.snz += __shfl_sync(full, q.z, k);
        }
        if (io.curv) a.sc += __shfl_sync(full, cv, k);
        if (io.rgba) {
          const uint32_t ck = __shfl_sync(full, c, k);
          a.sr += (float)((ck >> 16) & 0xff);
          a.sg += (float)((ck >> 8) & 0xff);
          a.sb += (float)(ck & 0xff);
          a.sa += (float)((ck >> 24) & 0xff);
        }
      }
      a.cnt += nv;
      if (nv < 32) break;
    }
    if (lane == 0) voxel_write(io, m0, a);
  }
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t* p, int n, int32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace lc3d
