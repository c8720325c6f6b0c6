// voxel.cuh — pcl::VoxelGrid::applyFilter (cloud_downsampling.cpp:73-76):
// voxel index keys -> stable radix sort -> run heads -> scan -> one thread per voxel sums
// its points sequentially in float32 (ascending input index), CentroidPoint semantics.
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

__global__ void __launch_bounds__(128)
    voxel_centroid_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                          const uint32_t* __restrict__ flags, const uint32_t* __restrict__ rank, int n,
                          VoxelIO io) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !flags[j]) return;
  const uint32_t key = keys[j];
  const uint32_t m = rank[j];
  float sx = 0, sy = 0, sz = 0, snx = 0, sny = 0, snz = 0, sc = 0, sr = 0, sg = 0, sb = 0, sa = 0;
  int cnt = 0;
  for (int t = j; t < n && keys[t] == key; ++t) {
    const uint32_t pi = vals[t];
    const float4 p = io.xyz[pi];
    sx += p.x;
    sy += p.y;
    sz += p.z;
    if (io.nrm) {
      const float4 q = io.nrm[pi];
      snx += q.x;
      sny += q.y;
      snz += q.z;
    }
    if (io.curv) sc += io.curv[pi];
    if (io.rgba) {
      const uint32_t c = io.rgba[pi];
      sr += (float)((c >> 16) & 0xff);
      sg += (float)((c >> 8) & 0xff);
      sb += (float)(c & 0xff);
      sa += (float)((c >> 24) & 0xff);
    }
    if (io.voxel_of_point) io.voxel_of_point[pi] = (int32_t)m;
    ++cnt;
  }
  const float fn = (float)cnt;
  io.out_xyz[3 * (size_t)m + 0] = sx / fn;
  io.out_xyz[3 * (size_t)m + 1] = sy / fn;
  io.out_xyz[3 * (size_t)m + 2] = sz / fn;
  if (io.out_nrm && io.nrm) {  // CentroidPoint: summed normal, normalised
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
  if (io.out_curv && io.curv) io.out_curv[m] = sc / fn;
  if (io.out_rgba && io.rgba)
    io.out_rgba[m] = ((uint32_t)(sa / fn) << 24) | ((uint32_t)(sr / fn) << 16) |
                     ((uint32_t)(sg / fn) << 8) | (uint32_t)(sb / fn);
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t* p, int n, int32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace lc3d
