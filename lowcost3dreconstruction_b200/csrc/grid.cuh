// grid.cuh — uniform-grid spatial index over a cloud (replaces pcl::search::KdTree /
// KdTreeFLANN: implicit in icp.setInputTarget, fine_registration.cpp:109; explicit at
// normal_estimation.cpp:89-90).
//
// Layout in HBM (all SoA, 16-byte vector loads):
//   pts[j]        float4 (x, y, z, bits(original index)), sorted by linear cell id
//                 (x fastest), stable => ascending original index inside a cell
//   nrm[j]        float4 (nx, ny, nz, curvature) in the same order (optional)
//   cell_start[c] uint32, ncell+1 entries: points of cell c are [cell_start[c], cell_start[c+1]).
//                 Because x is the fastest-varying cell coordinate, a run of cells
//                 along x is ONE contiguous point range: a 3x3x3 neighbourhood is 9
//                 (start,end) lookups, not 27.
//   coarse_cnt[C] uint32 occupancy of 8x8x8-cell super-cells: far-field rejection and
//                 ring expansion run on this small table.
// Build: bbox reduce -> density probe -> cell keys (+histogram) -> hand-written radix
// sort -> exclusive scan of the histogram -> gather.
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>

#include "common.cuh"
#include "scan_sort.cuh"

namespace lc3d {

constexpr int kCoarseShift = 3;  // super-cell = 8^3 cells
constexpr int kCoarse = 1 << kCoarseShift;
constexpr float kCellSlack = 1e-3f;  // cells; covers float rounding of cell coordinates
constexpr int kMaxDim = 2048;
constexpr int kMaxDimX = 4096;  // x-subcells
// Cap of the dense cell table.  A surface cloud occupies a vanishing fraction of its bounding
// volume, so the table (4 B per cell) grows with extent^3 while the points grow with extent^2: the
// cap is what keeps x-subdivided millimetre cells affordable.  Default 2^26 cells (256 MB of the
// 180 GB), raised for multi-million-point clouds up to 2^29 (2 GB) so that they keep the cell edge
// and x subdivision the search is tuned for instead of a coarser grid; LC3D_MAX_CELLS_LOG2 overrides.
inline int64_t max_cells_for(int64_t n) {
  if (const char* e = std::getenv("LC3D_MAX_CELLS_LOG2")) {
    const int b = std::atoi(e);
    if (b >= 16 && b <= 30) return (int64_t)1 << b;
  }
  int64_t cap = (int64_t)1 << 26;
  while (cap < ((int64_t)1 << 29) && cap < 48 * n) cap <<= 1;
  return cap;
}

struct GridDev {
  float ox, oy, oz;
  float c, inv_c;
  // cells are c x c x c, except that x is subdivided `xs` (power of two) times: rows (y,z) stay
  // few while the x-run of a row clips tightly to the search ball.  dx counts x-SUBcells,
  // inv_cx = xs / c, inv_xs = 1 / xs.
  float inv_cx, inv_xs;
  int xs, xs_shift;
  int dx, dy, dz;
  int cdx, cdy, cdz;
  int n;  // finite points indexed
  const uint32_t* cell_start;
  const uint32_t* coarse_cnt;
  const float4* pts;
  const float4* nrm;
  // dilated occupancy bits at c-cell resolution (x packed 32 cells per word), or null: a clear
  // bit proves that no indexed point lies within (occ_r - 0.01) * c of ANY position inside the cell
  const uint32_t* occ;
  int occ_wx, occ_r;
};

struct Grid {
  GridDev v{};
  DevBuf cell_start, coarse_cnt, pts, nrm, occ_raw, occ;
  int64_t ncell = 0;
  double pts_per_cell_est = 0.0;  // (cell edge / point spacing)^2 / x subdivision, from the density probe
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
  void release() {
    cell_start.release();
    coarse_cnt.release();
    pts.release();
    nrm.release();
    occ_raw.release();
    occ.release();
  }
};

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
inline float ord2f(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  float f;
  std::memcpy(&f, &b, 4);
  return f;
}

struct BBoxOut {
  uint32_t lo[3], hi[3];
  uint32_t count;
  uint32_t pad;
};

__global__ void bbox_init(BBoxOut* o) {
  if (threadIdx.x == 0) {
    o->lo[0] = o->lo[1] = o->lo[2] = 0xffffffffu;
    o->hi[0] = o->hi[1] = o->hi[2] = 0u;
    o->count = 0;
    o->pad = 0;  // receives the probe's occupied-cell count
  }
}

__global__ void __launch_bounds__(256) bbox_reduce(const float4* __restrict__ p, int n, BBoxOut* o) {
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  uint32_t cnt = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 q = p[i];
    if (finite3(q.x, q.y, q.z)) {
      lo[0] = fminf(lo[0], q.x); hi[0] = fmaxf(hi[0], q.x);
      lo[1] = fminf(lo[1], q.y); hi[1] = fmaxf(hi[1], q.y);
      lo[2] = fminf(lo[2], q.z); hi[2] = fmaxf(hi[2], q.z);
      ++cnt;
    }
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fminf(lo[d], __shfl_down_sync(0xffffffffu, lo[d], o2));
      hi[d] = fmaxf(hi[d], __shfl_down_sync(0xffffffffu, hi[d], o2));
    }
    cnt += __shfl_down_sync(0xffffffffu, cnt, o2);
  }
  // block combine in shared memory, then 7 atomics per BLOCK (not per warp)
  __shared__ float s_lo[8][3], s_hi[8][3];
  __shared__ uint32_t s_cnt[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      s_lo[w][d] = lo[d];
      s_hi[w][d] = hi[d];
    }
    s_cnt[w] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int ww = 1; ww < (int)(blockDim.x >> 5); ++ww) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        lo[d] = fminf(lo[d], s_lo[ww][d]);
        hi[d] = fmaxf(hi[d], s_hi[ww][d]);
      }
      cnt += s_cnt[ww];
    }
    if (cnt > 0) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        atomicMin(&o->lo[d], f2ord(lo[d]));
        atomicMax(&o->hi[d], f2ord(hi[d]));
      }
      atomicAdd(&o->count, cnt);
    }
  }
}

// Density probe: occupancy bitmask of a 64^3 grid over the bbox; the number of occupied
// probe cells estimates the sampled surface area and hence the point spacing.
constexpr int kProbe = 64;
__device__ __forceinline__ float ord2f_dev(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// The bbox is read from device memory (written by bbox_reduce just before on the same stream),
// so the build needs ONE host round trip (bbox + occupancy together) instead of two.
__global__ void __launch_bounds__(256)
    probe_mark(const float4* __restrict__ p, int n, const BBoxOut* __restrict__ bb,
               uint32_t* __restrict__ bits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || bb->count <= 1u) return;
  const float ox = ord2f_dev(bb->lo[0]), oy = ord2f_dev(bb->lo[1]), oz = ord2f_dev(bb->lo[2]);
  const double ex = (double)ord2f_dev(bb->hi[0]) - (double)ox, ey = (double)ord2f_dev(bb->hi[1]) - (double)oy,
               ez = (double)ord2f_dev(bb->hi[2]) - (double)oz;
  const double maxext = fmax(ex, fmax(ey, ez));
  if (!(maxext > 0.0)) return;
  const float sx = (float)((double)kProbe / fmax(ex, maxext * 1e-6));
  const float sy = (float)((double)kProbe / fmax(ey, maxext * 1e-6));
  const float sz = (float)((double)kProbe / fmax(ez, maxext * 1e-6));
  float4 q = p[i];
  if (!finite3(q.x, q.y, q.z)) return;
  int ix = min(max((int)((q.x - ox) * sx), 0), kProbe - 1);
  int iy = min(max((int)((q.y - oy) * sy), 0), kProbe - 1);
  int iz = min(max((int)((q.z - oz) * sz), 0), kProbe - 1);
  int c = (iz * kProbe + iy) * kProbe + ix;
  // neighbouring points share probe cells: one atomic per distinct word per warp
  const int word = c >> 5;
  const uint32_t m = 1u << (c & 31);
  const unsigned peers = __match_any_sync(__activemask(), word);
  const uint32_t combined = __reduce_or_sync(peers, m);
  if ((threadIdx.x & 31) == __ffs(peers) - 1 && (bits[word] & combined) != combined) atomicOr(&bits[word], combined);
}
__global__ void __launch_bounds__(256) probe_count(const uint32_t* __restrict__ bits, int nwords,
                                                   uint32_t* out) {
  uint32_t c = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x)
    c += __popc(bits[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// Cell coordinate (in cells, float) — the single definition used for build and query.
__device__ __forceinline__ float cell_coord(float v, float o, float inv_c) {
  return __fmul_rn(__fsub_rn(v, o), inv_c);
}

// key = linear cell id (x fastest); non-finite points get the sentinel `ncell`.
// Also histograms the cells (-> cell_start by scan) and the 8^3 super-cells.
__global__ void __launch_bounds__(256)
    grid_keys(const float4* __restrict__ p, int n, GridDev g, uint32_t ncell,
              uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
              uint32_t* __restrict__ cell_cnt, uint32_t* __restrict__ coarse_cnt,
              uint32_t* __restrict__ occ_bits = nullptr, int occ_wx = 0, uint32_t* __restrict__ rank = nullptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 q = p[i];
  uint32_t key = ncell;
  uint32_t my_rank = 0;
  if (finite3(q.x, q.y, q.z)) {
    int ix = min(max((int)floorf(cell_coord(q.x, g.ox, g.inv_cx)), 0), g.dx - 1);
    int iy = min(max((int)floorf(cell_coord(q.y, g.oy, g.inv_c)), 0), g.dy - 1);
    int iz = min(max((int)floorf(cell_coord(q.z, g.oz, g.inv_c)), 0), g.dz - 1);
    key = (uint32_t)((iz * g.dy + iy) * g.dx + ix);
    // points arrive in scan order, so the lanes of a warp share a handful of cells: one
    // atomic per distinct cell (and per distinct super-cell) per warp instead of one per point
    const int lane = threadIdx.x & 31;
    const unsigned act = __activemask();
    if (cell_cnt) {
      // the value the atomic returns is this warp's first slot in the cell: a (run-dependent) rank
      // inside the cell for the counting sort; the order is made canonical afterwards
      const unsigned peers = __match_any_sync(act, key);
      const int leader = __ffs(peers) - 1;
      uint32_t base = 0;
      if (lane == leader) base = atomicAdd(&cell_cnt[key], (uint32_t)__popc(peers));
      base = __shfl_sync(peers, base, leader);
      my_rank = base + __popc(peers & ((1u << lane) - 1u));
    }
    if (coarse_cnt) {
      const uint32_t ck = (uint32_t)(((iz >> kCoarseShift) * g.cdy + (iy >> kCoarseShift)) * g.cdx +
                                     (ix >> (kCoarseShift + g.xs_shift)));
      const unsigned peers = __match_any_sync(act, ck);
      if (lane == __ffs(peers) - 1) atomicAdd(&coarse_cnt[ck], (uint32_t)__popc(peers));
    }
    if (occ_bits) {
      const int cx = ix >> g.xs_shift;
      const int word = (iz * g.dy + iy) * occ_wx + (cx >> 5);
      const uint32_t m = 1u << (cx & 31);
      const unsigned peers = __match_any_sync(act, word);
      const uint32_t combined = __reduce_or_sync(peers, m);
      if (lane == __ffs(peers) - 1 && (occ_bits[word] & combined) != combined) atomicOr(&occ_bits[word], combined);
    }
  } else if (rank) {
    my_rank = atomicAdd(&cell_cnt[ncell], 1u);  // non-finite points: sentinel cell, sorted last
  }
  keys[i] = key;
  if (rank)
    rank[i] = my_rank;
  else
    vals[i] = (uint32_t)i;
}

// ---- counting sort by cell: scatter with the atomic ranks, then canonical order ---------------
// tmp[start[key] + rank] = point index: groups the points by cell in a run-dependent order.
__global__ void __launch_bounds__(256)
    cs_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ start,
               int n, uint32_t* __restrict__ tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tmp[start[keys[i]] + rank[i]] = (uint32_t)i;
}
// Canonical (= stable) order inside every cell: the final slot of point i is its cell's start plus
// the number of points of the same cell with a smaller index (cells hold a handful of points, so
// the count is a short scan of the cell's segment); the point (and its normal) is written there.
// The layout is thereby identical to a stable sort by cell, whatever order the atomics ran in.
__global__ void __launch_bounds__(256)
    cs_fixup_gather(const uint32_t* __restrict__ tmp, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ start,
                    const float4* __restrict__ p, const float4* __restrict__ nrm, int n, float4* __restrict__ out_p,
                    float4* __restrict__ out_n, float4* __restrict__ out_p2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t i = tmp[t];
  const uint32_t key = keys[i];
  const uint32_t s = start[key], e = start[key + 1];
  uint32_t r = t - s;  // pathological cells (thousands of coincident points) keep the scattered order
  if (e - s <= 4096u) {
    r = 0;
    for (uint32_t u = s; u < e; ++u) r += tmp[u] < i ? 1u : 0u;
  }
  float4 q = p[i];
  q.w = __int_as_float((int)i);
  out_p[s + r] = q;
  if (out_p2) out_p2[s + r] = q;
  if (nrm) out_n[s + r] = nrm[i];
}

// Box dilation (Chebyshev radius r cells) of the occupancy bits: out bit = OR of all bits within r cells.
__global__ void __launch_bounds__(256)
    occ_dilate(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int wx, int dy, int dz, int r) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= wx * dy * dz) return;
  const int w = idx % wx, y = (idx / wx) % dy, z = idx / (wx * dy);
  uint32_t acc = 0;
  for (int zz = max(z - r, 0); zz <= min(z + r, dz - 1); ++zz)
    for (int yy = max(y - r, 0); yy <= min(y + r, dy - 1); ++yy) {
      const uint32_t* row = in + (size_t)(zz * dy + yy) * wx;
      const uint32_t a = row[w], l = w > 0 ? row[w - 1] : 0u, h = w + 1 < wx ? row[w + 1] : 0u;
      acc |= a;
      for (int k = 1; k <= r; ++k) acc |= (a << k) | (l >> (32 - k)) | (a >> k) | (h << (32 - k));
    }
  out[idx] = acc;
}

// true: no indexed point within (occ_r - 0.01) * c of the query (proven by the dilated occupancy)
__device__ __forceinline__ bool occ_proves_empty(const GridDev& g, int ix, int iy, int iz) {
  if (!g.occ) return false;
  const int cx = ix >> g.xs_shift, ncx = g.dx >> g.xs_shift;
  if ((unsigned)cx < (unsigned)ncx && (unsigned)iy < (unsigned)g.dy && (unsigned)iz < (unsigned)g.dz)
    return ((__ldg(&g.occ[(size_t)(iz * g.dy + iy) * g.occ_wx + (cx >> 5)]) >> (cx & 31)) & 1u) == 0u;
  // outside the grid: Chebyshev distance (cells) to the grid box
  const int ox = max(max(-cx, cx - (ncx - 1)), 0), oy = max(max(-iy, iy - (g.dy - 1)), 0), oz = max(max(-iz, iz - (g.dz - 1)), 0);
  return max(ox, max(oy, oz)) > g.occ_r;
}

__global__ void __launch_bounds__(256)
    gather_sorted(const float4* __restrict__ p, const float4* __restrict__ nrm,
                  const uint32_t* __restrict__ vals, int n, float4* __restrict__ out_p,
                  float4* __restrict__ out_n, float4* __restrict__ out_p2) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t i = vals[j];
  float4 q = p[i];
  q.w = __int_as_float((int)i);
  out_p[j] = q;
  if (out_p2) out_p2[j] = q;
  if (nrm) out_n[j] = nrm[i];
}

inline int bit_length(uint32_t v) {
  int b = 0;
  while (v) {
    ++b;
    v >>= 1;
  }
  return b;
}

// Scratch slots of ctx->scratch used by the index build.
enum { kScrBBox = 0, kScrProbe, kScrKeys, kScrVals, kScrKeysAlt, kScrValsAlt, kScrHist, kScrScan,
       kScrGridEnd };

// Builds the grid over `xyz` (n float4, input order).  cell_factor: cell edge in units of
// the estimated point spacing.  min_cell > 0 forces a lower bound on the cell edge.
// Two phases so that callers can overlap independent work with the second one:
//   grid_plan  bbox + density probe + ONE host round trip -> origin, cell edge, dims (everything
//              a query needs to compute cell coordinates, e.g. the Morton order of the source);
//   grid_fill  keys, sort, scan, gather (stream-ordered, no host sync).
// What a caller may already know about the cloud (lc3d_prepare_view after VoxelGrid: every point is
// finite, lies inside the input's bounding box and the spacing is about the leaf size): with a
// hint grid_plan runs no kernel and makes no host round trip.
struct GridHint {
  float lo[3], hi[3];  // a box that contains every point (not necessarily tight)
  int64_t nfinite;     // all n points must be finite
  double spacing;      // estimated point spacing
};
inline void grid_plan(lc3d_ctx* ctx, Grid& G, const float4* xyz, int64_t n64, double cell_factor,
                      double min_cell = 0.0, int xsub = 1, const GridHint* hint = nullptr) {
  const int n = (int)n64;
  cudaStream_t st = ctx->stream;
  G.v = GridDev{};
  if (n == 0) {
    G.v.dx = G.v.dy = G.v.dz = G.v.cdx = G.v.cdy = G.v.cdz = 1;
    G.v.c = G.v.inv_c = G.v.inv_cx = G.v.inv_xs = 1.0f;
    G.v.xs = 1;
    G.ncell = 1;
    G.cell_start.ensure(2 * 4);
    G.coarse_cnt.ensure(4);
    G.pts.ensure(16);
    LC3D_CUDA(cudaMemsetAsync(G.cell_start.p, 0, 8, st));
    LC3D_CUDA(cudaMemsetAsync(G.coarse_cnt.p, 0, 4, st));
    G.v.cell_start = G.cell_start.as<uint32_t>();
    G.v.coarse_cnt = G.coarse_cnt.as<uint32_t>();
    G.v.pts = G.pts.as<float4>();
    return;
  }
  BBoxOut bb{};
  int nfinite = 0;
  double ext[3];
  if (hint) {
    nfinite = (int)hint->nfinite;
    for (int d = 0; d < 3; ++d) {
      G.lo[d] = hint->lo[d];
      G.hi[d] = hint->hi[d];
    }
  } else {
    // 1. bounding box of the finite points
    ctx->scratch[kScrBBox].ensure(sizeof(BBoxOut) + 16);
    BBoxOut* d_bb = ctx->scratch[kScrBBox].as<BBoxOut>();
    LC3D_LAUNCH(ctx, bbox_init, 1, 32, 0, d_bb);
    int nb = std::min(div_up(n, 256), ctx->num_sms * 2);
    LC3D_LAUNCH(ctx, bbox_reduce, nb, 256, 0, xyz, n, d_bb);
    // 2. density probe (64^3 occupancy over the bbox, bbox read on the device), then ONE round trip
    const int nwords = kProbe * kProbe * kProbe / 32;
    ctx->scratch[kScrProbe].ensure(nwords * 4 + 16);
    uint32_t* bits = ctx->scratch[kScrProbe].as<uint32_t>();
    LC3D_CUDA(cudaMemsetAsync(bits, 0, nwords * 4 + 16, st));
    LC3D_LAUNCH(ctx, probe_mark, div_up(n, 256), 256, 0, xyz, n, d_bb, bits);
    LC3D_LAUNCH(ctx, probe_count, 32, 256, 0, bits, nwords, &d_bb->pad);
    LC3D_CUDA(cudaMemcpyAsync(&bb, d_bb, sizeof bb, cudaMemcpyDeviceToHost, st));
    LC3D_CUDA(cudaStreamSynchronize(st));
    nfinite = (int)bb.count;
    if (nfinite == 0) {
      for (int d = 0; d < 3; ++d) G.lo[d] = G.hi[d] = 0.0f;
    } else {
      for (int d = 0; d < 3; ++d) {
        G.lo[d] = ord2f(bb.lo[d]);
        G.hi[d] = ord2f(bb.hi[d]);
      }
    }
  }
  double maxext = 0;
  for (int d = 0; d < 3; ++d) {
    ext[d] = (double)G.hi[d] - (double)G.lo[d];
    maxext = std::max(maxext, ext[d]);
  }
  // point spacing (surface area ~ occupied probe cells) -> cell edge
  double cell, spacing_est = 0.0;
  if (nfinite <= 1 || maxext <= 0) {
    cell = maxext > 0 ? maxext : 1.0;
  } else if (hint && hint->spacing > 0) {
    spacing_est = hint->spacing;
    cell = cell_factor * spacing_est;
  } else {
    double pe[3];
    for (int d = 0; d < 3; ++d) pe[d] = std::max(ext[d], maxext * 1e-6) / kProbe;
    const uint32_t occ = bb.pad;
    double cell_area = std::pow(pe[0] * pe[1] * pe[2], 2.0 / 3.0);
    double area = std::max(1.0, (double)occ) * cell_area;
    double spacing = std::sqrt(area / (double)nfinite);
    spacing_est = spacing;
    cell = cell_factor * spacing;
  }
  cell = std::max(cell, min_cell);
  cell = std::max(cell, maxext / (kMaxDim - 2));
  cell = std::max(cell, 1e-30);
  int dims[3];
  int xs = 1, xs_shift = 0;
  while (xs * 2 <= xsub && xs < 16) {
    xs *= 2;
    ++xs_shift;
  }
  const int64_t kMaxCells = max_cells_for(n64);
  for (int iter = 0; iter < 64; ++iter) {
    int64_t tot = 1;
    for (int d = 0; d < 3; ++d) {
      dims[d] = (int)std::floor(ext[d] / cell) + 1;
      tot *= dims[d];
    }
    // the x subdivision must keep dx within the range the cell-coordinate slack covers
    while (xs > 1 && ((int64_t)dims[0] * xs > kMaxDimX || tot * xs > kMaxCells)) {
      xs >>= 1;
      --xs_shift;
    }
    if (tot * xs <= kMaxCells) break;
    cell *= std::cbrt((double)tot / (double)kMaxCells) * 1.02;
  }
  GridDev& g = G.v;
  g.ox = G.lo[0];
  g.oy = G.lo[1];
  g.oz = G.lo[2];
  g.c = (float)cell;
  g.inv_c = 1.0f / g.c;
  g.xs = xs;
  g.xs_shift = xs_shift;
  g.inv_xs = 1.0f / (float)xs;
  g.inv_cx = g.inv_c * (float)xs;  // exact (power of two)
  // x-subcells: every c-cell of the isotropic layout splits into xs slices
  g.dx = dims[0] * xs;
  g.dy = dims[1];
  g.dz = dims[2];
  g.cdx = (dims[0] + kCoarse - 1) >> kCoarseShift;
  g.cdy = (g.dy + kCoarse - 1) >> kCoarseShift;
  g.cdz = (g.dz + kCoarse - 1) >> kCoarseShift;
  g.n = nfinite;
  G.ncell = (int64_t)g.dx * g.dy * g.dz;
  G.pts_per_cell_est = spacing_est > 0 ? (cell / spacing_est) * (cell / spacing_est) / (double)xs : (double)nfinite;
}

// before_gather: called (stream-ordered on ctx->stream) right before the normals are first read,
// so a caller may still be uploading them while the keys are sorted.
inline void grid_fill(lc3d_ctx* ctx, Grid& G, const float4* xyz, const float4* nrm, int64_t n64,
                      const std::function<void()>& before_gather = nullptr, double occ_reach = 0.0) {
  const int n = (int)n64;
  if (n == 0) return;  // grid_plan set up the empty grid
  cudaStream_t st = ctx->stream;
  GridDev& g = G.v;
  const int64_t ncoarse = (int64_t)g.cdx * g.cdy * g.cdz;
  // 3. keys + histograms
  G.cell_start.ensure((size_t)(G.ncell + 4) * 4);
  G.coarse_cnt.ensure((size_t)ncoarse * 4);
  G.pts.ensure((size_t)n * 16 + 16);
  if (nrm) G.nrm.ensure((size_t)n * 16 + 16);
  uint32_t* cell_start = G.cell_start.as<uint32_t>();
  LC3D_CUDA(cudaMemsetAsync(cell_start, 0, (size_t)(G.ncell + 4) * 4, st));
  LC3D_CUDA(cudaMemsetAsync(G.coarse_cnt.p, 0, (size_t)ncoarse * 4, st));
  ctx->scratch[kScrKeys].ensure((size_t)n * 4);
  ctx->scratch[kScrVals].ensure((size_t)n * 4);
  ctx->scratch[kScrKeysAlt].ensure((size_t)n * 4);
  ctx->scratch[kScrValsAlt].ensure((size_t)n * 4);
  ctx->scratch[kScrHist].ensure(sort_hist_bytes(n));
  ctx->scratch[kScrScan].ensure(
      std::max(scan_scratch_bytes(G.ncell + 4), scan_scratch_bytes((int64_t)kRadix * div_up(n, kSortTile))) + 64);
  uint32_t* keys = ctx->scratch[kScrKeys].as<uint32_t>();
  uint32_t* vals = ctx->scratch[kScrVals].as<uint32_t>();
  // dilated occupancy (rejects queries with nothing within occ_reach in O(1)): only for small radii
  g.occ = nullptr;
  g.occ_wx = g.occ_r = 0;
  uint32_t* occ_raw = nullptr;
  int occ_r = 0;
  const int occ_wx = ((g.dx >> g.xs_shift) + 31) >> 5;
  const size_t occ_words = (size_t)occ_wx * g.dy * g.dz;
  if (occ_reach > 0.0 && std::isfinite(occ_reach)) occ_r = (int)std::ceil(occ_reach / (double)g.c + 0.02);
  if (occ_r >= 1 && occ_r <= 4 && occ_words <= ((size_t)1 << 24)) {
    G.occ_raw.ensure(occ_words * 4);
    G.occ.ensure(occ_words * 4);
    occ_raw = G.occ_raw.as<uint32_t>();
    LC3D_CUDA(cudaMemsetAsync(occ_raw, 0, occ_words * 4, st));
  } else {
    occ_r = 0;
  }
  // Counting sort (one histogram pass with atomic ranks, scan, scatter, canonical in-cell order)
  // when cells hold a handful of points — the ICP / k-NN indexes; the multi-pass radix sort when a
  // forced minimum cell edge packs hundreds of points into a cell (the in-cell ordering step is
  // quadratic in the cell population).
  const bool counting = G.pts_per_cell_est <= 192.0 && !std::getenv("LC3D_RADIX_INDEX");
  LC3D_LAUNCH(ctx, grid_keys, div_up(n, 256), 256, 0, xyz, n, g, (uint32_t)G.ncell, keys, vals,
              cell_start, G.coarse_cnt.as<uint32_t>(), occ_raw, occ_wx, counting ? vals : (uint32_t*)nullptr);
  if (counting) {
    uint32_t* rank = vals;
    uint32_t* tmp = ctx->scratch[kScrKeysAlt].as<uint32_t>();
    exclusive_scan_u32(ctx, cell_start, cell_start, G.ncell + 2, ctx->scratch[kScrScan].as<uint32_t>());
    LC3D_LAUNCH(ctx, cs_scatter, div_up(n, 256), 256, 0, keys, rank, cell_start, n, tmp);
    if (occ_raw) {
      LC3D_LAUNCH(ctx, occ_dilate, div_up((int64_t)occ_words, 256), 256, 0, occ_raw, G.occ.as<uint32_t>(), occ_wx, g.dy,
                  g.dz, occ_r);
      g.occ = G.occ.as<uint32_t>();
      g.occ_wx = occ_wx;
      g.occ_r = occ_r;
    }
    if (before_gather) before_gather();
    LC3D_LAUNCH(ctx, cs_fixup_gather, div_up(n, 256), 256, 0, tmp, keys, cell_start, xyz, nrm, n, G.pts.as<float4>(),
                nrm ? G.nrm.as<float4>() : nullptr, (float4*)nullptr);
    g.cell_start = cell_start;
    g.coarse_cnt = G.coarse_cnt.as<uint32_t>();
    g.pts = G.pts.as<float4>();
    g.nrm = nrm ? G.nrm.as<float4>() : nullptr;
    return;
  }
  if (occ_raw) {
    LC3D_LAUNCH(ctx, occ_dilate, div_up((int64_t)occ_words, 256), 256, 0, occ_raw, G.occ.as<uint32_t>(), occ_wx, g.dy, g.dz,
                occ_r);
    g.occ = G.occ.as<uint32_t>();
    g.occ_wx = occ_wx;
    g.occ_r = occ_r;
  }
  // 4. stable radix sort of (cell id, point index)
  SortScratch ss{ctx->scratch[kScrKeysAlt].as<uint32_t>(), ctx->scratch[kScrValsAlt].as<uint32_t>(),
                 ctx->scratch[kScrHist].as<uint32_t>(), ctx->scratch[kScrScan].as<uint32_t>()};
  uint32_t *skeys = nullptr, *svals = nullptr;
  radix_sort_pairs(ctx, keys, vals, n, bit_length((uint32_t)G.ncell), ss, &skeys, &svals);
  vals = svals;
  // 5. cell_start = exclusive scan of the cell histogram (ncell+1 entries)
  exclusive_scan_u32(ctx, cell_start, cell_start, G.ncell + 1, ctx->scratch[kScrScan].as<uint32_t>());
  // 6. gather into sorted SoA float4 (finite points come first: sentinel key sorts last)
  if (before_gather) before_gather();
  LC3D_LAUNCH(ctx, gather_sorted, div_up(n, 256), 256, 0, xyz, nrm, vals, n, G.pts.as<float4>(),
              nrm ? G.nrm.as<float4>() : nullptr, (float4*)nullptr);
  g.cell_start = cell_start;
  g.coarse_cnt = G.coarse_cnt.as<uint32_t>();
  g.pts = G.pts.as<float4>();
  g.nrm = nrm ? G.nrm.as<float4>() : nullptr;
}

// Target normals gathered into the index order AFTER the fill (host-buffer ICP: the index is built
// and the first search runs while the normals are still crossing PCIe).  pts[j].w = original index.
__global__ void __launch_bounds__(256)
    gather_normals_sorted(const float4* __restrict__ pts, const float4* __restrict__ nrm_in, int n,
                          float4* __restrict__ nrm_sorted) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  nrm_sorted[j] = nrm_in[__float_as_int(pts[j].w)];
}
inline void grid_attach_normals(lc3d_ctx* ctx, Grid& G, const float4* nrm, int64_t n64) {
  const int n = (int)n64;  // every slot of the sorted array carries its original index in w
  if (n == 0 || G.v.n == 0) return;
  G.nrm.ensure((size_t)n64 * 16 + 16);
  LC3D_LAUNCH(ctx, gather_normals_sorted, div_up(n, 256), 256, 0, G.v.pts, nrm, n, G.nrm.as<float4>());
  G.v.nrm = G.nrm.as<float4>();
}

inline void grid_build(lc3d_ctx* ctx, Grid& G, const float4* xyz, const float4* nrm, int64_t n64,
                       double cell_factor, double min_cell = 0.0, int xsub = 1, const GridHint* hint = nullptr) {
  grid_plan(ctx, G, xyz, n64, cell_factor, min_cell, xsub, hint);
  grid_fill(ctx, G, xyz, nrm, n64);
}

// Orders `xyz` (n float4) along a Morton (Z-order) curve over the cells of grid g (clamped),
// for query locality: G consecutive queries form a compact 3-D patch, so the lanes of a
// tile group share one small search region.  (The TARGET stays in linear x-fastest cell
// order: that is what makes x-runs of cells contiguous.)  out[j].w = bits(orig idx).
__device__ __forceinline__ uint32_t morton_spread10(uint32_t v) {
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
// hist != null: counting-sort mode — histogram of the keys with the atomic's return value as the
// (run-dependent) rank inside the key, written to vals; otherwise vals = point index (radix mode).
__global__ void __launch_bounds__(256)
    query_keys(const float4* __restrict__ p, int n, GridDev g, int shift, uint32_t sentinel,
               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ hist = nullptr) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 q = p[i];
  uint32_t key = sentinel;  // non-finite queries sort last
  if (finite3(q.x, q.y, q.z)) {
    int ix = min(max((int)floorf(cell_coord(q.x, g.ox, g.inv_cx)), 0), g.dx - 1) >> g.xs_shift;
    int iy = min(max((int)floorf(cell_coord(q.y, g.oy, g.inv_c)), 0), g.dy - 1);
    int iz = min(max((int)floorf(cell_coord(q.z, g.oz, g.inv_c)), 0), g.dz - 1);
    key = morton_spread10((uint32_t)ix >> shift) | (morton_spread10((uint32_t)iy >> shift) << 1) |
          (morton_spread10((uint32_t)iz >> shift) << 2);
  }
  keys[i] = key;
  if (hist) {
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(__activemask(), key);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&hist[key], (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    vals[i] = base + __popc(peers & ((1u << lane) - 1u));
  } else {
    vals[i] = (uint32_t)i;
  }
}

// Needs only grid_plan's result.  scr: first of 6 consecutive scratch slots (keys, vals, alt
// keys, alt vals, histogram, scan) — pass a private range to run concurrently with grid_fill.
inline void sort_queries_by_cell(lc3d_ctx* ctx, const Grid& G, const float4* xyz, int64_t n64,
                                 float4* out_sorted, float4* out_sorted2 = nullptr, int scr = kScrKeys) {
  const int n = (int)n64;
  if (n == 0) return;
  const int sKeys = scr, sVals = scr + 1, sKeysAlt = scr + 2, sValsAlt = scr + 3, sHist = scr + 4, sScan = scr + 5;
  ctx->scratch[sKeys].ensure((size_t)n * 4);
  ctx->scratch[sVals].ensure((size_t)n * 4);
  ctx->scratch[sKeysAlt].ensure((size_t)n * 4);
  ctx->scratch[sValsAlt].ensure((size_t)n * 4);
  ctx->scratch[sHist].ensure(sort_hist_bytes(n));
  ctx->scratch[sScan].ensure(scan_scratch_bytes((int64_t)kRadix * div_up(n, kSortTile)) + 64);
  uint32_t* keys = ctx->scratch[sKeys].as<uint32_t>();
  uint32_t* vals = ctx->scratch[sVals].as<uint32_t>();
  int maxdim = std::max(G.v.dx >> G.v.xs_shift, std::max(G.v.dy, G.v.dz));
  int shift = 0;
  while ((maxdim >> shift) > 1024) ++shift;  // 10 bits per axis
  // Counting sort over the Morton cells (histogram with atomic ranks, scan, scatter, canonical
  // in-cell order: 5 launches instead of the 11 of a 3-pass radix sort).  The key table has
  // 2^(3 bits-per-axis) entries, so the Morton cells are coarsened until it fits 2^22; cells that
  // would then hold hundreds of points fall back to the radix sort.
  int cshift = shift;
  while (3 * bit_length((uint32_t)((maxdim - 1) >> cshift)) > 22) ++cshift;
  const double per_key = G.pts_per_cell_est * (double)G.v.xs * std::pow(4.0, cshift);  // a surface: ~4^shift cells merge
  if (per_key <= 256.0 && !std::getenv("LC3D_RADIX_INDEX")) {
    const int cbc = bit_length((uint32_t)((maxdim - 1) >> cshift));
    const int64_t nkeys = ((int64_t)1 << (3 * cbc)) + 1;  // + the non-finite sentinel
    ctx->scratch[sHist].ensure((size_t)(nkeys + 4) * 4);
    ctx->scratch[sScan].ensure(scan_scratch_bytes(nkeys + 4) + 64);
    uint32_t* table = ctx->scratch[sHist].as<uint32_t>();
    uint32_t* tmp = ctx->scratch[sKeysAlt].as<uint32_t>();
    LC3D_CUDA(cudaMemsetAsync(table, 0, (size_t)(nkeys + 4) * 4, ctx->stream));
    LC3D_LAUNCH(ctx, query_keys, div_up(n, 256), 256, 0, xyz, n, G.v, cshift, (uint32_t)(nkeys - 1), keys, vals, table);
    exclusive_scan_u32(ctx, table, table, nkeys + 1, ctx->scratch[sScan].as<uint32_t>());
    LC3D_LAUNCH(ctx, cs_scatter, div_up(n, 256), 256, 0, keys, vals, table, n, tmp);
    LC3D_LAUNCH(ctx, cs_fixup_gather, div_up(n, 256), 256, 0, tmp, keys, table, xyz, (const float4*)nullptr, n, out_sorted,
                (float4*)nullptr, out_sorted2);
    return;
  }
  const int cb0 = bit_length((uint32_t)((maxdim - 1) >> shift));
  LC3D_LAUNCH(ctx, query_keys, div_up(n, 256), 256, 0, xyz, n, G.v, shift, (uint32_t)1u << (3 * cb0), keys, vals,
              (uint32_t*)nullptr);
  SortScratch ss{ctx->scratch[sKeysAlt].as<uint32_t>(), ctx->scratch[sValsAlt].as<uint32_t>(),
                 ctx->scratch[sHist].as<uint32_t>(), ctx->scratch[sScan].as<uint32_t>()};
  // key bits actually used: 3 interleaved coordinates of bit_length((maxdim-1) >> shift) bits,
  // plus the non-finite sentinel bit 30 only if such points can exist (checked by the caller's
  // bbox count): sorting all 31 bits costs a fourth radix pass for nothing
  const int cb = bit_length((uint32_t)((maxdim - 1) >> shift));
  uint32_t *skeys = nullptr, *svals = nullptr;
  radix_sort_pairs(ctx, keys, vals, n, std::min(31, 3 * cb + 1), ss, &skeys, &svals);
  vals = svals;
  LC3D_LAUNCH(ctx, gather_sorted, div_up(n, 256), 256, 0, xyz, (const float4*)nullptr, vals, n,
              out_sorted, (float4*)nullptr, out_sorted2);
}

}  // namespace lc3d
