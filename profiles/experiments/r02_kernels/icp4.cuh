// icp4.cuh — the ICP iteration as a pipeline of small dense kernels (fourth generation).
//
// One iteration of pcl::IterativeClosestPoint::align (pcl_tools/fine_registration.cpp:121;
// SURVEY A.2-A.5) = five launches chained with programmatic dependent launch, then
// icp_solve_kernel (icp.cuh):
//   icp4_prepare   one thread per source point: incremental float32 transform; the previous match
//                  is re-evaluated (an exact upper bound on the nearest-neighbour distance);
//                  points that provably have nothing within the gate are finished here (proven-
//                  empty radius minus the accumulated motion, or the dilated occupancy bits of
//                  the index).  Everything else goes to work queue 0.
//   icp4_pass<0>   queue 0: the 2x2 cell rows nearest to the point, +-1 cell along x
//   icp4_pass<1>   queue 1: the 4x4 nearest rows, +-2 cells along x
//   icp4_pass<2>   queue 2: every row that intersects the search ball; balls wider than the row
//                  table go through the warp-cooperative ring search
//                  A pass resolves a point when the rows and the x-range it examined cover the
//                  ball of the best distance found — exact by construction — and forwards it to
//                  the next queue otherwise.  Each pass is a dense grid-stride kernel over its
//                  queue (one thread per queued point, no block barriers, high occupancy): the
//                  cheap majority never waits for the expensive few, and the expensive few are
//                  spread over the whole GPU.
//   icp4_reduce    one thread per source point in the original order: estimator terms staged in
//                  shared memory, lane v of each warp accumulates estimator value v over the
//                  warp's 32 points in a fixed order (fp64 FMAs of exactly representable
//                  products), one partial row per block -> bit-reproducible sums whatever order
//                  the queues were filled in.
// All passes walk the precomputed centre-out row table (icp2.cuh) with walk_rows (icp3.cuh).
#pragma once
#include "icp3.cuh"

namespace lc3d {

struct Icp4Queues {
  int* q[3];          // work queues (source point indices)
  unsigned* count;    // [4]: lengths of the three queues (+ spare); reset by the solve kernel
};

constexpr int kI4PrepThreads = 256;
#ifndef LC3D_I4_PASS_THREADS
#define LC3D_I4_PASS_THREADS 128
#endif
#ifndef LC3D_I4_PASS_MINBLOCKS
#define LC3D_I4_PASS_MINBLOCKS 8
#endif
constexpr int kI4PassThreads = LC3D_I4_PASS_THREADS;

// warp-aggregated append to a global queue
__device__ __forceinline__ void queue_push(bool p, int* __restrict__ q, unsigned* __restrict__ count, int v) {
  const unsigned m = __ballot_sync(0xffffffffu, p);
  if (m) {
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (p) q[base + __popc(m & ((1u << lane) - 1u))] = v;
  }
}

// W[i] = (match position or -1, bits of its squared distance / of the bound, original index of
// the match, bits of the proven-empty radius to keep) — the working state of this iteration.
__global__ void __launch_bounds__(kI4PrepThreads)
    icp4_prepare(const IcpState* __restrict__ st, const __grid_constant__ IcpConfig cfg,
                 const __grid_constant__ GridDev g, float4* __restrict__ X, const int2* __restrict__ MB,
                 int4* __restrict__ W, int n, const Icp4Queues qs) {
  __shared__ float sT[16];
  __shared__ int s_flags[2];
  pdl_wait();
  pdl_trigger();
  if (threadIdx.x == 0) {
    s_flags[0] = st->done;
    s_flags[1] = st->iter;
  }
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T[threadIdx.x];
  __syncthreads();
  if (s_flags[0]) return;
  const int iter = s_flags[1];
  const int i = blockIdx.x * kI4PrepThreads + threadIdx.x;
  bool active = i < n;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  int2 mb = make_int2(-1, 0);
  if (active) q = X[i];
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
  float Lb = fmaxf(__int_as_float(mb.y) - delta, 0.0f);
  Best b;
  b.d2 = cfg.gate_ext;
  b.j = -1;
  b.oi = 0x7fffffff;
  bool need = active && g.n > 0;
  if (need) {
    if (mj >= 0) {
      consider(__ldg(&g.pts[mj]), mj, q.x, q.y, q.z, b);  // previous match: an upper bound
    } else if (Lb * 0.9999f > cfg.gate_dist) {
      need = false;  // still nothing within the gate
    }
    if (need && b.j < 0) {
      // no candidate at all: the dilated occupancy may prove that nothing lies within the gate
      const QueryCell qc = query_cell(g, q.x, q.y, q.z);
      if (occ_proves_empty(g, qc.ix, qc.iy, qc.iz)) {
        need = false;
        Lb = ((float)g.occ_r - 0.01f) * g.c;
      }
    }
  }
  if (i < n) W[i] = make_int4(b.j, __float_as_int(b.d2), b.oi, __float_as_int(need ? 0.0f : Lb));
  queue_push(need, qs.q[0], qs.count + 0, i);
}

template <int PASS, bool STATS>
__global__ void __launch_bounds__(kI4PassThreads, LC3D_I4_PASS_MINBLOCKS)
    icp4_pass(const IcpState* __restrict__ st, const __grid_constant__ IcpConfig cfg,
              const __grid_constant__ GridDev g, const float4* __restrict__ X, int4* __restrict__ W,
              const Icp4Queues qs, const int2* __restrict__ rowtab) {
  pdl_wait();
  pdl_trigger();
  if (st->done) return;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned count = *(volatile unsigned*)(qs.count + PASS);
  const int* __restrict__ qin = qs.q[PASS];
  // warps take contiguous chunks of 32 queue entries (appended together: Morton neighbours)
  const unsigned wstride = gridDim.x * (kI4PassThreads / 32) * 32;
  for (unsigned e0 = (blockIdx.x * (kI4PassThreads / 32) + (threadIdx.x >> 5)) * 32; e0 < count; e0 += wstride) {
    const unsigned e = e0 + lane;
    const bool act = e < count;
    const int i = act ? qin[e] : 0;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 w4 = make_int4(-1, 0, 0x7fffffff, 0);
    if (act) {
      q = X[i];
      w4 = W[i];
    }
    Best b;
    b.j = w4.x;
    b.d2 = __int_as_float(w4.y);
    b.oi = w4.z;
    bool ok;
    if (PASS == 0) {
      ok = walk_rows<false>(g, rowtab, act, q.x, q.y, q.z, 4, 1, b, 0.f, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    } else if (PASS == 1) {
      ok = walk_rows<false>(g, rowtab, act, q.x, q.y, q.z, 16, 2, b, 0.f, nullptr, 0, nullptr, nullptr, nullptr, nullptr);
    } else {
      const float Rc = sqrtf(b.d2) * g.inv_c * 1.0001f + 0.01f;  // ball radius in cells
      const bool walk = act && Rc <= cfg.tab_wmax;
      ok = false;
      if (__any_sync(full, walk)) {
        const bool okw = walk_rows<false>(g, rowtab, walk, q.x, q.y, q.z, kTabN, 1 << 20, b, 0.f, nullptr, 0, nullptr,
                                          nullptr, nullptr, nullptr);
        if (walk) ok = okw;
      }
      // balls wider than the row table (huge gates, no gate): warp-cooperative ring search
      unsigned todo = __ballot_sync(full, act && !ok);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        Best wb;
        wb.d2 = __shfl_sync(full, b.d2, src);
        wb.j = __shfl_sync(full, b.j, src);
        wb.oi = __shfl_sync(full, b.oi, src);
        const float wx = __shfl_sync(full, q.x, src), wy = __shfl_sync(full, q.y, src), wz = __shfl_sync(full, q.z, src);
        nn_phase2_warp(g, wx, wy, wz, wb);
        if (lane == src) {
          b = wb;
          ok = true;
        }
      }
    }
    if (act) {
      // resolved without a match: nothing lies within r_cap (the bound the search ran against)
      const float L = (ok && b.j < 0) ? cfg.r_cap * 0.9999f : 0.0f;
      W[i] = make_int4(b.j, __float_as_int(b.d2), b.oi, __float_as_int(L));
    }
    if (PASS < 2) queue_push(act && !ok, qs.q[PASS + 1], qs.count + PASS + 1, i);
  }
}

template <int MODE>
__global__ void __launch_bounds__(kI3Threads)
    icp4_reduce(const IcpState* __restrict__ st, const __grid_constant__ IcpConfig cfg,
                const __grid_constant__ GridDev g, const float4* __restrict__ X, const int4* __restrict__ W,
                int2* __restrict__ MB, int n, double* __restrict__ partials, int32_t* __restrict__ dump_idx,
                float* __restrict__ dump_d2) {
  constexpr int NV = MODE == LC3D_ICP_POINT_TO_PLANE ? kNvP2Plane : kNvP2P;
  constexpr int kWarps = kI3Threads / 32;
  const unsigned full = 0xffffffffu;
  __shared__ double s_U[kWarps][32][kUW];
  __shared__ double s_part[kWarps][32];
  pdl_wait();
  pdl_trigger();
  if (st->done) return;
  const int iter = st->iter;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int i = blockIdx.x * kI3Threads + tid;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  int4 w4 = make_int4(-1, 0, 0, 0);
  if (i < n) {
    q = X[i];
    w4 = W[i];
    MB[i] = make_int2(w4.x, w4.w);
  }
  const int j = w4.x;
  const float d2 = __int_as_float(w4.y);
  const bool has = j >= 0 && d2 <= cfg.gate;
  float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 nn = make_float4(0.f, 0.f, 0.f, 0.f);
  if (has) {
    d = __ldg(&g.pts[j]);
    if (MODE == LC3D_ICP_POINT_TO_PLANE) nn = __ldg(&g.nrm[j]);
  }
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
      if (has && finite3(nn.x, nn.y, nn.z)) {
        // float32 products widened to double, as TransformationEstimationPointToPlaneLLS
        J[0] = nn.z * q.y - nn.y * q.z;
        J[1] = nn.x * q.z - nn.z * q.x;
        J[2] = nn.y * q.x - nn.x * q.y;
        J[3] = nn.x;
        J[4] = nn.y;
        J[5] = nn.z;
        r = nn.x * d.x + nn.y * d.y + nn.z * d.z - nn.x * q.x - nn.y * q.y - nn.z * q.z;
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
    for (int ww = 0; ww < kWarps; ++ww) s += s_part[ww][lane];
    partials[(size_t)lane * gridDim.x + blockIdx.x] = s;
  }
}

}  // namespace lc3d
