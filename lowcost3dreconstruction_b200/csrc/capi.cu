// capi.cu — the C ABI of include/lc3d.h over the CUDA kernels.  No CPU fallback.
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <functional>
#include <limits>

#include "common.cuh"
#include "dedup.cuh"
#include "cluster.cuh"
#include "grid.cuh"
#include "icp.cuh"
#include "knn.cuh"
#include "scan_sort.cuh"
#include "search.cuh"
#include "voxel.cuh"

using namespace lc3d;

namespace {

thread_local std::string g_create_error;

// scratch slots beyond the grid builder's
enum {
  kScrRawA = kScrGridEnd, kScrRawB, kScrRawC, kScrRawD, kScrSrcSorted, kScrSrcWork, kScrState, kScrPartials, kScrOutA,
  kScrOutB, kScrDumpIdx, kScrDumpD2, kScrBound, kScrPrevMatch, kScrMisc, kScrMisc2, kScrMisc3,
  kScrSrcSort0, kScrSrcSort1, kScrSrcSort2, kScrSrcSort3, kScrSrcSort4, kScrSrcSort5,  // private sort scratch
  kScrEnd
};
static_assert(kScrEnd <= 64, "scratch slots");

struct Guard {  // sets the device for the duration of a call
  explicit Guard(lc3d_ctx* c) { LC3D_CUDA(cudaSetDevice(c->device)); }
};

// Strided host floats -> device float4.  elems = 3 (xyz -> w=1) or 3+1 (normal + curvature).
__global__ void __launch_bounds__(256)
    unpack_strided(const unsigned char* __restrict__ raw, int64_t stride, int n, float w_default,
                   const unsigned char* __restrict__ raw_w, int64_t stride_w,
                   float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(raw + (size_t)i * stride);
  float4 v;
  v.x = p[0];
  v.y = p[1];
  v.z = p[2];
  v.w = raw_w ? *reinterpret_cast<const float*>(raw_w + (size_t)i * stride_w) : w_default;
  out[i] = v;
}

// Copies n records of `rec` bytes at `stride` from host memory into a device raw buffer
// (on stream `cs`) and returns the device pointer.
const unsigned char* stage_raw(lc3d_ctx* ctx, DevBuf& buf, const void* host, int64_t stride,
                               int64_t n, int rec, cudaStream_t cs = nullptr) {
  if (n == 0) return nullptr;
  size_t bytes = (size_t)(n - 1) * stride + rec;
  buf.ensure(bytes + 64);
  LC3D_CUDA(cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, cs ? cs : ctx->stream));
  ctx->h2d_bytes_call += bytes;
  return buf.as<unsigned char>();
}

// Pageable (unregistered) host memory?  cudaMemcpyAsync from it is staged by the driver at a
// fraction of the PCIe rate and blocks the calling thread.
bool host_is_pageable(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

constexpr int64_t kPackMinPoints = 131072;  // below this the plain copy wins (pool wake-up, per-chunk calls)
constexpr int64_t kPackChunkPoints = 32768; // 384 KiB of packed floats per DMA

// n records of 12 bytes at `stride` in PAGEABLE host memory -> device, packed (stride 12): host
// threads gather chunk after chunk into pinned staging memory and each finished chunk goes out
// with its own cudaMemcpyAsync on `cs`, so the DMA of one chunk overlaps the packing of the next
// and a 48-byte PCL point costs 12 bytes of PCIe per field instead of 48.  Returns the device
// pointer; the caller's buffer has been consumed when this returns.
const unsigned char* stage_packed(lc3d_ctx* ctx, int slot, DevBuf& buf, const void* host, int64_t stride, int64_t n,
                                  cudaStream_t cs) {
  if (n == 0) return nullptr;
  cudaStream_t st = cs ? cs : ctx->stream;
  buf.ensure((size_t)n * 12 + 64);
  if (!ctx->host_pool) {
    int t = 3;
    if (const char* e = std::getenv("LC3D_PACK_THREADS")) t = std::max(0, std::atoi(e) - 1);
    ctx->host_pool = new HostPool(std::min(t, std::max(0, (int)std::thread::hardware_concurrency() - 1)));
  }
  if (!ctx->stage_done[slot]) LC3D_CUDA(cudaEventCreateWithFlags(&ctx->stage_done[slot], cudaEventDisableTiming));
  else LC3D_CUDA(cudaEventSynchronize(ctx->stage_done[slot]));  // the slot's previous DMA has left it
  ctx->stage[slot].ensure((size_t)n * 12);
  unsigned char* pin = ctx->stage[slot].as<unsigned char>();
  unsigned char* dev = buf.as<unsigned char>();
  const unsigned char* src = static_cast<const unsigned char*>(host);
  const int njobs = (int)((n + kPackChunkPoints - 1) / kPackChunkPoints);
  const int device = ctx->device;
  std::atomic<int> failed{0};
  ctx->host_pool->run(njobs, [&](int j) {
    const int64_t a = (int64_t)j * kPackChunkPoints, b = std::min(n, a + kPackChunkPoints);
    if (stride == 12) {
      std::memcpy(pin + a * 12, src + a * 12, (size_t)(b - a) * 12);
    } else {
      for (int64_t i = a; i < b; ++i) std::memcpy(pin + i * 12, src + i * stride, 12);
    }
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaMemcpyAsync(dev + a * 12, pin + a * 12, (size_t)(b - a) * 12, cudaMemcpyHostToDevice, st) != cudaSuccess)
      failed.store(1);
  });
  if (failed.load()) {
    cudaGetLastError();
    throw CudaError{"staged upload failed"};
  }
  LC3D_CUDA(cudaEventRecord(ctx->stage_done[slot], st));
  ctx->h2d_bytes_call += (size_t)n * 12;
  return dev;
}

// An upload in flight: the raw strided bytes are copied on a (possibly separate) copy stream;
// the unpack into SoA float4 happens later on the compute stream.
struct PendingUpload {
  const lc3d_cloud* h = nullptr;
  lc3d_dcloud* d = nullptr;
  const unsigned char* raw_xyz = nullptr;
  const unsigned char* raw_nrm = nullptr;
  cudaEvent_t ready = nullptr;  // recorded on the copy stream after the last byte, or null
  // normals staged separately (upload_begin_normals) so that other arrays can go first
  bool normals_deferred = false;
  cudaEvent_t ready_nrm = nullptr;
  // strides of the STAGED records (12 when the host packed them, else the caller's)
  int64_t stride_xyz = 0, stride_nrm = 0;
  int slot = 0;        // staging slots slot (xyz) and slot + 1 (normals)
  bool packed = false; // pageable input packed by the host pool
};

PendingUpload upload_begin(lc3d_ctx* ctx, const lc3d_cloud* h, lc3d_dcloud* d, bool want_normals, DevBuf& raw_a,
                           DevBuf& raw_b, cudaStream_t cs, cudaEvent_t ready, bool defer_normals = false) {
  PendingUpload pu;
  pu.h = h;
  pu.d = d;
  const int64_t n = h->n;
  d->n = n;
  d->has_normal = false;
  if (n == 0) return pu;
  if (n > (int64_t)INT32_MAX / 2) throw CudaError{"cloud too large (n must be < 2^30)"};
  if (!h->xyz || h->xyz_stride < 12) throw CudaError{"cloud.xyz is NULL or xyz_stride < 12"};
  // the device unpacks the staged records with 4-byte loads: strides (and the offset between fields of
  // one record) must keep every float 4-byte aligned
  if (h->xyz_stride % 4 != 0) throw CudaError{"cloud.xyz_stride must be a multiple of 4 bytes"};
  d->xyz.ensure((size_t)n * 16);
  if (want_normals && h->normal) {
    if (h->normal_stride < 12) throw CudaError{"normal_stride < 12"};
    if (h->normal_stride % 4 != 0) throw CudaError{"cloud.normal_stride must be a multiple of 4 bytes"};
    d->normal.ensure((size_t)n * 16);
  }
  pu.slot = (&raw_a == &ctx->scratch[kScrRawA]) ? 0 : 2;
  pu.stride_xyz = h->xyz_stride;
  pu.stride_nrm = h->normal_stride;
  // pageable host memory (std::vector, numpy): the host pool packs every field into pinned staging
  // memory — 12 bytes per point and field over PCIe whatever the caller's record size
  // (packed 12-byte arrays gain nothing: the driver's own pageable staging runs at the same host
  // memcpy rate; records wider than the fields that are used — PCL's 48-byte points — halve their PCIe
  // bytes)
  int64_t pack_min = kPackMinPoints;
  if (const char* e = std::getenv("LC3D_PACK_MIN")) pack_min = std::max(1, std::atoi(e));
  pu.packed = n >= pack_min && h->xyz_stride > 16 && host_is_pageable(h->xyz) && !std::getenv("LC3D_NO_PACK");
  if (pu.packed) {
    pu.raw_xyz = stage_packed(ctx, pu.slot, raw_a, h->xyz, h->xyz_stride, n, cs);
    pu.stride_xyz = 12;
    if (want_normals && h->normal) {
      if (defer_normals) {
        pu.normals_deferred = true;
      } else {
        pu.raw_nrm = stage_packed(ctx, pu.slot + 1, raw_b, h->normal, h->normal_stride, n, cs);
        pu.stride_nrm = 12;
      }
      d->has_normal = true;
    }
    if (ready) {
      LC3D_CUDA(cudaEventRecord(ready, cs ? cs : ctx->stream));
      pu.ready = ready;
    }
    return pu;
  }
  // same AoS block as xyz (PCL 48-byte points)?  then one staged copy carries both
  const ptrdiff_t off = h->normal ? (const char*)h->normal - (const char*)h->xyz : -1;
  const bool same_block = want_normals && h->normal && h->normal_stride == h->xyz_stride && off >= 0 &&
                          off + 12 <= h->xyz_stride && off % 4 == 0;
  pu.raw_xyz = stage_raw(ctx, raw_a, h->xyz, h->xyz_stride, n, same_block ? (int)off + 12 : 12, cs);
  if (want_normals && h->normal) {
    if (same_block)
      pu.raw_nrm = pu.raw_xyz + off;
    else if (defer_normals)
      pu.normals_deferred = true;  // staged by upload_begin_normals
    else
      pu.raw_nrm = stage_raw(ctx, raw_b, h->normal, h->normal_stride, n, 12, cs);
    d->has_normal = true;  // contents arrive with upload_finish / upload_finish_normals
  }
  if (ready) {
    LC3D_CUDA(cudaEventRecord(ready, cs ? cs : ctx->stream));
    pu.ready = ready;
  }
  return pu;
}

// Second half of a split upload: the normals follow on the copy stream (after whatever the
// caller queued in between) with their own completion event.
void upload_begin_normals(lc3d_ctx* ctx, PendingUpload& pu, DevBuf& raw_b, cudaStream_t cs, cudaEvent_t ready) {
  if (!pu.normals_deferred || pu.h->n == 0) return;
  if (pu.packed) {
    pu.raw_nrm = stage_packed(ctx, pu.slot + 1, raw_b, pu.h->normal, pu.h->normal_stride, pu.h->n, cs);
    pu.stride_nrm = 12;
  } else {
    pu.raw_nrm = stage_raw(ctx, raw_b, pu.h->normal, pu.h->normal_stride, pu.h->n, 12, cs);
  }
  LC3D_CUDA(cudaEventRecord(ready, cs ? cs : ctx->stream));
  pu.ready_nrm = ready;
}
// Unpacks deferred normals on the CURRENT compute stream of the context once they have landed.
void upload_finish_normals(lc3d_ctx* ctx, const PendingUpload& pu) {
  if (!pu.normals_deferred || pu.h->n == 0) return;
  LC3D_CUDA(cudaStreamWaitEvent(ctx->stream, pu.ready_nrm, 0));
  LC3D_LAUNCH(ctx, unpack_strided, div_up(pu.h->n, 256), 256, 0, pu.raw_nrm, pu.stride_nrm, (int)pu.h->n,
              0.0f, (const unsigned char*)nullptr, (int64_t)0, pu.d->normal.as<float4>());
}

void upload_finish(lc3d_ctx* ctx, const PendingUpload& pu) {
  const int64_t n = pu.h->n;
  if (n == 0) return;
  if (pu.ready) LC3D_CUDA(cudaStreamWaitEvent(ctx->stream, pu.ready, 0));
  LC3D_LAUNCH(ctx, unpack_strided, div_up(n, 256), 256, 0, pu.raw_xyz, pu.stride_xyz, (int)n, 1.0f,
              (const unsigned char*)nullptr, (int64_t)0, pu.d->xyz.as<float4>());
  if (pu.raw_nrm && !pu.normals_deferred) {
    LC3D_LAUNCH(ctx, unpack_strided, div_up(n, 256), 256, 0, pu.raw_nrm, pu.stride_nrm, (int)n, 0.0f,
                (const unsigned char*)nullptr, (int64_t)0, pu.d->normal.as<float4>());
    pu.d->has_normal = true;
  }
}

void upload_cloud(lc3d_ctx* ctx, const lc3d_cloud* h, lc3d_dcloud* d, bool want_normals) {
  upload_finish(ctx, upload_begin(ctx, h, d, want_normals, ctx->scratch[kScrRawA], ctx->scratch[kScrRawB],
                                  nullptr, nullptr));
}

float gate_from_distance(double max_dist) {
  if (!(max_dist > 0) || std::isinf(max_dist)) return INFINITY;
  double d2 = max_dist * max_dist;
  if (d2 >= (double)FLT_MAX) return INFINITY;
  float g = (float)d2;
  if ((double)g > d2) g = std::nextafterf(g, 0.0f);
  return g;  // largest float with (double)g <= max_dist^2: "d2 > max^2" rejects exactly as PCL
}

// Cell edge in units of the estimated point spacing.  Measured optimum on B200 (profiles/
// r01_summary.md): ~4 for the 1-NN searches of ICP (fewer, longer rows; smaller index table),
// ~1.5 for the k-NN passes.  Overridable for experiments.
double cell_factor_env() {
  const char* e = std::getenv("LC3D_CELL_FACTOR");
  double f = e ? std::atof(e) : 4.0;
  return f > 0.1 ? f : 4.0;
}
// x subdivision of the 1-NN index cells (power of two): rows stay few, runs clip tightly.
int xsub_env() {
  const char* e = std::getenv("LC3D_XSUB");
  int x = e ? std::atoi(e) : 8;
  return x >= 1 && x <= 16 ? x : 8;
}
// k-NN passes: the cell edge (in point spacings) is chosen from k so that the guaranteed ball of the
// (2K+1)^3 block with K = 2 holds about kKnnFill * k points of a surface sampled at the estimated
// spacing (pi (K cf)^2 >= fill k): 25 cell rows = one lane-parallel step of the search, one sort, and
// a second, larger block only where the cloud is locally sparser.  LC3D_KNN_CELL_FACTOR overrides.
constexpr double kKnnFill = 1.4;
double knn_cell_factor(int k) {
  if (const char* e = std::getenv("LC3D_KNN_CELL_FACTOR")) {
    const double f = std::atof(e);
    if (f > 0.1) return f;
  }
  double fill = kKnnFill;
  if (const char* e = std::getenv("LC3D_KNN_FILL")) fill = std::max(0.5, std::atof(e));
  const double f = std::sqrt(fill * std::max(k, 1) / 3.14159265358979) / 2.0;
  return std::min(std::max(f, 1.0), 6.0);
}
// first block half-width (cells) for k neighbours on an index built with cell factor cf
int knn_kfirst(double cf, int k) {
  double fill = kKnnFill;
  if (const char* e = std::getenv("LC3D_KNN_FILL")) fill = std::max(0.5, std::atof(e));
  return std::max(1, (int)std::ceil(std::sqrt(fill * std::max(k, 1) / 3.14159265358979) / cf - 1e-3));
}

// before_source: called after the target index is built and before the source is first touched
// (the host-buffer path finishes the source upload there, so that copy overlaps the build).
// overlap_download: move the registered cloud to the host on the copy stream while
// getFitnessScore runs.
// Hooks of the host-buffer path, each called stream-ordered right before the data is first read:
// the uploads still in flight are finished as late as possible.
struct IcpHooks {
  std::function<void()> before_source;          // source xyz (first read by the Morton ordering)
  std::function<void()> before_target_normals;  // target normals (first read by the index gather)
  std::function<void()> before_source_normals;  // source normals (first read by the registered-cloud transform)
};
void icp_run(lc3d_ctx* ctx, const lc3d_dcloud* src, const lc3d_dcloud* tgt, const lc3d_icp_params* p,
             lc3d_icp_result* res, const lc3d_icp_outputs* out, const IcpHooks& hooks = IcpHooks(),
             bool overlap_download = false, bool sharded = false, double* fitness_parts = nullptr) {
  const std::function<void()>& before_source = hooks.before_source;
  cudaStream_t st = ctx->stream;
  const int n = (int)src->n;
  if (p->max_iterations <= 0) throw CudaError{"max_iterations needs to be greater than zero."};
  if (p->mode != LC3D_ICP_POINT_TO_POINT && p->mode != LC3D_ICP_POINT_TO_PLANE)
    throw CudaError{"unknown ICP mode"};
  if (p->mode == LC3D_ICP_POINT_TO_PLANE && tgt->n > 0 && !tgt->has_normal)
    throw CudaError{"point-to-plane ICP needs target normals"};
  // ---- spatial index over the target (KdTreeFLANN::setInputCloud) ----
  ctx->tm[1].start(st);
  Grid& G = *ctx->grid;
  // plan (bbox, cell size: one host round trip), then the target fill (keys, sort, scan, gather)
  // runs on the auxiliary stream while this stream orders the source along the Morton curve of
  // the planned cells — two independent chains of short launch-bound kernels
  GridHint hint{};
  if (tgt->has_hint && !std::getenv("LC3D_NO_GRID_HINT")) {
    for (int d = 0; d < 3; ++d) {
      hint.lo[d] = tgt->hint_lo[d];
      hint.hi[d] = tgt->hint_hi[d];
    }
    hint.nfinite = tgt->n;
    hint.spacing = tgt->hint_spacing;
  }
  grid_plan(ctx, G, tgt->xyz.as<float4>(), tgt->n, cell_factor_env(), 0.0, xsub_env(),
            hint.nfinite > 0 ? &hint : nullptr);
  // host-buffer point-to-plane on a slow PCIe link: while the target normals are still in flight,
  // build the index and run the search of iteration 0 without them, gather them into index order
  // afterwards and let icp_estimate_kernel compute iteration 0's sums.  At 43 GB/s the normals land
  // 40 us after the index is built and the extra gather + estimate launches cost 30 us: no gain; at
  // 25 GB/s (the same box type, another host) they land 200 us later (profiles/r02_summary.md).
  // LC3D_DEFER_NORMALS=0/1 forces the choice.
  // Default: decided by the host->device rate the previous host-buffer call on this context measured —
  // below kDeferBelowGbs the normals are the last thing the first iteration would wait for.
  constexpr double kDeferBelowGbs = 35.0;
  bool want_defer = ctx->h2d_gbs > 0.0 && ctx->h2d_gbs < kDeferBelowGbs;
  if (const char* e = std::getenv("LC3D_DEFER_NORMALS")) want_defer = std::atoi(e) != 0;
  const bool defer_normals = hooks.before_target_normals && p->mode == LC3D_ICP_POINT_TO_PLANE && tgt->has_normal &&
                             tgt->n > 0 && n > 0 && !sharded && !std::getenv("LC3D_STATS") && want_defer;
  const bool two_streams = ctx->aux_stream != nullptr && !std::getenv("LC3D_NO_AUX");
  {
    struct StreamSwap {
      lc3d_ctx* c;
      cudaStream_t saved;
      ~StreamSwap() { c->stream = saved; }
    } swap{ctx, ctx->stream};
    if (two_streams) ctx->stream = ctx->aux_stream;
    const float gate0 = gate_from_distance(p->max_correspondence_distance);
    grid_fill(ctx, G, tgt->xyz.as<float4>(),
              tgt->has_normal && !defer_normals ? tgt->normal.as<float4>() : nullptr, tgt->n,
              defer_normals ? std::function<void()>() : hooks.before_target_normals,
              std::isinf(gate0) || std::getenv("LC3D_NO_OCC") ? 0.0 : (double)std::nextafterf(std::sqrt(gate0), INFINITY));
    if (two_streams) LC3D_CUDA(cudaEventRecord(ctx->ev_aux, ctx->aux_stream));
  }
  if (before_source) before_source();
  ctx->scratch[kScrSrcSorted].ensure((size_t)n * 16 + 16);
  ctx->scratch[kScrSrcWork].ensure((size_t)n * 16 + 16);
  float4* X0 = ctx->scratch[kScrSrcSorted].as<float4>();
  float4* X = ctx->scratch[kScrSrcWork].as<float4>();
  ctx->scratch[kScrBound].ensure((size_t)n * 4 + 16);
  float* Bnd = ctx->scratch[kScrBound].as<float>();
  ctx->scratch[kScrPrevMatch].ensure((size_t)n * 4 + 16);
  int* Mj = ctx->scratch[kScrPrevMatch].as<int>();
  sort_queries_by_cell(ctx, G, src->xyz.as<float4>(), n, X0, X, kScrSrcSort0);  // pristine + working copy
  if (two_streams) LC3D_CUDA(cudaStreamWaitEvent(st, ctx->ev_aux, 0));
  ctx->tm[1].stop(st);
  // ---- the loop ----
  ctx->scratch[kScrState].ensure(sizeof(IcpState));
  IcpState* d_state = ctx->scratch[kScrState].as<IcpState>();
  const int nblk_fit = kFitReduceBlocks;
  const int nblk = std::max(1, div_up(n, kIcpThreads));
  const int nwarps_icp = nblk * (kIcpThreads / 32);
  ctx->scratch[kScrPartials].ensure((size_t)32 * std::max(nwarps_icp, nblk_fit) * 8 + 32 * 8);
  double* partials = ctx->scratch[kScrPartials].as<double>();
  double* reduced = partials + (size_t)32 * std::max(nwarps_icp, nblk_fit);
  IcpConfig cfg;
  cfg.gate = gate_from_distance(p->max_correspondence_distance);
  if (std::isinf(cfg.gate)) {
    cfg.gate_ext = INFINITY;
    cfg.gate_dist = INFINITY;
    cfg.margin_slack = 0.0f;
    cfg.slack_floor = INFINITY;  // no skipping without a finite gate
  } else {
    // searches use an extended gate so that rejected queries learn how far beyond the gate
    // their nearest neighbour is (the slack that lets later iterations skip them)
    const double r = std::sqrt((double)cfg.gate);
    // margin: extra reach of the search beyond the gate = what a rejected query can learn as
    // slack.  Small on purpose: every query without a correspondence pays for the whole extended
    // ball whenever its slack runs out (measured: 2 cells -> 0.15 r cut the loop by 14 %)
    const char* me = std::getenv("LC3D_GATE_MARGIN");  // in cell edges
    const double mc = me ? std::atof(me) : 0.25;
    const char* mre = std::getenv("LC3D_GATE_MARGIN_R");  // as a fraction of the gate distance
    const double mr = mre ? std::atof(mre) : 0.15;
    const double margin = std::max(mr * r, mc * (double)G.v.c);
    cfg.gate_dist = std::nextafterf((float)r, INFINITY);
    float ge = (float)((r + margin) * (r + margin));
    cfg.gate_ext = std::isfinite(ge) ? ge : INFINITY;
    cfg.slack_floor = (float)(1e-4 * (r + margin));
    cfg.margin_slack = std::isfinite(ge) ? (float)(margin * 0.9999) : 0.0f;
    if (!std::isfinite(ge)) cfg.slack_floor = INFINITY;
  }
  cfg.max_iterations = p->max_iterations;
  cfg.rot_thr = 1.0 - p->transformation_epsilon;
  cfg.transl_thr = p->transformation_epsilon;
  cfg.rel_mse = p->euclidean_fitness_epsilon;
  cfg.abs_mse = 1e-12;
  cfg.dump_iteration = p->dump_iteration;
  cfg.mode = p->mode;
  cfg.stats = nullptr;
  const bool want_stats = std::getenv("LC3D_STATS") != nullptr;
  if (want_stats) {
    ctx->scratch[kScrMisc2].ensure(sizeof(SearchStats) * (p->max_iterations + 1));
    cfg.stats = ctx->scratch[kScrMisc2].as<SearchStats>();
    LC3D_CUDA(cudaMemsetAsync(cfg.stats, 0, sizeof(SearchStats) * (p->max_iterations + 1), st));
    for (int it = 0; it <= p->max_iterations; ++it)
      LC3D_CUDA(cudaMemsetAsync(&cfg.stats[it].c[11], 0xff, 8, st));
  }
  unsigned long long* d_block_log = nullptr;
  const char* block_log_path = want_stats ? std::getenv("LC3D_BLOCK_LOG") : nullptr;
  if (block_log_path) {
    LC3D_CUDA(cudaMalloc(&d_block_log, (size_t)p->max_iterations * nblk * 24));
    LC3D_CUDA(cudaMemset(d_block_log, 0, (size_t)p->max_iterations * nblk * 24));
    LC3D_CUDA(cudaMemcpyToSymbol(g_block_log, &d_block_log, sizeof(d_block_log)));
  }
  int32_t* d_dump_idx = nullptr;
  float* d_dump_d2 = nullptr;
  const bool dump = out && (out->corr_index || out->corr_dist2) && p->dump_iteration >= 0 && n > 0;
  if (dump) {
    ctx->scratch[kScrDumpIdx].ensure((size_t)n * 4);
    ctx->scratch[kScrDumpD2].ensure((size_t)n * 4);
    d_dump_idx = ctx->scratch[kScrDumpIdx].as<int32_t>();
    d_dump_d2 = ctx->scratch[kScrDumpD2].as<float>();
    LC3D_CUDA(cudaMemsetAsync(d_dump_idx, 0xff, (size_t)n * 4, st));
    LC3D_CUDA(cudaMemsetAsync(d_dump_d2, 0x7f, (size_t)n * 4, st));
  }
  ctx->tm[2].start(st);
  LC3D_LAUNCH(ctx, icp_state_init, 1, 32, 0, d_state);
  std::vector<cudaEvent_t> iter_ev;
  if (want_stats) {
    iter_ev.resize(p->max_iterations + 1);
    for (auto& e : iter_ev) LC3D_CUDA(cudaEventCreate(&e));
    LC3D_CUDA(cudaEventRecord(iter_ev[0], st));
  }
  // The loop runs on the device (solve + convergence test in the kernel); the host only stops
  // ENQUEUEING: launches go out in chunks of kChunk, after each chunk the `done` flag is
  // copied to pinned memory asynchronously, and the host looks at chunk c's flag only after
  // chunk c+1 is already queued — no bubble, no per-iteration round trip, and far fewer no-op
  // launches than enqueueing all max_iterations up front.
  int chunk_env = 4;
  if (const char* e = std::getenv("LC3D_CHUNK")) chunk_env = std::max(1, std::atoi(e));
  const int kChunk = want_stats ? p->max_iterations : chunk_env;
  ctx->pinned[1].ensure(sizeof(int) * (size_t)(p->max_iterations / kChunk + 2));
  int* h_done = ctx->pinned[1].as<int>();
  cudaEvent_t chunk_ev[2] = {ctx->chunk.a, ctx->chunk.b};
  const bool pdl = !(std::getenv("LC3D_PDL") && std::atoi(std::getenv("LC3D_PDL")) == 0);
  // source-sharded mode: partial rows go to the IPC-exported exchange buffer (double-buffered by
  // iteration parity) and the solve kernel reads every rank's rows over peer memory
  ShardView sv{};
  ShardHeader* shard_hdr = nullptr;
  double* shard_rows = nullptr;
  if (sharded) {
    auto& sh = ctx->shard;
    if (sh.rank < 0 || sh.world < 1) throw CudaError{"lc3d_icp_align_sharded: call lc3d_shard_export / lc3d_shard_connect first"};
    if (nblk > sh.row_stride) throw CudaError{"lc3d_icp_align_sharded: source shard larger than the exported capacity"};
    if (p->max_iterations >= 65535) throw CudaError{"lc3d_icp_align_sharded: max_iterations must be < 65535"};
    shard_hdr = reinterpret_cast<ShardHeader*>(sh.xbuf.p);
    shard_rows = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(sh.xbuf.p) + 256);
    sh.epoch += 1;
    sv.world = sh.world;
    sv.rank = sh.rank;
    sv.base = sh.epoch << 16;
    sv.row_stride = sh.row_stride;
    for (int r = 0; r < sh.world; ++r) {
      unsigned char* base = reinterpret_cast<unsigned char*>(r == sh.rank ? sh.xbuf.p : sh.peer[r]);
      sv.hdr[r] = reinterpret_cast<const ShardHeader*>(base);
      sv.rows[r] = reinterpret_cast<const double*>(base + 256);
    }
  }
  int ring_lo = 1, ring_hi = 2;  // LC3D_RING_ITERS=lo-hi (e.g. "1-2"; "1-0" switches the ring walk off)
  if (const char* e = std::getenv("LC3D_RING_ITERS")) {
    if (std::sscanf(e, "%d-%d", &ring_lo, &ring_hi) != 2) ring_lo = 1, ring_hi = 2;
  }
  auto launch_one = [&](int it) {
    if (want_stats && it > 0) LC3D_CUDA(cudaEventRecord(iter_ev[it], st));
    if (sharded) {
      // this rank's rows of parity q = it & 1 start at q * 32 * row_stride; inside that region they
      // are value-major with the pitch the search kernel writes (its grid size, published as nblk)
      double* rows_q = shard_rows + (size_t)(it & 1) * 32 * sv.row_stride;
      if (p->mode == LC3D_ICP_POINT_TO_PLANE) {
        LC3D_LAUNCH_PDL(ctx, pdl, (icp_iteration_kernel<LC3D_ICP_POINT_TO_PLANE, false>), nblk, kIcpThreads, d_state,
                        cfg, G.v, X, Bnd, Mj, n, rows_q, d_dump_idx, d_dump_d2);
        LC3D_LAUNCH_PDL(ctx, pdl, shard_publish_kernel, 1, 32, d_state, shard_hdr, (unsigned)nblk, sv.base);
        LC3D_LAUNCH_PDL(ctx, pdl, icp_solve_sharded_kernel<LC3D_ICP_POINT_TO_PLANE>, kNvP2Plane, kSolveThreads, d_state,
                        cfg, sv, reduced);
      } else {
        LC3D_LAUNCH_PDL(ctx, pdl, (icp_iteration_kernel<LC3D_ICP_POINT_TO_POINT, false>), nblk, kIcpThreads, d_state,
                        cfg, G.v, X, Bnd, Mj, n, rows_q, d_dump_idx, d_dump_d2);
        LC3D_LAUNCH_PDL(ctx, pdl, shard_publish_kernel, 1, 32, d_state, shard_hdr, (unsigned)nblk, sv.base);
        LC3D_LAUNCH_PDL(ctx, pdl, icp_solve_sharded_kernel<LC3D_ICP_POINT_TO_POINT>, kNvP2P, kSolveThreads, d_state, cfg,
                        sv, reduced);
      }
      return;
    }
    // programmatic dependent launch: each kernel of the chain is staged while its predecessor
    // drains (the kernels call pdl_wait() before reading anything the predecessor wrote)
    if (it == 0 && defer_normals) {
      if (ring_lo <= 0 && ring_hi >= 0)
        LC3D_LAUNCH_PDL(ctx, pdl, (icp_iteration_kernel<LC3D_ICP_POINT_TO_PLANE, false, true, true>), nblk, kIcpThreads,
                        d_state, cfg, G.v, X, Bnd, Mj, n, partials, d_dump_idx, d_dump_d2);
      else
        LC3D_LAUNCH_PDL(ctx, pdl, (icp_iteration_kernel<LC3D_ICP_POINT_TO_PLANE, false, false, true>), nblk, kIcpThreads,
                        d_state, cfg, G.v, X, Bnd, Mj, n, partials, d_dump_idx, d_dump_d2);
      hooks.before_target_normals();
      grid_attach_normals(ctx, G, tgt->normal.as<float4>(), tgt->n);
      LC3D_LAUNCH(ctx, icp_estimate_kernel<LC3D_ICP_POINT_TO_PLANE>, nblk, kIcpThreads, 0, d_state, cfg, G.v, X, Mj, n,
                  partials);
      LC3D_LAUNCH_PDL(ctx, pdl, icp_solve_kernel<LC3D_ICP_POINT_TO_PLANE>, kNvP2Plane, kSolveThreads, d_state, cfg,
                      partials, nblk, reduced);
      return;
    }
    // iterations ring_lo..ring_hi (the ones right after the large first pose updates, when the
    // previous matches are stale seeds and the search balls several cells wide) walk centre-out
    const bool rings = it >= ring_lo && it <= ring_hi;
#define LC3D_ITER_LAUNCH(MODE_, STATS_, RINGS_)                                                              \
  LC3D_LAUNCH_PDL(ctx, pdl, (icp_iteration_kernel<MODE_, STATS_, RINGS_>), nblk, kIcpThreads, d_state, cfg, G.v, \
                  X, Bnd, Mj, n, partials, d_dump_idx, d_dump_d2)
    if (p->mode == LC3D_ICP_POINT_TO_PLANE) {
      if (want_stats) {
        if (rings) LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_PLANE, true, true);
        else LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_PLANE, true, false);
      } else {
        if (rings) LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_PLANE, false, true);
        else LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_PLANE, false, false);
      }
      LC3D_LAUNCH_PDL(ctx, pdl, icp_solve_kernel<LC3D_ICP_POINT_TO_PLANE>, kNvP2Plane, kSolveThreads, d_state, cfg,
                      partials, nblk, reduced);
    } else {
      if (want_stats) {
        if (rings) LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_POINT, true, true);
        else LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_POINT, true, false);
      } else {
        if (rings) LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_POINT, false, true);
        else LC3D_ITER_LAUNCH(LC3D_ICP_POINT_TO_POINT, false, false);
      }
      LC3D_LAUNCH_PDL(ctx, pdl, icp_solve_kernel<LC3D_ICP_POINT_TO_POINT>, kNvP2P, kSolveThreads, d_state, cfg,
                      partials, nblk, reduced);
    }
#undef LC3D_ITER_LAUNCH
  };
  {
    int it = 0, chunk = 0;
    int pending = -1;  // chunk whose flag has been requested but not yet inspected
    while (it < p->max_iterations) {
      const int end = std::min(it + kChunk, p->max_iterations);
      for (; it < end; ++it) launch_one(it);
      h_done[chunk] = 0;
      LC3D_CUDA(cudaMemcpyAsync(&h_done[chunk], &d_state->done, sizeof(int), cudaMemcpyDeviceToHost, st));
      LC3D_CUDA(cudaEventRecord(chunk_ev[chunk & 1], st));
      if (pending >= 0) {
        LC3D_CUDA(cudaEventSynchronize(chunk_ev[pending & 1]));
        if (h_done[pending]) break;
      }
      pending = chunk++;
    }
  }
  if (want_stats) LC3D_CUDA(cudaEventRecord(iter_ev[p->max_iterations], st));
  ctx->tm[2].stop(st);
  // ---- registered cloud (fine_registration.cpp:121 output) ----
  cudaStream_t ds = st;  // stream the registered cloud is downloaded on
  const bool want_reg = out && out->registered_xyz && n > 0;
  if (want_reg) {
    ctx->scratch[kScrOutA].ensure((size_t)n * 12);
    const bool wn = out->registered_normal && src->has_normal;
    if (wn) ctx->scratch[kScrOutB].ensure((size_t)n * 12);
    if (wn && hooks.before_source_normals) hooks.before_source_normals();
    LC3D_LAUNCH(ctx, transform_kernel, div_up(n, 256), 256, 0, d_state->Tfinal,
                src->xyz.as<float4>(), wn ? src->normal.as<float4>() : nullptr, n,
                ctx->scratch[kScrOutA].as<float>(), wn ? ctx->scratch[kScrOutB].as<float>() : nullptr);
    if (overlap_download && ctx->copy_stream) {
      LC3D_CUDA(cudaEventRecord(ctx->ev_copy[2], st));
      LC3D_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy[2], 0));
      ds = ctx->copy_stream;
    }
    LC3D_CUDA(cudaMemcpyAsync(out->registered_xyz, ctx->scratch[kScrOutA].p, (size_t)n * 12,
                              cudaMemcpyDeviceToHost, ds));
    if (wn)
      LC3D_CUDA(cudaMemcpyAsync(out->registered_normal, ctx->scratch[kScrOutB].p, (size_t)n * 12,
                                cudaMemcpyDeviceToHost, ds));
  }
  // ---- getFitnessScore ----
  ctx->tm[3].start(st);
  if (p->compute_fitness) {
    // Bnd is dead after the loop: it becomes the per-point squared distance array
    float* d2_all = Bnd;
    ctx->scratch[kScrMisc3].ensure((size_t)(2 * (size_t)n + 4) * 4);
    FitQueue fq;
    fq.idx = ctx->scratch[kScrMisc3].as<int>();
    fq.seed = fq.idx + n;
    fq.count = reinterpret_cast<unsigned*>(fq.seed + n);
    LC3D_CUDA(cudaMemsetAsync(fq.count, 0, 4, st));
    LC3D_LAUNCH(ctx, icp_fitness_kernel, std::max(1, div_up(n, kIcpThreads)), kIcpThreads, 0, d_state, G.v, X0,
                Mj, 1, n, d2_all, fq,
                cfg.stats ? cfg.stats + p->max_iterations : (SearchStats*)nullptr);
    LC3D_LAUNCH(ctx, icp_fitness_hard_kernel, ctx->num_sms * (1024 / kHardThreads), kHardThreads, 0, d_state, G.v, X0,
                d2_all, fq);
    LC3D_LAUNCH(ctx, icp_fitness_reduce_kernel, kFitReduceBlocks, 256, 0, d_state, d2_all, n, partials,
                fq.count);
  }
  ctx->tm[3].stop(st);
  // ---- results ----
  ctx->tm[4].start(st);
  ctx->pinned[0].ensure(sizeof(IcpState));
  IcpState* h_state = ctx->pinned[0].as<IcpState>();
  LC3D_CUDA(cudaMemcpyAsync(h_state, d_state, sizeof(IcpState), cudaMemcpyDeviceToHost, st));
  if (dump) {
    if (out->corr_index)
      LC3D_CUDA(cudaMemcpyAsync(out->corr_index, d_dump_idx, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (out->corr_dist2)
      LC3D_CUDA(cudaMemcpyAsync(out->corr_dist2, d_dump_d2, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  }
  ctx->tm[4].stop(st);
  ctx->tm[5].stop(st);
  LC3D_CUDA(cudaStreamSynchronize(st));
  if (ds != st) LC3D_CUDA(cudaStreamSynchronize(ds));
  if (d_block_log) {
    std::vector<unsigned long long> hb((size_t)p->max_iterations * nblk * 3);
    LC3D_CUDA(cudaMemcpy(hb.data(), d_block_log, hb.size() * 8, cudaMemcpyDeviceToHost));
    if (FILE* f = std::fopen(block_log_path, "wb")) {
      const int hdr[2] = {p->max_iterations, nblk};
      std::fwrite(hdr, 4, 2, f);
      std::fwrite(hb.data(), 8, hb.size(), f);
      std::fclose(f);
    }
    unsigned long long* null_log = nullptr;
    LC3D_CUDA(cudaMemcpyToSymbol(g_block_log, &null_log, sizeof(null_log)));
    cudaFree(d_block_log);
  }
  if (want_stats) {
    std::vector<SearchStats> hs(p->max_iterations + 1);
    LC3D_CUDA(cudaMemcpy(hs.data(), cfg.stats, sizeof(SearchStats) * (p->max_iterations + 1), cudaMemcpyDeviceToHost));
    {
      const SearchStats& f = hs[p->max_iterations];
      std::fprintf(stderr,
                   "[lc3d stats] fitness searched %llu prev-seed %llu probe %llu borrowed %llu walked %llu "
                   "fallback %llu | warp cand-iters %llu rows %llu\n",
                   f.c[0], f.c[1], f.c[2], f.c[3], f.c[4], f.c[5], f.c[6], f.c[7]);
    }
    std::fprintf(stderr, "[lc3d stats] cell %.5f xsub %d dims %dx%dx%d n_src %d\n", G.v.c, G.v.xs, G.v.dx, G.v.dy, G.v.dz, n);
    for (int it = 0; it < h_state->iter + 1 && it < p->max_iterations; ++it) {
      float ms = 0;
      cudaEventElapsedTime(&ms, iter_ev[it], iter_ev[it + 1]);
      std::fprintf(stderr, "[lc3d stats] it %2d kernel %.1f us\n", it, ms * 1e3f);
    }
    for (auto& e : iter_ev) cudaEventDestroy(e);
    for (int it = 0; it < h_state->iter; ++it)
      std::fprintf(stderr,
                   "[lc3d stats] it %2d searched %7llu prev-seed %7llu probe %7llu borrowed %6llu walked %7llu "
                   "fallback %6llu | warp cand-iters %9llu rows %8llu | ns: search %llu, final reduce %llu, solve %llu\n",
                   it, hs[it].c[0], hs[it].c[1], hs[it].c[2], hs[it].c[3], hs[it].c[4], hs[it].c[5],
                   hs[it].c[6], hs[it].c[7], hs[it].c[8], hs[it].c[9], hs[it].c[10]);
  }
  std::memcpy(res->transformation, h_state->Tfinal, sizeof res->transformation);
  res->converged = h_state->converged;
  res->iterations = h_state->iter;
  res->state = h_state->state;
  res->last_mse = h_state->last_mse;
  res->last_correspondences = h_state->last_corr;
  res->fitness = 0.0;
  if (p->compute_fitness)
    res->fitness = h_state->fitness_cnt > 0 ? h_state->fitness_sum / (double)h_state->fitness_cnt
                                            : std::numeric_limits<double>::max();
  if (fitness_parts) {
    fitness_parts[0] = p->compute_fitness ? h_state->fitness_sum : 0.0;
    fitness_parts[1] = p->compute_fitness ? (double)h_state->fitness_cnt : 0.0;
  }
  if (sharded && h_state->pad[0]) throw CudaError{"lc3d_icp_align_sharded: timed out waiting for a peer's partial sums"};
  res->ms_index = ctx->tm[1].ms();
  res->ms_loop = ctx->tm[2].ms();
  res->ms_fitness = ctx->tm[3].ms();
  res->ms_download = ctx->tm[4].ms();
  res->ms_total = ctx->tm[5].ms();
}

template <typename F>
int guarded(lc3d_ctx* ctx, F&& f) {
  if (!ctx) return LC3D_ERR_INVALID;
  try {
    ctx->err.clear();
    Guard g(ctx);
    f();
    return LC3D_OK;
  } catch (const CudaError& e) {
    ctx->err = e.msg;
    cudaGetLastError();
    return e.msg.find("failed at") != std::string::npos ? LC3D_ERR_CUDA : LC3D_ERR_INVALID;
  } catch (const std::exception& e) {
    ctx->err = e.what();
    return LC3D_ERR_INTERNAL;
  }
}

// argument validation failure: the message is what lc3d_last_error returns
int invalid(lc3d_ctx* ctx, const char* msg) {
  if (ctx) ctx->err = msg;
  return LC3D_ERR_INVALID;
}

struct TmpClouds {
  lc3d_dcloud &a, &b;
};
TmpClouds tmp_clouds(lc3d_ctx* ctx) { return TmpClouds{ctx->tmp_a, ctx->tmp_b}; }
Grid& ctx_grid(lc3d_ctx* ctx) { return *ctx->grid; }

}  // namespace

extern "C" {

const char* lc3d_version(void) { return "lc3d-b200 0.1 (sm_100a)"; }

int lc3d_create(int device, void* stream, lc3d_ctx** out) {
  if (!out) return LC3D_ERR_INVALID;
  *out = nullptr;
  lc3d_ctx* ctx = nullptr;
  try {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      throw CudaError{std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                      "); lc3d has no CPU fallback"};
    if (device < 0 || device >= count) throw CudaError{"invalid device ordinal"};
    LC3D_CUDA(cudaSetDevice(device));
    ctx = new lc3d_ctx;
    ctx->device = device;
    cudaDeviceProp prop;
    LC3D_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    if (stream) {
      ctx->stream = reinterpret_cast<cudaStream_t>(stream);
      ctx->own_stream = false;
    } else {
      LC3D_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
      ctx->own_stream = true;
    }
    for (auto& t : ctx->tm) t.init();
    ctx->chunk.init();
    LC3D_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (auto& e : ctx->ev_copy) LC3D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : ctx->ev_up) LC3D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    LC3D_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
    LC3D_CUDA(cudaEventCreateWithFlags(&ctx->ev_aux, cudaEventDisableTiming));
    ctx->grid = new Grid;
    *out = ctx;
    return LC3D_OK;
  } catch (const CudaError& e) {
    g_create_error = e.msg;
    cudaGetLastError();
    delete ctx;
    return LC3D_ERR_CUDA;
  }
}

void lc3d_destroy(lc3d_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& b : ctx->scratch) b.release();
  for (auto& b : ctx->pinned) b.release();
  for (auto& b : ctx->stage) b.release();
  for (auto& e : ctx->stage_done)
    if (e) cudaEventDestroy(e);
  for (auto& e : ctx->ev_h2d)
    if (e) cudaEventDestroy(e);
  delete ctx->host_pool;
  ctx->tmp_a.release();
  ctx->tmp_b.release();
  ctx->pool.release_all();
  lc3d_shard_close(ctx);
  if (ctx->grid) {
    ctx->grid->release();
    delete ctx->grid;
  }
  for (auto& t : ctx->tm) t.destroy();
  ctx->chunk.destroy();
  for (auto& e : ctx->ev_copy)
    if (e) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
  for (auto& e : ctx->ev_up)
    if (e) cudaEventDestroy(e);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* lc3d_last_error(const lc3d_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_create_error.c_str();
}

int64_t lc3d_launch_count(const lc3d_ctx* ctx) { return ctx ? ctx->launches : 0; }

void lc3d_debug_grid_info(const lc3d_ctx* ctx, double out[8]) {
  for (int i = 0; i < 8; ++i) out[i] = 0;
  if (!ctx || !ctx->grid) return;
  const GridDev& g = ctx->grid->v;
  out[0] = g.c;
  out[1] = g.dx;
  out[2] = g.dy;
  out[3] = g.dz;
  out[4] = (double)ctx->grid->ncell;
  out[5] = g.n;
}

int64_t lc3d_debug_alloc_count(void) { return (int64_t)lc3d::devbuf_alloc_counter().load(); }

int lc3d_host_register(void* ptr, uint64_t bytes) {
  if (!ptr || bytes == 0) return LC3D_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return LC3D_ERR_CUDA;
  }
  return LC3D_OK;
}

int lc3d_host_unregister(void* ptr) {
  if (!ptr) return LC3D_ERR_INVALID;
  const cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return LC3D_ERR_CUDA;
  }
  return LC3D_OK;
}

int lc3d_cloud_upload(lc3d_ctx* ctx, const lc3d_cloud* host, lc3d_dcloud** out) {
  if (!out || !host) return LC3D_ERR_INVALID;
  *out = nullptr;
  lc3d_dcloud* d = new lc3d_dcloud;
  int rc = guarded(ctx, [&] {
    if (host->n > 0) {
      d->xyz = ctx->pool.acquire((size_t)host->n * 16);
      if (host->normal) d->normal = ctx->pool.acquire((size_t)host->n * 16);
    }
    upload_cloud(ctx, host, d, true);
    LC3D_CUDA(cudaStreamSynchronize(ctx->stream));
  });
  if (rc != LC3D_OK) {
    d->release();
    delete d;
    return rc;
  }
  *out = d;
  return LC3D_OK;
}

void lc3d_cloud_free(lc3d_ctx* ctx, lc3d_dcloud* dc) {
  if (!dc) return;
  if (ctx) {
    cudaSetDevice(ctx->device);
    // every use of the cloud was stream-ordered on ctx->stream and the entry points return
    // synchronised, so the buffers can be handed out again right away
    dc->park(ctx->pool);
  } else {
    dc->release();
  }
  delete dc;
}

int64_t lc3d_dcloud_size(const lc3d_dcloud* dc) { return dc ? dc->n : 0; }

int lc3d_icp_align(lc3d_ctx* ctx, const lc3d_cloud* source, const lc3d_cloud* target,
                   const lc3d_icp_params* params, lc3d_icp_result* result,
                   const lc3d_icp_outputs* outputs) {
  if (!source || !target || !params || !result) return invalid(ctx, "lc3d_icp_align: NULL argument");
  return guarded(ctx, [&] {
    std::memset(result, 0, sizeof *result);
    TmpClouds tc = tmp_clouds(ctx);
    ctx->tm[5].start(ctx->stream);
    ctx->tm[0].start(ctx->stream);
    const bool need_src_normals = outputs && outputs->registered_normal;
    // All four host->device copies go out on the copy stream right away, in the order the
    // device needs them: target xyz (bbox, cell size), source xyz (Morton ordering, concurrent
    // with the target fill), target normals (first read by the index gather), source normals
    // (first read after the loop).  Each array is unpacked, stream-ordered, right before its
    // first use.
    cudaStream_t cs = ctx->copy_stream;
    ctx->h2d_bytes_call = 0;
    for (auto& e : ctx->ev_h2d)
      if (!e) LC3D_CUDA(cudaEventCreate(&e));
    try {
      LC3D_CUDA(cudaEventRecord(ctx->ev_h2d[0], cs ? cs : ctx->stream));
      PendingUpload pt = upload_begin(ctx, target, &tc.b, params->mode == LC3D_ICP_POINT_TO_PLANE,
                                      ctx->scratch[kScrRawA], ctx->scratch[kScrRawB], cs, ctx->ev_up[0], true);
      PendingUpload ps = upload_begin(ctx, source, &tc.a, need_src_normals, ctx->scratch[kScrRawC],
                                      ctx->scratch[kScrRawD], cs, ctx->ev_up[1], true);
      upload_begin_normals(ctx, pt, ctx->scratch[kScrRawB], cs, ctx->ev_up[2]);
      upload_begin_normals(ctx, ps, ctx->scratch[kScrRawD], cs, ctx->ev_up[3]);
      LC3D_CUDA(cudaEventRecord(ctx->ev_h2d[1], cs ? cs : ctx->stream));
      upload_finish(ctx, pt);
      ctx->tm[0].stop(ctx->stream);
      IcpHooks hooks;
      hooks.before_source = [&] { upload_finish(ctx, ps); };
      hooks.before_target_normals = [&] { upload_finish_normals(ctx, pt); };
      hooks.before_source_normals = [&] { upload_finish_normals(ctx, ps); };
      icp_run(ctx, &tc.a, &tc.b, params, result, outputs, hooks, true);
      // every staged copy has been consumed on the paths above except in odd output
      // combinations (normals requested without xyz): never return with a copy in flight
      if (cs) LC3D_CUDA(cudaStreamSynchronize(cs));
      // the rate the four uploads achieved back to back on the copy stream (only meaningful when it
      // carried nothing else: the separate copy stream, pinned or packed-staged sources)
      float ms_h2d = 0.0f;
      if (cs && ctx->h2d_bytes_call >= ((size_t)4 << 20) &&
          cudaEventElapsedTime(&ms_h2d, ctx->ev_h2d[0], ctx->ev_h2d[1]) == cudaSuccess && ms_h2d > 0.0f)
        ctx->h2d_gbs = (double)ctx->h2d_bytes_call / ((double)ms_h2d * 1e6);
    } catch (...) {
      if (cs) cudaStreamSynchronize(cs);  // no copy may still read the caller's buffers
      throw;
    }
    result->ms_upload = ctx->tm[0].ms();
  });
}

int lc3d_icp_align_resident(lc3d_ctx* ctx, const lc3d_dcloud* source, const lc3d_dcloud* target,
                            const lc3d_icp_params* params, lc3d_icp_result* result,
                            const lc3d_icp_outputs* outputs) {
  if (!source || !target || !params || !result) return invalid(ctx, "lc3d_icp_align_resident: NULL argument");
  return guarded(ctx, [&] {
    std::memset(result, 0, sizeof *result);
    ctx->tm[5].start(ctx->stream);
    icp_run(ctx, source, target, params, result, outputs);
  });
}

int lc3d_shard_export(lc3d_ctx* ctx, int64_t max_shard_points, unsigned char handle_out[64]) {
  if (!handle_out || max_shard_points <= 0) return invalid(ctx, "lc3d_shard_export: bad argument");
  return guarded(ctx, [&] {
    auto& sh = ctx->shard;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    sh.row_stride = div_up(max_shard_points, kIcpThreads) + 1;
    const size_t bytes = 256 + (size_t)2 * 32 * sh.row_stride * 8;
    if (sh.xbuf.p) sh.xbuf.release();
    // a dedicated allocation: IPC exports whole cudaMalloc allocations
    LC3D_CUDA(cudaMalloc(&sh.xbuf.p, bytes));
    sh.xbuf.cap = bytes;
    LC3D_CUDA(cudaMemset(sh.xbuf.p, 0, bytes));
    cudaIpcMemHandle_t h;
    LC3D_CUDA(cudaIpcGetMemHandle(&h, sh.xbuf.p));
    std::memcpy(handle_out, &h, 64);
    sh.epoch = 0;
  });
}

int lc3d_shard_connect(lc3d_ctx* ctx, int32_t rank, int32_t world, const unsigned char* handles) {
  if (!handles || world < 1 || world > kShardMaxWorld || rank < 0 || rank >= world)
    return invalid(ctx, "lc3d_shard_connect: bad rank / world (at most 8 ranks) or NULL handles");
  return guarded(ctx, [&] {
    auto& sh = ctx->shard;
    if (!sh.xbuf.p) throw CudaError{"lc3d_shard_connect: call lc3d_shard_export first"};
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      cudaIpcMemHandle_t h;
      std::memcpy(&h, handles + (size_t)r * 64, 64);
      LC3D_CUDA(cudaIpcOpenMemHandle(&sh.peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    sh.rank = rank;
    sh.world = world;
  });
}

void lc3d_shard_close(lc3d_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  auto& sh = ctx->shard;
  for (int r = 0; r < kShardMaxWorld; ++r)
    if (sh.peer[r]) {
      cudaIpcCloseMemHandle(sh.peer[r]);
      sh.peer[r] = nullptr;
    }
  sh.xbuf.release();
  sh.rank = -1;
  sh.world = 0;
}

int lc3d_icp_align_sharded(lc3d_ctx* ctx, const lc3d_dcloud* source_shard, const lc3d_dcloud* target,
                           const lc3d_icp_params* params, lc3d_icp_result* result, const lc3d_icp_outputs* outputs,
                           double fitness_sum_count[2]) {
  if (!source_shard || !target || !params || !result) return invalid(ctx, "lc3d_icp_align_sharded: NULL argument");
  return guarded(ctx, [&] {
    std::memset(result, 0, sizeof *result);
    ctx->tm[5].start(ctx->stream);
    icp_run(ctx, source_shard, target, params, result, outputs, IcpHooks(), false, true, fitness_sum_count);
  });
}

int lc3d_nn(lc3d_ctx* ctx, const lc3d_cloud* cloud, const lc3d_cloud* queries, double max_dist,
            int32_t* out_index, float* out_dist2) {
  if (!cloud || !out_index || !out_dist2) return invalid(ctx, "lc3d_nn: NULL argument");
  return guarded(ctx, [&] {
    TmpClouds tc = tmp_clouds(ctx);
    upload_cloud(ctx, cloud, &tc.b, false);
    const lc3d_dcloud* q = &tc.b;
    if (queries) {
      upload_cloud(ctx, queries, &tc.a, false);
      q = &tc.a;
    }
    const int n = (int)q->n;
    if (n == 0) return;
    Grid& G = ctx_grid(ctx);
    grid_build(ctx, G, tc.b.xyz.as<float4>(), nullptr, tc.b.n, cell_factor_env(), 0.0, xsub_env());
    ctx->scratch[kScrSrcSorted].ensure((size_t)n * 16 + 16);
    float4* Q = ctx->scratch[kScrSrcSorted].as<float4>();
    sort_queries_by_cell(ctx, G, q->xyz.as<float4>(), n, Q);
    ctx->scratch[kScrDumpIdx].ensure((size_t)n * 4);
    ctx->scratch[kScrDumpD2].ensure((size_t)n * 4);
    LC3D_LAUNCH(ctx, nn_kernel, div_up(n, 256), 256, 0, G.v, Q, n, gate_from_distance(max_dist),
                ctx->scratch[kScrDumpIdx].as<int32_t>(), ctx->scratch[kScrDumpD2].as<float>());
    LC3D_CUDA(cudaMemcpyAsync(out_index, ctx->scratch[kScrDumpIdx].p, (size_t)n * 4,
                              cudaMemcpyDeviceToHost, ctx->stream));
    LC3D_CUDA(cudaMemcpyAsync(out_dist2, ctx->scratch[kScrDumpD2].p, (size_t)n * 4,
                              cudaMemcpyDeviceToHost, ctx->stream));
    LC3D_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

int lc3d_transform(lc3d_ctx* ctx, const lc3d_cloud* cloud, const float matrix[16], float* out_xyz,
                   float* out_normal) {
  if (!cloud || !matrix || !out_xyz) return invalid(ctx, "lc3d_transform: NULL argument");
  return guarded(ctx, [&] {
    TmpClouds tc = tmp_clouds(ctx);
    upload_cloud(ctx, cloud, &tc.a, out_normal != nullptr);
    const int n = (int)tc.a.n;
    if (n == 0) return;
    ctx->scratch[kScrMisc].ensure(64);
    LC3D_CUDA(cudaMemcpyAsync(ctx->scratch[kScrMisc].p, matrix, 64, cudaMemcpyHostToDevice, ctx->stream));
    const bool wn = out_normal && tc.a.has_normal;
    ctx->scratch[kScrOutA].ensure((size_t)n * 12);
    if (wn) ctx->scratch[kScrOutB].ensure((size_t)n * 12);
    LC3D_LAUNCH(ctx, transform_kernel, div_up(n, 256), 256, 0, ctx->scratch[kScrMisc].as<float>(),
                tc.a.xyz.as<float4>(), wn ? tc.a.normal.as<float4>() : nullptr, n,
                ctx->scratch[kScrOutA].as<float>(), wn ? ctx->scratch[kScrOutB].as<float>() : nullptr);
    LC3D_CUDA(cudaMemcpyAsync(out_xyz, ctx->scratch[kScrOutA].p, (size_t)n * 12, cudaMemcpyDeviceToHost,
                              ctx->stream));
    if (wn)
      LC3D_CUDA(cudaMemcpyAsync(out_normal, ctx->scratch[kScrOutB].p, (size_t)n * 12,
                                cudaMemcpyDeviceToHost, ctx->stream));
    LC3D_CUDA(cudaStreamSynchronize(ctx->stream));
  });
}

}  // extern "C"

#include "capi_filters.inc"
