// common.cuh — context, device buffers, error handling, exact-float helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lc3d.h"

namespace lc3d {

constexpr int kNumSmsB200 = 148;

struct CudaError {
  std::string msg;
};

#define LC3D_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      throw ::lc3d::CudaError{std::string(#expr) + " failed at " + __FILE__ + ":" +           \
                              std::to_string(__LINE__) + ": " + cudaGetErrorString(_e)};      \
    }                                                                                         \
  } while (0)

// Grow-only device buffer; reused across calls so the steady state allocates nothing.
// device allocations made so far by this process (lc3d_debug_alloc_count: the steady state of a view
// chain is supposed to allocate nothing — cudaFree synchronises the whole device)
inline std::atomic<long long>& devbuf_alloc_counter() {
  static std::atomic<long long> c{0};
  return c;
}
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    // a buffer that has to grow again grows by half: the sizes that depend on the data (cell tables,
    // per-view point counts) settle after a few calls instead of creeping up 12 % at a time
    size_t want = p ? bytes + bytes / 2 + 256 : bytes + bytes / 8 + 256;
    if (p) LC3D_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;
    LC3D_CUDA(cudaMalloc(&p, want));
    cap = want;
    devbuf_alloc_counter().fetch_add(1);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  void ensure(size_t bytes) {
    if (bytes <= cap) return;
    if (p) LC3D_CUDA(cudaFreeHost(p));
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    LC3D_CUDA(cudaMallocHost(&p, want));
    cap = want;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// Fork-join pool of host threads (packing of pageable host clouds into pinned staging memory).
class HostPool {
 public:
  explicit HostPool(int nthreads) {
    for (int t = 0; t < nthreads; ++t) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lock(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int size() const { return (int)workers_.size(); }
  // runs job(0..njobs-1) on the workers and the calling thread; returns when all are done
  void run(int njobs, const std::function<void(int)>& job) {
    {
      std::lock_guard<std::mutex> lock(mu_);
      job_ = &job;
      next_ = 0;
      njobs_ = njobs;
      pending_ = njobs;
    }
    cv_.notify_all();
    for (;;) {  // the caller works too
      int j;
      {
        std::lock_guard<std::mutex> lock(mu_);
        if (next_ >= njobs_) break;
        j = next_++;
      }
      job(j);
      std::lock_guard<std::mutex> lock(mu_);
      --pending_;
    }
    std::unique_lock<std::mutex> lock(mu_);
    done_.wait(lock, [this] { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop() {
    for (;;) {
      int j;
      const std::function<void(int)>* job;
      {
        std::unique_lock<std::mutex> lock(mu_);
        cv_.wait(lock, [this] { return stop_ || (job_ && next_ < njobs_); });
        if (stop_) return;
        j = next_++;
        job = job_;
      }
      (*job)(j);
      std::lock_guard<std::mutex> lock(mu_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  int next_ = 0, njobs_ = 0, pending_ = 0;
  bool stop_ = false;
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  void init() {
    LC3D_CUDA(cudaEventCreate(&a));
    LC3D_CUDA(cudaEventCreate(&b));
  }
  void destroy() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
    a = b = nullptr;
  }
  void start(cudaStream_t s) { LC3D_CUDA(cudaEventRecord(a, s)); }
  void stop(cudaStream_t s) { LC3D_CUDA(cudaEventRecord(b, s)); }
  float ms() {
    float v = 0;
    LC3D_CUDA(cudaEventSynchronize(b));
    LC3D_CUDA(cudaEventElapsedTime(&v, a, b));
    return v;
  }
};

// ---- exact float32 helpers: every parity-critical expression is written with the
// round-to-nearest intrinsics so that nvcc never contracts it into an FMA
// (PCL/FLANN distro builds evaluate mul and add separately).
__device__ __forceinline__ float dist2_exact(float qx, float qy, float qz, float px, float py,
                                             float pz) {
  float d = __fsub_rn(qx, px);
  float r = __fmul_rn(d, d);
  d = __fsub_rn(qy, py);
  r = __fadd_rn(r, __fmul_rn(d, d));
  d = __fsub_rn(qz, pz);
  r = __fadd_rn(r, __fmul_rn(d, d));
  return r;
}

// row r of T times (x,y,z,1): ((T0*x + T1*y) + T2*z) + T3, float32, no FMA.
__device__ __forceinline__ float xform_row(const float* __restrict__ T, int r, float x, float y,
                                           float z) {
  float s = __fmul_rn(T[r * 4 + 0], x);
  s = __fadd_rn(s, __fmul_rn(T[r * 4 + 1], y));
  s = __fadd_rn(s, __fmul_rn(T[r * 4 + 2], z));
  return __fadd_rn(s, T[r * 4 + 3]);
}
__device__ __forceinline__ float rot_row(const float* __restrict__ T, int r, float x, float y,
                                         float z) {
  float s = __fmul_rn(T[r * 4 + 0], x);
  s = __fadd_rn(s, __fmul_rn(T[r * 4 + 1], y));
  return __fadd_rn(s, __fmul_rn(T[r * 4 + 2], z));
}
__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace lc3d

namespace lc3d {
struct Grid;
}

namespace lc3d {
// Free list of device buffers of resident clouds.  cudaMalloc / cudaFree cost 0.1-5 ms each (and
// cudaFree synchronises the device; both get slower as the process maps more memory), which would
// dominate a view chain that uploads and frees a cloud per view: freed cloud buffers are parked
// here and handed out again, so the steady state of a chain allocates nothing.
struct BufPool {
  std::vector<DevBuf> free_list;
  std::mutex mu;  // a resident cloud may be freed from another host thread than the one that made it
  static constexpr size_t kMaxParked = 32;
  // a parked buffer of at least `bytes` (and not absurdly larger), or a fresh allocation
  DevBuf acquire(size_t bytes) {
    std::unique_lock<std::mutex> lock(mu);
    int best = -1;
    for (int i = 0; i < (int)free_list.size(); ++i)
      if (free_list[i].cap >= bytes && free_list[i].cap <= 4 * bytes + (1u << 20) &&
          (best < 0 || free_list[i].cap < free_list[best].cap))
        best = i;
    DevBuf b;
    if (best >= 0) {
      b = free_list[best];
      free_list.erase(free_list.begin() + best);
    } else {
      lock.unlock();
      b.ensure(bytes);
    }
    return b;
  }
  void park(DevBuf& b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> lock(mu);
    if (free_list.size() >= kMaxParked) {  // drop the smallest parked buffer
      int small = 0;
      for (int i = 1; i < (int)free_list.size(); ++i)
        if (free_list[i].cap < free_list[small].cap) small = i;
      free_list[small].release();
      free_list.erase(free_list.begin() + small);
    }
    free_list.push_back(b);
    b.p = nullptr;
    b.cap = 0;
  }
  void release_all() {
    std::lock_guard<std::mutex> lock(mu);
    for (auto& b : free_list) b.release();
    free_list.clear();
  }
};
}  // namespace lc3d

struct lc3d_dcloud {
  int64_t n = 0;
  lc3d::DevBuf xyz;     // float4 (x,y,z,1) in input order
  lc3d::DevBuf normal;  // float4 (nx,ny,nz,curvature) in input order, or empty
  bool has_normal = false;
  // what lc3d_prepare_view knows about the cloud it produced (all points finite, inside this box,
  // about `spacing` apart): lets the ICP index be planned without kernels or a host round trip
  bool has_hint = false;
  float hint_lo[3] = {0, 0, 0}, hint_hi[3] = {0, 0, 0};
  double hint_spacing = 0.0;
  void release() {
    xyz.release();
    normal.release();
    n = 0;
    has_normal = false;
    has_hint = false;
  }
  // buffers go back to the context's pool instead of cudaFree
  void park(lc3d::BufPool& pool) {
    pool.park(xyz);
    pool.park(normal);
    n = 0;
    has_normal = false;
    has_hint = false;
  }
};

// The opaque context of include/lc3d.h.
struct lc3d_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  int64_t launches = 0;
  int num_sms = lc3d::kNumSmsB200;
  lc3d::DevBuf scratch[64];  // grow-only scratch arena, slots named by the users
  lc3d::PinnedBuf pinned[2];
  lc3d::Timer tm[6];
  lc3d::Timer chunk;  // two events used to poll the ICP loop's done flag
  cudaStream_t copy_stream = nullptr;  // H2D / D2H of the host-buffer entry points (overlaps compute)
  cudaEvent_t ev_copy[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_up[4] = {nullptr, nullptr, nullptr, nullptr};  // per-array upload completion (host-buffer ICP)
  cudaStream_t aux_stream = nullptr;  // target index fill, concurrent with the source ordering
  cudaEvent_t ev_aux = nullptr;
  lc3d::Grid* grid = nullptr;  // spatial index reused across calls
  lc3d_dcloud tmp_a, tmp_b;    // staging clouds of the host-buffer entry points
  lc3d::BufPool pool;          // parked buffers of freed resident clouds
  // pageable host clouds: packed by host threads into pinned staging memory, chunk by chunk, each
  // chunk's DMA overlapping the packing of the next (slots: target xyz / normals, source xyz / normals)
  // host->device rate seen by the previous host-buffer alignment (GB/s, 0 = unknown) and this
  // call's staged bytes: PCIe differs 2x between otherwise identical boxes, and the alignment
  // schedules itself around the uploads (icp_run: iteration 0 before the target normals arrive)
  double h2d_gbs = 0.0;
  size_t h2d_bytes_call = 0;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
  lc3d::PinnedBuf stage[4];
  cudaEvent_t stage_done[4] = {nullptr, nullptr, nullptr, nullptr};  // last DMA out of the slot
  lc3d::HostPool* host_pool = nullptr;
  // source-sharded single-pair mode (lc3d_shard_*): exchange buffer + the peers' mappings
  struct Shard {
    int rank = -1, world = 0;
    lc3d::DevBuf xbuf;          // [header 256 B][2 x 32 x row_stride doubles], exported with CUDA IPC
    long long row_stride = 0;
    void* peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // mapped peer buffers
    unsigned epoch = 0;         // alignments run so far (the same on every rank: the calls are collective)
  } shard;
};

namespace lc3d {
// Programmatic dependent launch (sm_90+): a kernel launched with the attribute may be scheduled
// before its predecessor in the stream has drained; it must call pdl_wait() before touching
// anything the predecessor wrote.  pdl_trigger() lets the successor start launching early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(cudaStream_t stream, bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
}  // namespace lc3d
#define LC3D_LAUNCH_PDL(ctx, pdl, kernel, grid, block, ...)                                  \
  do {                                                                                        \
    LC3D_CUDA(lc3d::launch_pdl((ctx)->stream, (pdl), kernel, dim3(grid), dim3(block), __VA_ARGS__)); \
    ++(ctx)->launches;                                                                        \
  } while (0)

#define LC3D_LAUNCH(ctx, kernel, grid, block, smem, ...)                      \
  do {                                                                        \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);          \
    ++(ctx)->launches;                                                        \
    LC3D_CUDA(cudaGetLastError());                                            \
  } while (0)
