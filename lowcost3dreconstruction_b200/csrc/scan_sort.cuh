// scan_sort.cuh — hand-written exclusive scan and stable LSD radix sort (key32 + value32).
//
// These build the spatial index (cell-id keys) and the VoxelGrid ordering.  No CUB /
// Thrust: the north star asks for a hand-written sort.  Both are HBM/L2 streaming
// kernels: vectorised coalesced loads, shared-memory histograms, warp match ranking.
#pragma once
#include "common.cuh"

namespace lc3d {

// ------------------------------------------------------------------ scan -----
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  // exclusive scan of one value per thread across the block (kScanThreads threads)
  __shared__ uint32_t warp_tot[kScanThreads / 32];
  __shared__ uint32_t block_tot;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t t = lane < kScanThreads / 32 ? warp_tot[lane] : 0u;
    uint32_t ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    if (lane < kScanThreads / 32) warp_tot[lane] = ti - t;
    if (lane == 31) block_tot = ti;
  }
  __syncthreads();
  uint32_t r = inc - v + warp_tot[w];
  *total = block_tot;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t* __restrict__ in,
                                                               int64_t n,
                                                               uint32_t* __restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t s = 0;
  if (base + kScanItems <= n) {
    const uint4* p = reinterpret_cast<const uint4*>(in + base);
    uint4 a = p[0], b = p[1];
    s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
  } else {
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) s += in[base + k];
  }
  uint32_t tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads) scan_block_sums(uint32_t* sums, int nb) {
  // single block: exclusive scan of the tile sums, chunked with a running carry
  uint32_t carry = 0;
  for (int base = 0; base < nb; base += kScanThreads) {
    int i = base + threadIdx.x;
    uint32_t v = i < nb ? sums[i] : 0u;
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(v, &tot);
    if (i < nb) sums[i] = ex + carry;
    carry += tot;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t* __restrict__ in,
                                                           uint32_t* __restrict__ out, int64_t n,
                                                           const uint32_t* __restrict__ sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  const bool full = base + kScanItems <= n;
  if (full) {
    const uint4* p = reinterpret_cast<const uint4*>(in + base);
    uint4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = base + k < n ? in[base + k] : 0u;
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) s += v[k];
  uint32_t tot;
  uint32_t ex = block_exclusive_scan(s, &tot) + sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    uint32_t t = v[k];
    v[k] = ex;
    ex += t;
  }
  if (full) {
    uint4* p = reinterpret_cast<uint4*>(out + base);
    p[0] = make_uint4(v[0], v[1], v[2], v[3]);
    p[1] = make_uint4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) out[base + k] = v[k];
  }
}

// Single-block exclusive scan for small inputs (the radix histograms: 256 x #tiles entries):
// one launch instead of three.  Each thread owns a contiguous chunk.
constexpr int kSmallScanThreads = 1024;
constexpr int64_t kSmallScanMax = 2048;  // beyond this the 3-kernel tiled scan is faster
__global__ void __launch_bounds__(kSmallScanThreads)
    scan_small_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n) {
  __shared__ uint32_t wtot[kSmallScanThreads / 32];
  const int per = (n + kSmallScanThreads - 1) / kSmallScanThreads;
  const int lo = min(threadIdx.x * per, n), hi = min(lo + per, n);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += in[i];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t t = wtot[lane], ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    wtot[lane] = ti - t;
  }
  __syncthreads();
  uint32_t run = inc - s + wtot[w];
  for (int i = lo; i < hi; ++i) {
    const uint32_t v = in[i];
    out[i] = run;
    run += v;
  }
}

// One-block exclusive scan for mid-sized inputs (the radix histograms of a pass: 256 x #tiles,
// tens of thousands of entries): ONE launch instead of three launch-bound ones.  Each thread
// owns a contiguous chunk of up to 40 entries and issues all its loads up front (one memory
// round trip), then one block scan of the thread totals.
constexpr int kOneBlockVec = 10;                                  // uint4 per thread
constexpr int64_t kOneBlockScanMax = 1024 * 4 * kOneBlockVec;     // 40960
__global__ void __launch_bounds__(1024)
    scan_one_block_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n) {
  __shared__ uint32_t wtot[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // chunk length: multiple of 4 so that every chunk starts 16-byte aligned
  const int per = (((n + 1023) >> 10) + 3) & ~3;
  const int lo = threadIdx.x * per;
  uint4 v[kOneBlockVec];
#pragma unroll
  for (int k = 0; k < kOneBlockVec; ++k) {
    const int idx = lo + 4 * k;
    v[k] = make_uint4(0u, 0u, 0u, 0u);
    if (4 * k < per) {
      if (idx + 4 <= n) {
        v[k] = *reinterpret_cast<const uint4*>(in + idx);
      } else {
        if (idx < n) v[k].x = in[idx];
        if (idx + 1 < n) v[k].y = in[idx + 1];
        if (idx + 2 < n) v[k].z = in[idx + 2];
      }
    }
  }
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kOneBlockVec; ++k) sum += v[k].x + v[k].y + v[k].z + v[k].w;
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wtot[w] = inc;
  __syncthreads();
  if (w == 0) {
    const uint32_t t = wtot[lane];
    uint32_t ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += u;
    }
    wtot[lane] = ti - t;
  }
  __syncthreads();
  uint32_t ex = wtot[w] + inc - sum;
#pragma unroll
  for (int k = 0; k < kOneBlockVec; ++k) {
    const int idx = lo + 4 * k;
    if (4 * k < per) {
      uint4 q;
      q.x = ex; ex += v[k].x;
      q.y = ex; ex += v[k].y;
      q.z = ex; ex += v[k].z;
      q.w = ex; ex += v[k].w;
      if (idx + 4 <= n) {
        *reinterpret_cast<uint4*>(out + idx) = q;
      } else {
        if (idx < n) out[idx] = q.x;
        if (idx + 1 < n) out[idx + 1] = q.y;
        if (idx + 2 < n) out[idx + 2] = q.z;
      }
    }
  }
}

// Single-pass exclusive scan with decoupled look-back (Merrill & Garland): every tile publishes
// its aggregate, then its inclusive prefix, in a 64-bit status word (flag << 32 | value); a tile
// obtains its exclusive prefix by walking back over its predecessors' status words, 32 at a time
// with one warp, until it meets a published prefix.  Tiles are claimed in order from a global
// counter, so a tile only ever waits for tiles that are already running.  One launch and
// 2 x 4 B of traffic per element instead of three launches and 3 x 4 B.
__global__ void __launch_bounds__(kScanThreads)
    scan_lookback_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
                         unsigned long long* __restrict__ status, unsigned* __restrict__ counter) {
  const unsigned long long kScanFlagAgg = 1ull << 32, kScanFlagPrefix = 2ull << 32;
  __shared__ unsigned s_tile;
  __shared__ uint32_t s_prefix;
  if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1u);
  __syncthreads();
  const unsigned tile = s_tile;
  const int64_t base = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  const bool full = base + kScanItems <= n;
  if (full) {
    const uint4* p = reinterpret_cast<const uint4*>(in + base);
    uint4 a = p[0], b = p[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = base + k < n ? in[base + k] : 0u;
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) s += v[k];
  uint32_t agg;
  uint32_t ex = block_exclusive_scan(s, &agg);
  if (threadIdx.x < 32) {
    const unsigned fullm = 0xffffffffu;
    const int lane = threadIdx.x;
    volatile unsigned long long* st = status;
    if (lane == 0) {
      st[tile] = (tile == 0 ? kScanFlagPrefix : kScanFlagAgg) | agg;
      __threadfence();
    }
    uint32_t exclusive = 0;
    if (tile > 0) {
      int64_t t = (int64_t)tile - 1 - lane;
      for (;;) {
        unsigned long long w = kScanFlagPrefix;  // before the first tile: prefix 0
        if (t >= 0) w = st[t];
        while (__any_sync(fullm, (w >> 32) == 0ull)) {
          if ((w >> 32) == 0ull) w = st[t];
        }
        const unsigned pm = __ballot_sync(fullm, (w >> 32) == 2ull);
        const int first = pm ? __ffs(pm) - 1 : 32;
        uint32_t val = lane <= first ? (uint32_t)w : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(fullm, val, o);
        exclusive += val;
        if (pm) break;
        t -= 32;
      }
      if (lane == 0) {
        st[tile] = kScanFlagPrefix | (unsigned long long)(uint32_t)(exclusive + agg);
        __threadfence();
      }
    }
    if (lane == 0) s_prefix = exclusive;
  }
  __syncthreads();
  ex += s_prefix;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    uint32_t t = v[k];
    v[k] = ex;
    ex += t;
  }
  if (full) {
    uint4* p = reinterpret_cast<uint4*>(out + base);
    p[0] = make_uint4(v[0], v[1], v[2], v[3]);
    p[1] = make_uint4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (base + k < n) out[base + k] = v[k];
  }
}

// Exclusive scan of n uint32 (in may alias out).  `sums` scratch: scan_scratch_bytes(n) bytes,
// 8-byte aligned.  Requires in/out 16-byte aligned (cudaMalloc'd).
inline void exclusive_scan_u32(lc3d_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n,
                               uint32_t* sums) {
  if (n <= 0) return;
  if (n <= kSmallScanMax) {
    LC3D_LAUNCH(ctx, scan_small_kernel, 1, kSmallScanThreads, 0, in, out, (int)n);
    return;
  }
  if (n <= kOneBlockScanMax) {
    LC3D_LAUNCH(ctx, scan_one_block_kernel, 1, 1024, 0, in, out, (int)n);
    return;
  }
  const int nb = div_up(n, kScanTile);
  // status words (one per tile) + the tile counter, zeroed before every scan
  LC3D_CUDA(cudaMemsetAsync(sums, 0, (size_t)nb * 8 + 16, ctx->stream));
  unsigned long long* status = reinterpret_cast<unsigned long long*>(sums);
  unsigned* counter = reinterpret_cast<unsigned*>(status + nb);
  LC3D_LAUNCH(ctx, scan_lookback_kernel, nb, kScanThreads, 0, in, out, n, status, counter);
}
inline size_t scan_scratch_bytes(int64_t n) { return (size_t)(div_up(n, kScanTile) + 2) * 8 + 16; }

// ------------------------------------------------------------ radix sort -----
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 8;                       // keys per thread
constexpr int kSortTile = kSortThreads * kSortRounds;  // 2048 keys per block
constexpr int kRadix = 256;

// hist[d * nblk + b] = number of keys of block b with digit d
__global__ void __launch_bounds__(kSortThreads) radix_hist(const uint32_t* __restrict__ keys,
                                                           int64_t n, int shift, int nblk,
                                                           uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[kRadix];
  for (int i = threadIdx.x; i < kRadix; i += kSortThreads) h[i] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = base + r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRadix; i += kSortThreads) hist[(size_t)i * nblk + blockIdx.x] = h[i];
}

// Stable scatter.  Warp w of block b owns the contiguous keys
// [b*tile + w*256, b*tile + (w+1)*256), processed as 8 rounds of 32 consecutive keys;
// __match_any_sync ranks equal digits inside a round in lane (= input) order.
__global__ void __launch_bounds__(kSortThreads)
    radix_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                  int shift, int nblk, const uint32_t* __restrict__ hist_scanned) {
  __shared__ uint32_t cnt[kSortWarps][kRadix + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * (kRadix + 1); i += kSortThreads)
    (&cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + (int64_t)w * (kSortRounds * 32);
  uint32_t key[kSortRounds], val[kSortRounds], rank[kSortRounds];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = wbase + r * 32 + lane;
    bool ok = i < n;
    key[r] = ok ? keys_in[i] : 0xffffffffu;
    val[r] = ok ? vals_in[i] : 0u;
    uint32_t d = ok ? ((key[r] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t prev = cnt[w][d];
    __syncwarp();
    if ((peers & lt) == 0) cnt[w][d] = prev + __popc(peers);
    __syncwarp();
    rank[r] = prev + __popc(peers & lt);
  }
  __syncthreads();
  // per-digit exclusive prefix over the warps of this block (thread d handles digit d)
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) {
    uint32_t run = hist_scanned[(size_t)d * nblk + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) {
      uint32_t c = cnt[ww][d];
      cnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    int64_t i = wbase + r * 32 + lane;
    if (i < n) {
      uint32_t d = (key[r] >> shift) & (kRadix - 1);
      uint32_t dst = cnt[w][d] + rank[r];
      keys_out[dst] = key[r];
      vals_out[dst] = val[r];
    }
  }
}

// NOTE: radix_hist above indexes keys block-strided, the scatter warp-contiguous; both
// cover exactly the same tile [b*tile, (b+1)*tile), so per-block digit counts agree.

struct SortScratch {
  uint32_t* keys_alt;
  uint32_t* vals_alt;
  uint32_t* hist;      // kRadix * nblk (+ scan sums appended)
  uint32_t* scan_sums;
};
inline size_t sort_hist_bytes(int64_t n) { return (size_t)kRadix * div_up(n, kSortTile) * 4 + 64; }

// Sorts (keys, vals) ascending by the low `bits` bits of key, stable.  The sorted data end up
// in (keys, vals) or in the scratch pair, depending on the parity of the pass count: the
// buffers holding the result are returned through out_keys / out_vals when given (no copy
// back); otherwise the result is copied back into (keys, vals).
inline void radix_sort_pairs(lc3d_ctx* ctx, uint32_t* keys, uint32_t* vals, int64_t n, int bits,
                             const SortScratch& s, uint32_t** out_keys = nullptr,
                             uint32_t** out_vals = nullptr) {
  if (out_keys) *out_keys = keys;
  if (out_vals) *out_vals = vals;
  if (n <= 1) return;
  int passes = (bits + 7) / 8;
  if (passes < 1) passes = 1;
  int nblk = div_up(n, kSortTile);
  uint32_t *kin = keys, *vin = vals, *kout = s.keys_alt, *vout = s.vals_alt;
  for (int p = 0; p < passes; ++p) {
    int shift = p * 8;
    LC3D_LAUNCH(ctx, radix_hist, nblk, kSortThreads, 0, kin, n, shift, nblk, s.hist);
    exclusive_scan_u32(ctx, s.hist, s.hist, (int64_t)kRadix * nblk, s.scan_sums);
    LC3D_LAUNCH(ctx, radix_scatter, nblk, kSortThreads, 0, kin, vin, kout, vout, n, shift, nblk,
                s.hist);
    std::swap(kin, kout);
    std::swap(vin, vout);
  }
  if (out_keys || out_vals) {
    if (out_keys) *out_keys = kin;
    if (out_vals) *out_vals = vin;
  } else if (kin != keys) {
    LC3D_CUDA(cudaMemcpyAsync(keys, kin, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    LC3D_CUDA(cudaMemcpyAsync(vals, vin, n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  }
}

}  // namespace lc3d
