// Device primitives: RAII buffers, a hand-written exclusive scan, and the
// radix sort used by the LBVH / grid builders and the overlay.
#pragma once
#include "rjb_common.cuh"

namespace rjb {

// Grow-only device buffer (allocations are reused across queries).
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // contents are NOT preserved when growing
  T* ensure(size_t n) {
    if (n > cap) {
      release();
      size_t want = n + n / 8 + 16;
      RJB_CUDA(cudaMalloc((void**) &p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

// ---------------------------------------------------------------------------
// exclusive scan of uint32 (n + 1 outputs; out[n] = total).  Three passes:
// per-tile sums, one-block scan of the tile sums, per-tile scan + offset.
// HBM traffic: 4n (read) + 4n (read) + 4n (write) bytes.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

static __device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// exclusive scan across the block of one value per thread; returns the
// exclusive prefix, *total = block sum
static __device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[kScanThreads / 32];
  __shared__ uint32_t wtot;
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = warp_incl_scan(v, lane);
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t s = lane < kScanThreads / 32 ? wsum[lane] : 0;
    uint32_t si = warp_incl_scan(s, lane);
    if (lane < kScanThreads / 32) wsum[lane] = si - s;
    if (lane == kScanThreads / 32 - 1) wtot = si;
  }
  __syncthreads();
  uint32_t r = inc - v + wsum[w];
  *total = wtot;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_tile_sums(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ tile_sums) {
  uint32_t base = blockIdx.x * kScanTile;
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    uint32_t idx = base + i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  uint32_t tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_tile_offsets(uint32_t* __restrict__ tile_sums, uint32_t n_tiles, uint32_t* __restrict__ total_out) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_tiles; base += kScanThreads) {
    uint32_t idx = base + threadIdx.x;
    uint32_t v = idx < n_tiles ? tile_sums[idx] : 0;
    uint32_t tot;
    uint32_t ex = block_excl_scan(v, &tot);
    if (idx < n_tiles) tile_sums[idx] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const uint32_t* __restrict__ in, uint32_t n, const uint32_t* __restrict__ tile_offs,
             uint32_t* __restrict__ out) {
  // thread owns kScanItems consecutive elements (blocked arrangement)
  uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  uint32_t tot;
  uint32_t ex = block_excl_scan(s, &tot) + tile_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; i++) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

struct ScanTemp {
  DBuf<uint32_t> tiles;
};

// out must hold n + 1 elements; in == out is allowed.
static inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n,
                                      ScanTemp& tmp, cudaStream_t st) {
  uint32_t n_tiles = div_up(n, kScanTile);
  if (n_tiles == 0) {
    RJB_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t), st));
    return;
  }
  uint32_t* tiles = tmp.tiles.ensure(n_tiles);
  k_scan_tile_sums<<<n_tiles, kScanThreads, 0, st>>>(in, n, tiles);
  k_scan_tile_offsets<<<1, kScanThreads, 0, st>>>(tiles, n_tiles, out + n);
  k_scan_apply<<<n_tiles, kScanThreads, 0, st>>>(in, n, tiles, out);
  RJB_CUDA(cudaGetLastError());
}

}  // namespace rjb

#include "rjb_sort.cuh"
