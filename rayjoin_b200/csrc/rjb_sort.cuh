// Hand-written onesweep LSD radix sort of (u64 key, u32 value) pairs.
//
// Used by the LBVH build (Morton keys -> leaf order; replaces the
// thrust::stable_sort_by_key of reference deps/lbvh/lbvh/bvh.cuh:346-353), by
// the Morton ordering of queries and by the overlay (sort by edge id).
//
// Structure (Adinets & Merrill, "Onesweep"):
//   1. ONE histogram kernel reads the keys once and counts the 8-bit digits of
//      every pass (shared-memory privatised, then global atomics);
//   2. a 1-block kernel turns the per-pass histograms into exclusive digit bases;
//   3. per pass ONE kernel: a tile of 256 threads x 8 keys ranks its keys by
//      digit in shared memory (warp-level __match_any_sync multi-split, stable),
//      obtains the tile's global digit offsets with a chained scan over the
//      preceding tiles (decoupled look-back: aggregate / inclusive-prefix flags,
//      tile ids handed out by an atomic ticket so predecessors always run), and
//      writes keys and values out in digit-contiguous runs.
// HBM traffic: 8n (histogram read) + passes x 2 x 12n bytes; the sort is stable.
#pragma once
#include <algorithm>

#include "rjb_common.cuh"

namespace rjb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;  // 2048 keys per tile
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortMaxPasses = 8;
constexpr uint32_t kFlagAggregate = 1u << 30, kFlagInclusive = 2u << 30, kValueMask = (1u << 30) - 1;

// digit histograms of all passes in one sweep
__global__ void __launch_bounds__(kSortThreads)
k_sort_histogram(const uint64_t* __restrict__ keys, uint32_t n, int begin_bit, int end_bit,
                 int n_passes, uint32_t* __restrict__ hist /* [n_passes][256] */) {
  __shared__ uint32_t s_hist[kSortMaxPasses * 256];
  for (int i = threadIdx.x; i < n_passes * 256; i += kSortThreads) s_hist[i] = 0;
  __syncthreads();
  for (uint32_t i = blockIdx.x * kSortThreads + threadIdx.x; i < n; i += gridDim.x * kSortThreads) {
    const uint64_t k = keys[i];
    for (int p = 0; p < n_passes; p++) {
      const int sh = begin_bit + 8 * p;
      const uint32_t mask = (1u << min(8, end_bit - sh)) - 1;  // the last pass may be narrower
      atomicAdd(&s_hist[p * 256 + ((uint32_t) (k >> sh) & mask)], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_passes * 256; i += kSortThreads)
    if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// Histogram of packed words (key in the HIGH 32 bits): two words per 16-byte load, four loads
// of a thread in flight, shifts and masks of the passes hoisted -- 14 instead of 70 warp
// instructions per 32 keys (the generic kernel above was issue bound: ncu 67 % issue slots busy
// at 2.7 TB/s).
template <int NP>
__global__ void __launch_bounds__(kSortThreads)
k_sort_histogram_packed(const uint64_t* __restrict__ words, uint32_t n, int begin_bit, int end_bit,
                        uint32_t* __restrict__ hist /* [NP][256] */) {
  __shared__ uint32_t s_hist[NP * 256];
  for (int i = threadIdx.x; i < NP * 256; i += kSortThreads) s_hist[i] = 0;
  __syncthreads();
  uint32_t mask[NP];
#pragma unroll
  for (int p = 0; p < NP; p++) mask[p] = (1u << min(8, end_bit - (begin_bit + 8 * p))) - 1;
  const ulonglong2* __restrict__ w2 = reinterpret_cast<const ulonglong2*>(words);
  const uint32_t n2 = n / 2;
  constexpr int kLoads = 4;
  for (uint32_t i0 = blockIdx.x * (kSortThreads * kLoads); i0 < n2; i0 += gridDim.x * (kSortThreads * kLoads)) {
    ulonglong2 v[kLoads];
    bool ok[kLoads];
#pragma unroll
    for (int j = 0; j < kLoads; j++) {
      const uint32_t i = i0 + j * kSortThreads + threadIdx.x;
      ok[j] = i < n2;
      v[j] = ok[j] ? __ldg(&w2[i]) : make_ulonglong2(0, 0);
    }
#pragma unroll
    for (int j = 0; j < kLoads; j++) {
      if (!ok[j]) continue;
      const uint32_t ka = (uint32_t) (v[j].x >> 32) >> begin_bit, kb = (uint32_t) (v[j].y >> 32) >> begin_bit;
#pragma unroll
      for (int p = 0; p < NP; p++) {
        atomicAdd(&s_hist[p * 256 + ((ka >> (8 * p)) & mask[p])], 1u);
        atomicAdd(&s_hist[p * 256 + ((kb >> (8 * p)) & mask[p])], 1u);
      }
    }
  }
  if ((n & 1u) && blockIdx.x == 0 && threadIdx.x == 0) {
    const uint32_t k = (uint32_t) (words[n - 1] >> 32) >> begin_bit;
#pragma unroll
    for (int p = 0; p < NP; p++) atomicAdd(&s_hist[p * 256 + ((k >> (8 * p)) & mask[p])], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NP * 256; i += kSortThreads)
    if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// exclusive scan of each pass's 256 bins (one block, thread d owns digit d)
__global__ void __launch_bounds__(256)
k_sort_scan_hist(uint32_t* __restrict__ hist, int n_passes) {
  __shared__ uint32_t s[256];
  for (int p = 0; p < n_passes; p++) {
    const uint32_t v = hist[p * 256 + threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    // Hillis-Steele inclusive scan over 256 entries
    for (int o = 1; o < 256; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += t;
      __syncthreads();
    }
    hist[p * 256 + threadIdx.x] = s[threadIdx.x] - v;
    __syncthreads();
  }
}

// Decoupled look-back of thread d (digit d) of tile `tile`: the sum of the digit's counts over all
// preceding tiles.  The states of kLookBatch predecessors are requested TOGETHER and then consumed
// in order -- one round trip to L2 per batch instead of one per predecessor.  (ncu / launch list,
// 100 M random keys: with ~600 tiles in flight the nearest inclusive prefix is dozens of tiles
// back, and the one-at-a-time walk made a tile take 23 us: 0.95 ms per pass, 26 % of the HBM peak.)
constexpr int kLookBatch = 8;

static __device__ __forceinline__ uint32_t sort_lookback(volatile uint32_t* tile_state, uint32_t tile, uint32_t d) {
  uint32_t excl = 0;
  int j = (int) tile - 1;
  while (j >= 0) {
    uint32_t st[kLookBatch];
#pragma unroll
    for (int k = 0; k < kLookBatch; k++)
      st[k] = j - k >= 0 ? (uint32_t) tile_state[(size_t) (j - k) * 256 + d] : (uint32_t) (2u << 30);  // (before tile 0: inclusive, nothing)
    int used = 0;
    bool done = false;
#pragma unroll
    for (int k = 0; k < kLookBatch; k++) {
      if (done || used != k) continue;
      const uint32_t flag = st[k] & ~kValueMask;
      if (flag == 0) continue;  // not published yet: the next batch starts at this tile
      excl += st[k] & kValueMask;
      used = k + 1;
      done = flag == kFlagInclusive;
    }
    if (done) break;
    j -= used;
  }
  return excl;
}

__global__ void __launch_bounds__(kSortThreads)
k_sort_onesweep(const uint64_t* __restrict__ k_in, uint64_t* __restrict__ k_out,
                const uint32_t* __restrict__ v_in, uint32_t* __restrict__ v_out, uint32_t n,
                int shift, uint32_t mask, const uint32_t* __restrict__ digit_base /* [256] exclusive */,
                volatile uint32_t* tile_state /* [n_tiles][256], zeroed */,
                unsigned int* ticket /* zeroed */) {
  __shared__ uint64_t s_keys[kSortTile];
  __shared__ uint32_t s_vals[kSortTile];
  __shared__ uint32_t s_wcount[kSortWarps][256];  // per-warp digit counts -> exclusive offsets
  __shared__ uint32_t s_start[256];               // tile-local start of each digit
  __shared__ uint32_t s_gbase[256];               // global position of the tile's first key of a digit
  __shared__ unsigned int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&s_wcount[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_begin = tile * kSortTile;
  const uint32_t tile_n = min((uint32_t) kSortTile, n - tile_begin);

  // load (warp-striped: warp w owns items [w*256, w*256+256), lane-strided rounds
  // keep the input order) and rank within the warp, round by round
  uint64_t key[kSortItems];
  uint32_t val[kSortItems];
  uint32_t rank[kSortItems];  // rank among equal digits inside the warp (all earlier rounds included)
#pragma unroll
  for (int r = 0; r < kSortItems; r++) {
    const uint32_t local = warp * (32 * kSortItems) + r * 32 + lane;
    const bool ok = local < tile_n;
    key[r] = ok ? k_in[tile_begin + local] : ~0ull;
    val[r] = ok ? v_in[tile_begin + local] : 0;
    // padding keys get digit 255 and sit after every real key of that digit because
    // they are the last items of the last warps
    const uint32_t d = ok ? ((uint32_t) (key[r] >> shift) & mask) : 255u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) {
      base = s_wcount[warp][d];
      s_wcount[warp][d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[r] = base + __popc(peers & ((1u << lane) - 1));
    __syncwarp();
  }
  __syncthreads();

  // thread d: exclusive prefix of digit d over the warps, tile count of digit d
  const uint32_t d_own = threadIdx.x;  // kSortThreads == 256 digits
  uint32_t count = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; w++) {
    const uint32_t c = s_wcount[w][d_own];
    s_wcount[w][d_own] = count;
    count += c;
  }
  // the padding keys of a partial tile were counted under digit 255: remove them
  uint32_t real = count;
  if (d_own == 255) real = count - (kSortTile - tile_n);
  // publish this tile's contribution for the chained scan
  volatile uint32_t* my_state = tile_state + (size_t) tile * 256;
  if (tile == 0) my_state[d_own] = kFlagInclusive | real;
  else my_state[d_own] = kFlagAggregate | real;
  __threadfence();
  // tile-local exclusive scan over digits (warp scan + warp totals)
  uint32_t incl = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __shared__ uint32_t s_wtot[kSortWarps];
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; w++)
    if (w < warp) woff += s_wtot[w];
  s_start[d_own] = woff + incl - count;
  // decoupled look-back: sum the aggregates of the preceding tiles until one with
  // an inclusive prefix is met
  uint32_t excl = 0;
  if (tile > 0) {
    excl = sort_lookback(tile_state, tile, d_own);
    my_state[d_own] = kFlagInclusive | (excl + real);
  }
  s_gbase[d_own] = digit_base[d_own] + excl;
  __syncthreads();

  // scatter into shared memory in digit order, then write digit-contiguous runs
#pragma unroll
  for (int r = 0; r < kSortItems; r++) {
    const uint32_t local = warp * (32 * kSortItems) + r * 32 + lane;
    const uint32_t d = local < tile_n ? ((uint32_t) (key[r] >> shift) & mask) : 255u;
    const uint32_t pos = s_start[d] + s_wcount[warp][d] + rank[r];
    s_keys[pos] = key[r];
    s_vals[pos] = val[r];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < tile_n; i += kSortThreads) {
    const uint64_t k = s_keys[i];
    const uint32_t d = (uint32_t) (k >> shift) & mask;
    const uint32_t dst = s_gbase[d] + (i - s_start[d]);
    k_out[dst] = k;
    v_out[dst] = s_vals[i];
  }
}

// ---- packed pairs -------------------------------------------------------------------------
// Every sort of the engine has a key of at most 32 significant bits (Morton prefixes, grid cell
// ids, edge ids) and a 32-bit payload, so the pair travels as ONE 64-bit word, key in the high
// half: a pass moves 8 bytes per element each way instead of 12 through two arrays, the tile is
// one shared-memory array, and a tile holds 4096 pairs (16 per thread) so that the runs a tile
// writes per digit are twice as long.  Same structure as k_sort_onesweep above (warp-level
// multi-split, decoupled look-back, digit-contiguous write-out); stable.
constexpr int kSortPItems = 16;
constexpr int kSortPTile = kSortThreads * kSortPItems;  // 4096 words per tile

#ifndef RJB_SORT_MIN_CTAS
#define RJB_SORT_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kSortThreads, RJB_SORT_MIN_CTAS)
k_sort_onesweep_packed(const uint64_t* __restrict__ w_in, uint64_t* __restrict__ w_out, uint32_t n, int shift,
                       uint32_t mask, const uint32_t* __restrict__ digit_base /* [256] exclusive */,
                       volatile uint32_t* tile_state /* [n_tiles][256], zeroed */,
                       unsigned int* ticket /* zeroed */) {
  __shared__ uint64_t s_words[kSortPTile];
  __shared__ uint32_t s_wcount[kSortWarps][256];  // per-warp digit counts -> exclusive offsets
  __shared__ uint32_t s_start[256];               // tile-local start of each digit
  __shared__ uint32_t s_gbase[256];               // global position of the tile's first word of a digit
  __shared__ uint32_t s_wtot[kSortWarps];
  __shared__ unsigned int s_tile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&s_wcount[0][0])[i] = 0;
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t tile_begin = tile * kSortPTile;
  const uint32_t tile_n = min((uint32_t) kSortPTile, n - tile_begin);

  // all loads first (16 independent 8-byte loads per thread in flight), then the ranking
  uint64_t word[kSortPItems];
  uint16_t rank[kSortPItems];
#pragma unroll
  for (int r = 0; r < kSortPItems; r++) {
    const uint32_t local = warp * (32 * kSortPItems) + r * 32 + lane;
    word[r] = local < tile_n ? __ldg(&w_in[tile_begin + local]) : ~0ull;
  }
#pragma unroll
  for (int r = 0; r < kSortPItems; r++) {
    const uint32_t local = warp * (32 * kSortPItems) + r * 32 + lane;
    // padding words get digit 255 and sit after every real word of that digit because they
    // are the last items of the last warps
    const uint32_t d = local < tile_n ? ((uint32_t) (word[r] >> shift) & mask) : 255u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == leader) {
      base = s_wcount[warp][d];
      s_wcount[warp][d] = base + __popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    rank[r] = (uint16_t) (base + __popc(peers & ((1u << lane) - 1)));
    __syncwarp();
  }
  __syncthreads();

  // thread d: exclusive prefix of digit d over the warps, tile count of digit d
  const uint32_t d_own = threadIdx.x;  // kSortThreads == 256 digits
  uint32_t count = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; w++) {
    const uint32_t c = s_wcount[w][d_own];
    s_wcount[w][d_own] = count;
    count += c;
  }
  uint32_t real = count;
  if (d_own == 255) real = count - (kSortPTile - tile_n);  // minus the padding
  // (flag and value are ONE 32-bit word, written by one store and read by one load: nothing else
  // is published through it, so no fence is needed around it)
  volatile uint32_t* my_state = tile_state + (size_t) tile * 256;
  if (tile == 0) my_state[d_own] = kFlagInclusive | real;
  else my_state[d_own] = kFlagAggregate | real;
#ifdef RJB_SORT_FENCE
  __threadfence();
#endif
  uint32_t incl = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_wtot[warp] = incl;
  __syncthreads();
  uint32_t woff = 0;
#pragma unroll
  for (int w = 0; w < kSortWarps; w++)
    if (w < warp) woff += s_wtot[w];
  s_start[d_own] = woff + incl - count;
  uint32_t excl = 0;
  if (tile > 0) {
    excl = sort_lookback(tile_state, tile, d_own);
    my_state[d_own] = kFlagInclusive | (excl + real);
  }
  s_gbase[d_own] = digit_base[d_own] + excl;
  __syncthreads();

#pragma unroll
  for (int r = 0; r < kSortPItems; r++) {
    const uint32_t local = warp * (32 * kSortPItems) + r * 32 + lane;
    const uint32_t d = local < tile_n ? ((uint32_t) (word[r] >> shift) & mask) : 255u;
    s_words[s_start[d] + s_wcount[warp][d] + rank[r]] = word[r];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < tile_n; i += kSortThreads) {
    const uint64_t w = s_words[i];
    const uint32_t d = (uint32_t) (w >> shift) & mask;
    w_out[s_gbase[d] + (i - s_start[d])] = w;
  }
}

struct SortTemp {
  DBuf<uint8_t> hist, state, ticket;
};

// Sorts by key bits [begin_bit, end_bit).  The result is in (k_out, v_out); the
// inputs are used as ping-pong scratch when more than one pass is needed (their
// contents are destroyed).
static inline void sort_pairs_u64_u32(uint64_t* k_in, uint64_t* k_out, uint32_t* v_in,
                                      uint32_t* v_out, uint32_t n, int begin_bit, int end_bit,
                                      SortTemp& tmp, cudaStream_t st) {
  if (n == 0) return;
  RJB_REQUIRE(n < (1u << 30), "sort: at most 2^30 - 1 pairs");
  int n_passes = (end_bit - begin_bit + 7) / 8;
  RJB_REQUIRE(n_passes >= 1 && n_passes <= kSortMaxPasses, "sort: bad bit range");
  const uint32_t n_tiles = div_up(n, kSortTile);
  uint32_t* hist = (uint32_t*) tmp.hist.ensure(kSortMaxPasses * 256 * sizeof(uint32_t));
  uint32_t* state = (uint32_t*) tmp.state.ensure((size_t) n_passes * n_tiles * 256 * sizeof(uint32_t));
  unsigned int* ticket = (unsigned int*) tmp.ticket.ensure(kSortMaxPasses * sizeof(unsigned int));
  RJB_CUDA(cudaMemsetAsync(hist, 0, kSortMaxPasses * 256 * sizeof(uint32_t), st));
  RJB_CUDA(cudaMemsetAsync(state, 0, (size_t) n_passes * n_tiles * 256 * sizeof(uint32_t), st));
  RJB_CUDA(cudaMemsetAsync(ticket, 0, kSortMaxPasses * sizeof(unsigned int), st));
  const unsigned hb = min(div_up(n, kSortThreads * 16), (unsigned) (kNumSMs * 8));
  k_sort_histogram<<<hb, kSortThreads, 0, st>>>(k_in, n, begin_bit, end_bit, n_passes, hist);
  k_sort_scan_hist<<<1, 256, 0, st>>>(hist, n_passes);
  // ping-pong in -> out -> in ...; an odd number of passes ends in the out buffers
  uint64_t* ks = k_in;
  uint32_t* vs = v_in;
  uint64_t* kd = k_out;
  uint32_t* vd = v_out;
  for (int p = 0; p < n_passes; p++) {
    const uint32_t mask = (1u << std::min(8, end_bit - (begin_bit + 8 * p))) - 1;
    k_sort_onesweep<<<n_tiles, kSortThreads, 0, st>>>(ks, kd, vs, vd, n, begin_bit + 8 * p, mask,
                                                      hist + p * 256,
                                                      state + (size_t) p * n_tiles * 256, ticket + p);
    uint64_t* tk = ks; ks = kd; kd = tk;
    uint32_t* tv = vs; vs = vd; vd = tv;
  }
  RJB_CUDA(cudaGetLastError());
  // after the loop `ks`/`vs` point at the buffers written last
  if (ks != k_out) {
    RJB_CUDA(cudaMemcpyAsync(k_out, ks, (size_t) n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    RJB_CUDA(cudaMemcpyAsync(v_out, vs, (size_t) n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
}

// Sorts packed words (key in the high 32 bits, payload in the low 32) by key bits
// [begin_bit, end_bit) of the KEY.  `a` holds the input; `b` is scratch of the same size.
// Returns the buffer that holds the result (a or b): no final copy.
static inline uint64_t* sort_packed(uint64_t* a, uint64_t* b, uint32_t n, int begin_bit, int end_bit,
                                    SortTemp& tmp, cudaStream_t st) {
  if (n == 0) return a;
  RJB_REQUIRE(n < (1u << 30), "sort: at most 2^30 - 1 pairs");
  RJB_REQUIRE(begin_bit >= 0 && end_bit <= 32 && begin_bit < end_bit, "sort: bad key bit range");
  const int n_passes = (end_bit - begin_bit + 7) / 8;
  const uint32_t n_tiles = div_up(n, kSortPTile);
  uint32_t* hist = (uint32_t*) tmp.hist.ensure(kSortMaxPasses * 256 * sizeof(uint32_t));
  uint32_t* state = (uint32_t*) tmp.state.ensure((size_t) n_passes * n_tiles * 256 * sizeof(uint32_t));
  unsigned int* ticket = (unsigned int*) tmp.ticket.ensure(kSortMaxPasses * sizeof(unsigned int));
  RJB_CUDA(cudaMemsetAsync(hist, 0, kSortMaxPasses * 256 * sizeof(uint32_t), st));
  RJB_CUDA(cudaMemsetAsync(state, 0, (size_t) n_passes * n_tiles * 256 * sizeof(uint32_t), st));
  RJB_CUDA(cudaMemsetAsync(ticket, 0, kSortMaxPasses * sizeof(unsigned int), st));
  const unsigned hp = min(div_up(n, kSortThreads * 8), (unsigned) (kNumSMs * 8));
  switch (n_passes) {
    case 1: k_sort_histogram_packed<1><<<hp, kSortThreads, 0, st>>>(a, n, begin_bit, end_bit, hist); break;
    case 2: k_sort_histogram_packed<2><<<hp, kSortThreads, 0, st>>>(a, n, begin_bit, end_bit, hist); break;
    case 3: k_sort_histogram_packed<3><<<hp, kSortThreads, 0, st>>>(a, n, begin_bit, end_bit, hist); break;
    default: k_sort_histogram_packed<4><<<hp, kSortThreads, 0, st>>>(a, n, begin_bit, end_bit, hist); break;
  }
  k_sort_scan_hist<<<1, 256, 0, st>>>(hist, n_passes);
  uint64_t* src = a;
  uint64_t* dst = b;
  for (int p = 0; p < n_passes; p++) {
    const uint32_t mask = (1u << std::min(8, end_bit - (begin_bit + 8 * p))) - 1;
    k_sort_onesweep_packed<<<n_tiles, kSortThreads, 0, st>>>(src, dst, n, 32 + begin_bit + 8 * p, mask, hist + p * 256,
                                                             state + (size_t) p * n_tiles * 256, ticket + p);
    std::swap(src, dst);
  }
  RJB_CUDA(cudaGetLastError());
  return src;  // written last
}

}  // namespace rjb
