// LBVH over the edges of one map: leaves of <= G consecutive chain edges,
// 64-bit 2-D Morton keys of the leaf-box centre, radix sort, Karras hierarchy,
// bottom-up refit of quantised integer boxes.
//
// Replaces FillPrimitivesLBVH + lbvh::bvh::construct
// (reference: src/tree/primtive.h:33-57, deps/lbvh/lbvh/bvh.cuh:277-481).
// Differences that matter: boxes are conservative in the INTEGER domain
// (coord >> 16, floor on both ends is monotone so exact overlap => quantised
// overlap) instead of floats widened by 2 ulp; keys are 64-bit Morton codes
// with an index tie-break, so no uniqueness pass is needed; a node stores the
// boxes of both children so traversal reads one record per visit.
#pragma once
#include "rjb_prims.cuh"

namespace rjb {

constexpr int kMortonDrop = 32;  // Morton bits ignored by the leaf order (keys: 16 bits per axis)

struct Bvh {
  DBuf<int4> node_box;
  DBuf<int2> node_child;
  DBuf<uint2> leaf_rec;
  int4 root_box = {0, 0, -1, -1};
  uint32_t n_leaves = 0;
  int leaf_size = 0;
  bool built = false;
  // build scratch (kept for rebuilds)
  DBuf<uint32_t> chain_cnt, leaf_base, vals_a, vals_b, parent, flags, ag_packed, ag_cnt, ag_off;
  DBuf<uint64_t> keys_a, keys_b;
  DBuf<int4> leaf_box_u, leaf_box_s;
  DBuf<uint2> leaf_rec_u;
  DBuf<int4> root_box_d;
  DBuf<int4> top_box;
  DBuf<int> top_code;
  DBuf<uint32_t> occ, occ_pop, occ_rank, cell_cnt, cell_begin;
  DBuf<uint2> occ_dir;
  DBuf<uint4> cell_item;
  DBuf<unsigned long long> inc_counter;
  bool have_cells = false;
  uint32_t n_occ_cells = 0, n_incidences = 0;
  int top_levels = 3;
  double occ_fraction = 1.0;  // share of occupied cells (the filter pays off when it is small)
  ScanTemp scan_tmp;
  SortTemp sort_tmp;

  BvhView view() const {
    BvhView v;
    v.node_box = node_box.p;
    v.node_child = node_child.p;
    v.leaf_rec = leaf_rec.p;
    v.root_box = root_box;
    v.n_leaves = n_leaves;
    v.top_box = top_box.p;
    v.top_code = top_code.p;
    v.occ = occ.p;
    v.occ_dir = have_cells ? occ_dir.p : nullptr;
    v.cell_begin = cell_begin.p;
    v.cell_item = cell_item.p;
    v.top_levels = top_levels;
    return v;
  }
  size_t cell_directory_bytes() const {
    return have_cells ? (size_t) (kOccWords + 1) * sizeof(uint2) + (size_t) (n_occ_cells + 1) * 4 +
                            (size_t) n_incidences * sizeof(uint4) : 0;
  }
  size_t index_bytes() const {
    uint32_t n_int = n_leaves > 1 ? n_leaves - 1 : 1;
    return (size_t) n_int * (2 * sizeof(int4) + sizeof(int2)) + (size_t) n_leaves * sizeof(uint2) +
           (size_t) (top_levels == 4 ? kTopSlots4 : kTopSlots3) * (sizeof(int4) + sizeof(int)) +
           (size_t) kOccMaps * kOccDim * kOccDim / 8 + cell_directory_bytes();
  }
};

__global__ void k_chain_leaf_counts(const uint32_t* __restrict__ row_index, uint32_t n_chains,
                                    int G, uint32_t* __restrict__ cnt) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  uint32_t ne = row_index[c + 1] - row_index[c] - 1;
  cnt[c] = (ne + G - 1) / G;
}

static __device__ __forceinline__ uint64_t spread32(uint32_t v) {
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
  x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
  x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}

// 64-bit Morton code of a point in internal coordinates (47-bit range)
static __device__ __forceinline__ uint64_t morton64(long long x, long long y, long long imin) {
  uint32_t ux = (uint32_t) ((unsigned long long) (x - imin) >> 15);
  uint32_t uy = (uint32_t) ((unsigned long long) (y - imin) >> 15);
  return (spread32(uy) << 1) | spread32(ux);
}

// one thread per leaf: locate the chain, emit {record, quantised box, key}
__global__ void k_leaf_fill(MapView m, const uint32_t* __restrict__ leaf_base, uint32_t n_leaves,
                            int G, long long imin, uint2* __restrict__ rec,
                            int4* __restrict__ box, uint64_t* __restrict__ key) {
  uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_leaves) return;
  // chain c with leaf_base[c] <= l < leaf_base[c+1]
  uint32_t lo = 0, hi = m.n_chains;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (leaf_base[mid] <= l) lo = mid; else hi = mid;
  }
  uint32_t c = lo;
  uint32_t p_begin = m.row_index[c], p_end = m.row_index[c + 1];
  uint32_t k = l - leaf_base[c];
  uint32_t p1 = p_begin + k * G;                   // first point of the leaf
  uint32_t cnt = min((uint32_t) G, p_end - 1 - p1);  // edges in the leaf
  long long xmin, xmax, ymin, ymax;
  longlong2 p = m.pts[p1];
  xmin = xmax = p.x;
  ymin = ymax = p.y;
  for (uint32_t i = 1; i <= cnt; i++) {
    p = m.pts[p1 + i];
    xmin = min(xmin, p.x); xmax = max(xmax, p.x);
    ymin = min(ymin, p.y); ymax = max(ymax, p.y);
  }
  rec[l] = make_uint2(p1 - c, (cnt << 28) | c);
  box[l] = make_int4(quant(xmin), quant(ymin), quant(xmax), quant(ymax));
  // 32 significant key bits (cells of 2^-16 of the range per axis, about one edge
  // length on a 10 M-edge map) = 4 radix passes;
  // the low bits are cleared so that the sorted keys stay monotone for Karras' delta()
  // (packed with the leaf id in the cleared low half: sort_packed, rjb_sort.cuh)
  key[l] = (morton64(xmin + ((xmax - xmin) >> 1), ymin + ((ymax - ymin) >> 1), imin) & ~((1ull << kMortonDrop) - 1)) | l;
}

// ---- adaptive leaf grouping ---------------------------------------------------------------
// RayJoin's Adaptive Grouping (reference src/rt/primitive.h:120-260, flags -ag -ag_iter
// -enlarge) merges the boxes of CONSECUTIVE edges while the merged box stays small: in rounds,
// neighbours (2i, 2i+1) of the current group list are merged iff
// area(merged) / max(area(a), area(b)) < enlarge, until a round merges nothing or max_iter
// rounds have run.  There it shrinks the number of RT primitives (86.6 % fewer on County);
// here the same rule sizes the LBVH leaves: a straight run of a chain becomes one leaf of up
// to 2^iter edges, a sharp corner keeps its edges apart.  Leaves stay runs of ONE chain (the
// leaf record is {first eid, count, chain}) of at most 8 edges, so at most 3 rounds take
// effect, inside aligned blocks of 8 edges of a chain; areas are those of the exact integer
// boxes (+1 so that an axis-parallel edge has a non-zero area), not of float boxes.
constexpr int kAgBlock = 8;

// groups of one block of <= 8 consecutive edges (points p0 .. p0 + ne): returns the group
// lengths packed 4 bits each, first group in the low nibble
static __device__ __forceinline__ uint32_t ag_groups(const longlong2* __restrict__ pts, uint32_t p0,
                                                     uint32_t ne, int max_iter, float enlarge) {
  long long x0[kAgBlock], y0[kAgBlock], x1[kAgBlock], y1[kAgBlock];
  int len[kAgBlock];
  longlong2 a = pts[p0];
  for (uint32_t k = 0; k < (uint32_t) kAgBlock; k++) {
    if (k < ne) {
      const longlong2 b = pts[p0 + k + 1];
      x0[k] = min(a.x, b.x); x1[k] = max(a.x, b.x);
      y0[k] = min(a.y, b.y); y1[k] = max(a.y, b.y);
      a = b;
    } else {
      x0[k] = y0[k] = x1[k] = y1[k] = 0;
    }
    len[k] = 1;
  }
  int n = (int) ne;
  auto area = [](long long ax0, long long ay0, long long ax1, long long ay1) {
    return (double) (ax1 - ax0 + 1) * (double) (ay1 - ay0 + 1);
  };
  for (int it = 0; it < max_iter; it++) {
    int m = 0;
    for (int i = 0; i < n; i += 2) {
      if (i + 1 < n) {
        const long long mx0 = min(x0[i], x0[i + 1]), my0 = min(y0[i], y0[i + 1]);
        const long long mx1 = max(x1[i], x1[i + 1]), my1 = max(y1[i], y1[i + 1]);
        const double big = fmax(area(x0[i], y0[i], x1[i], y1[i]), area(x0[i + 1], y0[i + 1], x1[i + 1], y1[i + 1]));
        if ((float) (area(mx0, my0, mx1, my1) / big) < enlarge) {
          const int l = len[i] + len[i + 1];
          x0[m] = mx0; y0[m] = my0; x1[m] = mx1; y1[m] = my1;
          len[m] = l;
          m++;
          continue;
        }
        // not merged: both survive (m <= i, so nothing unread is overwritten)
        const long long bx0 = x0[i + 1], by0 = y0[i + 1], bx1 = x1[i + 1], by1 = y1[i + 1];
        const int bl = len[i + 1];
        x0[m] = x0[i]; y0[m] = y0[i]; x1[m] = x1[i]; y1[m] = y1[i]; len[m] = len[i];
        m++;
        x0[m] = bx0; y0[m] = by0; x1[m] = bx1; y1[m] = by1; len[m] = bl;
        m++;
      } else {
        x0[m] = x0[i]; y0[m] = y0[i]; x1[m] = x1[i]; y1[m] = y1[i]; len[m] = len[i];
        m++;
      }
    }
    if (m == n) break;
    n = m;
  }
  uint32_t packed = 0;
  for (int i = 0; i < n; i++) packed |= (uint32_t) len[i] << (4 * i);
  return packed;
}

// block b of chain c = edges [8k, 8k + 8) of the chain (block_base = blocks before the chain)
static __device__ __forceinline__ void ag_locate(const MapView& m, const uint32_t* __restrict__ block_base,
                                                 uint32_t b, uint32_t& c, uint32_t& p0, uint32_t& ne) {
  uint32_t lo = 0, hi = m.n_chains;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (block_base[mid] <= b) lo = mid; else hi = mid;
  }
  c = lo;
  const uint32_t p_begin = m.row_index[c], p_end = m.row_index[c + 1];
  p0 = p_begin + (b - block_base[c]) * kAgBlock;
  ne = min((uint32_t) kAgBlock, p_end - 1 - p0);
}

__global__ void k_ag_group(MapView m, const uint32_t* __restrict__ block_base, uint32_t n_blocks, int max_iter,
                           float enlarge, uint32_t* __restrict__ packed, uint32_t* __restrict__ cnt) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint32_t c, p0, ne;
  ag_locate(m, block_base, b, c, p0, ne);
  const uint32_t g = ag_groups(m.pts, p0, ne, max_iter, enlarge);
  packed[b] = g;
  uint32_t n = 0;
  for (uint32_t t = g; t; t >>= 4) n++;
  cnt[b] = n;
}

// one thread per block: emits the leaves of its groups {record, quantised box, key}
__global__ void k_leaf_fill_ag(MapView m, const uint32_t* __restrict__ block_base, uint32_t n_blocks,
                               const uint32_t* __restrict__ packed, const uint32_t* __restrict__ leaf_off,
                               long long imin, uint2* __restrict__ rec, int4* __restrict__ box,
                               uint64_t* __restrict__ key) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint32_t c, p0, ne;
  ag_locate(m, block_base, b, c, p0, ne);
  uint32_t l = leaf_off[b];
  uint32_t p1 = p0;
  for (uint32_t g = packed[b]; g; g >>= 4, l++) {
    const uint32_t cntg = g & 15u;
    longlong2 p = m.pts[p1];
    long long xmin = p.x, xmax = p.x, ymin = p.y, ymax = p.y;
    for (uint32_t i = 1; i <= cntg; i++) {
      p = m.pts[p1 + i];
      xmin = min(xmin, p.x); xmax = max(xmax, p.x);
      ymin = min(ymin, p.y); ymax = max(ymax, p.y);
    }
    rec[l] = make_uint2(p1 - c, (cntg << 28) | c);
    box[l] = make_int4(quant(xmin), quant(ymin), quant(xmax), quant(ymax));
    key[l] = (morton64(xmin + ((xmax - xmin) >> 1), ymin + ((ymax - ymin) >> 1), imin) & ~((1ull << kMortonDrop) - 1)) | l;
    p1 += cntg;
  }
}

__global__ void k_leaf_gather(const uint64_t* __restrict__ sorted, uint32_t n,
                              const uint2* __restrict__ rec_u, const int4* __restrict__ box_u,
                              uint2* __restrict__ rec_s, int4* __restrict__ box_s) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t j = (uint32_t) sorted[i];  // packed (key, leaf id)
  rec_s[i] = rec_u[j];
  box_s[i] = box_u[j];
}

// Karras 2012: common-prefix length of sorted keys i and j, index tie-break
static __device__ __forceinline__ int delta(const uint64_t* __restrict__ key, uint32_t n, int i, int j) {
  if (j < 0 || j >= (int) n) return -1;
  // packed words: the key is the high half
  uint64_t a = key[i] & 0xFFFFFFFF00000000ull, b = key[j] & 0xFFFFFFFF00000000ull;
  if (a == b) return 64 + __clz((uint32_t) i ^ (uint32_t) j);
  return __clzll((long long) (a ^ b));
}

// parent[] is indexed by node id: internal i -> i, leaf j -> (n - 1) + j
__global__ void k_karras(const uint64_t* __restrict__ key, uint32_t n, int2* __restrict__ child,
                         uint32_t* __restrict__ parent) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int) n - 1) return;
  int d = (delta(key, n, i, i + 1) - delta(key, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(key, n, i, i - d);
  int lmax = 2;
  while (delta(key, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(key, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(key, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (delta(key, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  int gamma = i + s * d + min(d, 0);
  int left, right;
  if (min(i, j) == gamma) { left = ~gamma; parent[(n - 1) + gamma] = i; }
  else { left = gamma; parent[gamma] = i; }
  if (max(i, j) == gamma + 1) { right = ~(gamma + 1); parent[(n - 1) + gamma + 1] = i; }
  else { right = gamma + 1; parent[gamma + 1] = i; }
  child[i] = make_int2(left, right);
  if (i == 0) parent[0] = 0xFFFFFFFFu;
}

static __device__ __forceinline__ int4 box_union(const int4& a, const int4& b) {
  return make_int4(min(a.x, b.x), min(a.y, b.y), max(a.z, b.z), max(a.w, b.w));
}

// 128-bit atomic exchange (ATOMG.E.EXCH.128 on sm_90+): deposits a whole box and
// returns the previous content in one L2 operation.
static __device__ __forceinline__ int4 atom_exch_box(int4* addr, const int4& v) {
  unsigned long long vlo = ((unsigned long long) (unsigned) v.y << 32) | (unsigned) v.x;
  unsigned long long vhi = ((unsigned long long) (unsigned) v.w << 32) | (unsigned) v.z;
  unsigned long long olo, ohi;
  asm volatile(
      "{\n\t.reg .b128 vv, oo;\n\tmov.b128 vv, {%2, %3};\n\t"
      "atom.global.relaxed.gpu.exch.b128 oo, [%4], vv;\n\tmov.b128 {%0, %1}, oo;\n\t}"
      : "=l"(olo), "=l"(ohi)
      : "l"(vlo), "l"(vhi), "l"(addr)
      : "memory");
  return make_int4((int) (unsigned) olo, (int) (unsigned) (olo >> 32), (int) (unsigned) ohi,
                   (int) (unsigned) (ohi >> 32));
}

__global__ void k_refit_init(int4* node_box, uint32_t n_int) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_int) node_box[2 * i] = empty_box();  // "nobody has arrived yet"
}

// Bottom-up refit.  One thread per leaf climbs; at every internal node the first
// arrival DEPOSITS its box in the node's first slot with a single 128-bit atomic
// exchange and stops, the second arrival gets that box back from the same exchange
// (so no flag, no fence, no second read), writes both child boxes into the node
// record and carries the union upward.  (The reference's refit,
// deps/lbvh/lbvh/bvh.cuh:425-460, uses an atomicCAS flag without a fence between
// the box store and the hand-off.)
__global__ void k_refit(const int4* __restrict__ leaf_box, uint32_t n, const int2* __restrict__ child,
                        const uint32_t* __restrict__ parent, int4* node_box, int4* root_box) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int4 box = leaf_box[j];
  int me = ~(int) j;
  uint32_t cur = parent[(n - 1) + j];
  const int4 none = empty_box();
  while (true) {
    const int4 old = atom_exch_box(&node_box[2 * cur], box);
    if (old.x == none.x && old.y == none.y && old.z == none.z && old.w == none.w) return;
    const bool left = child[cur].x == me;
    node_box[2 * cur] = left ? box : old;
    node_box[2 * cur + 1] = left ? old : box;
    box = box_union(box, old);
    me = (int) cur;
    const uint32_t up = parent[cur];
    if (up == 0xFFFFFFFFu) {
      *root_box = box;
      return;
    }
    cur = up;
  }
}

__global__ void k_single_leaf_root(const int4* __restrict__ leaf_box, int4* node_box,
                                   int2* child, int4* root_box) {
  node_box[0] = leaf_box[0];
  node_box[1] = empty_box();  // empty: never overlaps
  child[0] = make_int2(~0, ~0);
  *root_box = leaf_box[0];
}

// 32-ary top tree: thread P (a 15- or 20-bit root path) walks that many binary levels
// and publishes the node it stands on after 5, 10, 15 (and 20) steps into the slot
// named by the path prefix.  A leaf met early stays in the all-zero continuation of
// its path; every other slot below it is empty.
__global__ void k_top_tree(const int4* __restrict__ node_box, const int2* __restrict__ node_child,
                           const int4* __restrict__ root_box, uint32_t n_leaves, int depth_max,
                           int4* __restrict__ top_box, int* __restrict__ top_code) {
  uint32_t P = blockIdx.x * blockDim.x + threadIdx.x;
  if (P >= (1u << depth_max)) return;
  int code = 0;  // root = internal node 0
  int4 box = *root_box;
  bool alive = n_leaves > 0;
  const int4 empty = empty_box();
  for (int level = 0; level < depth_max; level++) {
    int bit = (P >> (depth_max - 1 - level)) & 1;
    if (alive) {
      if (code < 0) {
        if (bit) alive = false;  // a leaf lives only in the zero continuation
      } else {
        int2 ch = node_child[code];
        box = node_box[2 * code + bit];
        code = bit ? ch.y : ch.x;
        if (box.x > box.z) alive = false;  // empty child of the single-leaf root (empty_box)
      }
    }
    int depth = level + 1;
    if (depth % 5 == 0) {
      int rest = depth_max - depth;  // lower path bits must be zero: one writer per slot
      if ((P & ((1u << rest) - 1)) == 0) {
        int off = depth == 5 ? kTopOff0 : (depth == 10 ? kTopOff1 : (depth == 15 ? kTopOff2 : kTopOff3));
        uint32_t slot = P >> rest;
        top_box[off + slot] = alive ? box : empty;
        top_code[off + slot] = code;
      }
    }
  }
}

// occupancy bitmap: every cell the (quantised) box of a leaf touches.  Leaves, not
// edges: 4x fewer boxes, and the leaf box is what the traversal would test anyway.
// (A warp-merged variant with __match_any_sync measured slower than plain atomics:
// 71 vs 39 us for 1 M leaves.)
//
// Rows of a box are marked a bitmap WORD at a time (one atomicOr per 32 cells), and a
// box of more than kOccBigWords words is not walked by its thread at all: it goes to a
// list that k_occ_mark_big works off with one CTA per box, so a map-spanning edge costs
// 512 K word operations spread over 256 threads instead of 16.7 M atomics in one thread.
constexpr uint32_t kOccBigWords = 256;

static __device__ __forceinline__ void occ_mark_row_words(uint32_t* __restrict__ occ, int y, int x0, int x1,
                                                          int w) {
  // word w of row y, cells [x0, x1] clipped to the word
  const int lo = max(x0, 32 * w), hi = min(x1, 32 * w + 31);
  const uint32_t m = (hi - lo == 31) ? 0xFFFFFFFFu : (((1u << (hi - lo + 1)) - 1) << (lo & 31));
  uint32_t* p = &occ[(uint32_t) y * (kOccDim / 32) + w];
  if ((*p & m) != m) atomicOr(p, m);  // mostly set already
}

// stats: [0] (leaf, cell) incidences, [1] cells of the largest box, [2] boxes on the big list
__global__ void k_occ_mark(const int4* __restrict__ leaf_box, uint32_t n, uint32_t* __restrict__ occ,
                           unsigned long long* __restrict__ stats, uint32_t* __restrict__ big_list) {
  uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long area = 0;
  if (l < n) {
    const int4 b = leaf_box[l];
    const int x0 = occ_cell(b.x), x1 = occ_cell(b.z), y0 = occ_cell(b.y), y1 = occ_cell(b.w);
    area = (unsigned long long) (x1 - x0 + 1) * (unsigned long long) (y1 - y0 + 1);
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    if ((uint32_t) (w1 - w0 + 1) * (uint32_t) (y1 - y0 + 1) > kOccBigWords) {
      big_list[atomicAdd(&stats[2], 1ull)] = l;
    } else {
      for (int y = y0; y <= y1; y++)
        for (int w = w0; w <= w1; w++) occ_mark_row_words(occ, y, x0, x1, w);
    }
    if (area > 1024) atomicMax(&stats[1], area);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
  if ((threadIdx.x & 31) == 0 && area) atomicAdd(&stats[0], area);
}

// one CTA per box of the big list
__global__ void __launch_bounds__(256)
k_occ_mark_big(const int4* __restrict__ leaf_box, const uint32_t* __restrict__ big_list,
               const unsigned long long* __restrict__ stats, uint32_t* __restrict__ occ) {
  const uint32_t n_big = (uint32_t) stats[2];
  for (uint32_t i = blockIdx.x; i < n_big; i += gridDim.x) {
    const int4 b = leaf_box[big_list[i]];
    const int x0 = occ_cell(b.x), x1 = occ_cell(b.z), y0 = occ_cell(b.y), y1 = occ_cell(b.w);
    const int w0 = x0 >> 5, nw = (x1 >> 5) - w0 + 1;
    const uint32_t total = (uint32_t) nw * (uint32_t) (y1 - y0 + 1);
    for (uint32_t t = threadIdx.x; t < total; t += blockDim.x)
      occ_mark_row_words(occ, y0 + (int) (t / nw), x0, x1, w0 + (int) (t % nw));
  }
}

__global__ void k_occ_popc(const uint32_t* __restrict__ occ, uint32_t* __restrict__ pop) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < kOccWords) pop[w] = __popc(occ[w]);
}

// id of an occupied cell (bit must be set in occ)
static __device__ __forceinline__ uint32_t occ_cell_id(const uint32_t* __restrict__ occ,
                                                       const uint32_t* __restrict__ rank, uint32_t bit) {
  return __ldg(&rank[bit >> 5]) + __popc(__ldg(&occ[bit >> 5]) & ((1u << (bit & 31)) - 1));
}

// cell directory, pass 1 (fill == false): leaves per occupied cell; pass 2: the item records
template <bool kFill>
__global__ void k_cell_lists(const int4* __restrict__ leaf_box, const uint2* __restrict__ leaf_rec, uint32_t n,
                             const uint32_t* __restrict__ occ, const uint32_t* __restrict__ rank,
                             uint32_t* __restrict__ cnt, const uint32_t* __restrict__ begin,
                             uint4* __restrict__ cell_item) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n) return;
  const int4 b = leaf_box[l];
  uint32_t first_point = 0, count = 1;
  if (kFill) {
    const uint2 rec = leaf_rec[l];
    first_point = rec.x + (rec.y & 0x0FFFFFFFu);
    count = rec.y >> 28;
  }
  const int x0 = occ_cell(b.x), x1 = occ_cell(b.z), y0 = occ_cell(b.y), y1 = occ_cell(b.w);
  for (int y = y0; y <= y1; y++)
    for (int x = x0; x <= x1; x++) {
      const uint32_t id = occ_cell_id(occ, rank, (uint32_t) y * kOccDim + x);
      const uint32_t k = atomicAdd(&cnt[id], 1u);
      if (kFill) cell_item[__ldg(&begin[id]) + k] = cell_item_of(b, x, y, first_point, count);
    }
}

// {bitmap word, rank} side by side: one 8-byte load finds a cell
__global__ void k_occ_dir(const uint32_t* __restrict__ occ, const uint32_t* __restrict__ rank,
                          uint2* __restrict__ dir) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w <= kOccWords) dir[w] = make_uint2(w < kOccWords ? occ[w] : 0u, rank[w]);
}

// One doubling step of the dilation: out(x, y) = in(x, y) | in(x+s, y) | in(x, y+s) | in(x+s, y+s),
// one word per thread.  occ -> occ2 (s = 1) -> occ4 (s = 2) -> occ8 (s = 4): occK(x, y) is the OR
// of occ over the K x K cells at (x, y), what a box of that size at that min corner can touch.
__global__ void k_occ_dilate(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int s) {
  constexpr uint32_t kRowWords = kOccDim / 32;
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= kOccWords) return;
  const uint32_t y = w / kRowWords, wx = w % kRowWords;
  const bool has_row = y + s < (uint32_t) kOccDim, has_col = wx + 1 < kRowWords;
  const uint32_t a = in[w] | (has_row ? in[w + s * kRowWords] : 0u);
  const uint32_t an = has_col ? (in[w + 1] | (has_row ? in[w + 1 + s * kRowWords] : 0u)) : 0u;
  out[w] = a | (a >> s) | (an << (32 - s));
}

// ag_iter > 0: adaptive leaf grouping with that many merge rounds (at most 3 take effect) and
// area limit `enlarge`; else fixed runs of leaf_size edges
static inline void build_lbvh(Bvh& b, const MapView& m, int leaf_size, int ag_iter, float enlarge, long long imin,
                              bool want_cells, cudaStream_t st) {
  RJB_REQUIRE(leaf_size >= 1 && leaf_size <= 8, "lbvh_leaf_size must be in 1..8");
  const bool ag = ag_iter > 0;
  if (ag) leaf_size = kAgBlock;  // the unit the chains are cut into; leaves are groups inside a block
  RJB_REQUIRE(m.n_chains < (1u << 28), "too many chains for the leaf record (2^28)");
  // `built` stays false until the last call below has succeeded: a failed allocation or
  // launch must not leave an index that passes the checks of rjb_lsi / rjb_pip
  b.built = false;
  b.have_cells = false;
  b.leaf_size = leaf_size;
  b.n_leaves = 0;
  b.root_box = empty_box();
  if (m.n_edges == 0) {
    b.built = true;
    return;
  }
  const int T = 256;
  uint32_t* cnt = b.chain_cnt.ensure(m.n_chains + 1);
  uint32_t* base = b.leaf_base.ensure(m.n_chains + 1);
  k_chain_leaf_counts<<<div_up(m.n_chains, T), T, 0, st>>>(m.row_index, m.n_chains, leaf_size, cnt);
  exclusive_scan_u32(cnt, base, m.n_chains, b.scan_tmp, st);
  uint32_t n = 0;
  RJB_CUDA(cudaMemcpyAsync(&n, base + m.n_chains, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  RJB_CUDA(cudaStreamSynchronize(st));
  uint32_t n_blocks = 0;
  uint32_t* ag_packed = nullptr;
  uint32_t* ag_off = nullptr;
  if (ag) {
    // n = blocks of <= 8 edges; group each block, then count the leaves
    n_blocks = n;
    ag_packed = b.ag_packed.ensure(n_blocks + 1);
    uint32_t* ag_cnt = b.ag_cnt.ensure(n_blocks + 1);
    ag_off = b.ag_off.ensure(n_blocks + 1);
    k_ag_group<<<div_up(n_blocks, T), T, 0, st>>>(m, base, n_blocks, std::min(ag_iter, 3), enlarge, ag_packed, ag_cnt);
    exclusive_scan_u32(ag_cnt, ag_off, n_blocks, b.scan_tmp, st);
    RJB_CUDA(cudaMemcpyAsync(&n, ag_off + n_blocks, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RJB_CUDA(cudaStreamSynchronize(st));
  }
  b.n_leaves = n;
  uint32_t n_int = n > 1 ? n - 1 : 1;
  uint2* rec_u = b.leaf_rec_u.ensure(n);
  int4* box_u = b.leaf_box_u.ensure(n);
  int4* box_s = b.leaf_box_s.ensure(n);
  uint64_t* ka = b.keys_a.ensure(n);
  uint64_t* kb = b.keys_b.ensure(n);
  uint32_t* va = b.vals_a.ensure(n);  // scratch of k_occ_mark (big-box list)
  uint2* rec_s = b.leaf_rec.ensure(n);
  int4* nbox = b.node_box.ensure(2 * (size_t) n_int);
  int2* nchild = b.node_child.ensure(n_int);
  uint32_t* parent = b.parent.ensure(2 * (size_t) n);
  int4* root_d = b.root_box_d.ensure(1);

  if (ag)
    k_leaf_fill_ag<<<div_up(n_blocks, T), T, 0, st>>>(m, base, n_blocks, ag_packed, ag_off, imin, rec_u, box_u, ka);
  else
    k_leaf_fill<<<div_up(n, T), T, 0, st>>>(m, base, n, leaf_size, imin, rec_u, box_u, ka);
  const uint64_t* sorted = sort_packed(ka, kb, n, 0, 64 - kMortonDrop, b.sort_tmp, st);
  k_leaf_gather<<<div_up(n, T), T, 0, st>>>(sorted, n, rec_u, box_u, rec_s, box_s);
  if (n == 1) {
    k_single_leaf_root<<<1, 1, 0, st>>>(box_s, nbox, nchild, root_d);
  } else {
    k_refit_init<<<div_up(n_int, T), T, 0, st>>>(nbox, n_int);
    k_karras<<<div_up(n - 1, T), T, 0, st>>>(sorted, n, nchild, parent);
    k_refit<<<div_up(n, T), T, 0, st>>>(box_s, n, nchild, parent, nbox, root_d);
  }
  b.top_levels = n >= (1u << 18) ? 4 : 3;
  const int top_slots = b.top_levels == 4 ? kTopSlots4 : kTopSlots3;
  int4* tbox = b.top_box.ensure(top_slots);
  int* tcode = b.top_code.ensure(top_slots);
  k_top_tree<<<(1u << (5 * b.top_levels)) / 256, 256, 0, st>>>(nbox, nchild, root_d, n,
                                                              5 * b.top_levels, tbox, tcode);
  const uint32_t occ_words = kOccWords;
  uint32_t* occ = b.occ.ensure((size_t) kOccMaps * occ_words);  // occ, then the dilated occ2, occ4, occ8
  RJB_CUDA(cudaMemsetAsync(occ, 0, occ_words * sizeof(uint32_t), st));
  // [0] (leaf, cell) incidences, [1] cells of the largest box (> 1024 only), [2] big boxes
  unsigned long long* inc = b.inc_counter.ensure(3);
  RJB_CUDA(cudaMemsetAsync(inc, 0, 3 * sizeof(unsigned long long), st));
  k_occ_mark<<<div_up(n, T), T, 0, st>>>(box_s, n, occ, inc, va /* free after the sort */);
  k_occ_mark_big<<<kNumSMs, 256, 0, st>>>(box_s, va, inc, occ);
  for (int k = 1; k < kOccMaps; k++)
    k_occ_dilate<<<div_up(occ_words, T), T, 0, st>>>(occ + (size_t) (k - 1) * occ_words, occ + (size_t) k * occ_words,
                                                     1 << (k - 1));
  RJB_CUDA(cudaGetLastError());
  uint32_t n_occ = 0;
  unsigned long long n_inc = 0, max_area = 0;
  uint32_t* rank = nullptr;
  RJB_CUDA(cudaMemcpyAsync(&b.root_box, root_d, sizeof(int4), cudaMemcpyDeviceToHost, st));
  if (want_cells) {
    // rank of every bitmap word = number of occupied cells before it
    uint32_t* pop = b.occ_pop.ensure(occ_words);
    rank = b.occ_rank.ensure(occ_words + 1);
    k_occ_popc<<<div_up(occ_words, T), T, 0, st>>>(occ, pop);
    exclusive_scan_u32(pop, rank, occ_words, b.scan_tmp, st);
    RJB_CUDA(cudaMemcpyAsync(&n_occ, rank + occ_words, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RJB_CUDA(cudaMemcpyAsync(&n_inc, inc, sizeof(n_inc), cudaMemcpyDeviceToHost, st));
    RJB_CUDA(cudaMemcpyAsync(&max_area, inc + 1, sizeof(max_area), cudaMemcpyDeviceToHost, st));
  }
  RJB_CUDA(cudaStreamSynchronize(st));
  // share of occupied cells (exact with the directory, else an upper estimate: a leaf
  // touches ~1.5 cells): it steers the first query, the filter switches itself off when it
  // does not pay
  b.occ_fraction = want_cells ? (double) n_occ / ((double) kOccDim * kOccDim)
                              : std::min(1.0, 1.5 * (double) n / ((double) kOccDim * kOccDim));
  // Cell directory: only for sparse maps (the ones the occupancy filter is used for), and
  // only when the leaves are small against the cells (else the lists explode)
  b.have_cells = false;
  b.n_occ_cells = n_occ;
  b.n_incidences = 0;
  // (k_cell_lists walks a box in one thread: no directory when some box spans > 1024 cells)
  if (want_cells && b.occ_fraction < 0.25 && n_inc <= 8ull * n + 1024 && max_area <= 1024) {
    b.n_incidences = (uint32_t) n_inc;
    uint32_t* ccnt = b.cell_cnt.ensure(n_occ + 1);
    uint32_t* cbeg = b.cell_begin.ensure(n_occ + 1);
    uint4* clist = b.cell_item.ensure(n_inc ? n_inc : 1);
    uint2* dir = b.occ_dir.ensure(occ_words + 1);
    RJB_CUDA(cudaMemsetAsync(ccnt, 0, (n_occ + 1) * sizeof(uint32_t), st));
    k_cell_lists<false><<<div_up(n, T), T, 0, st>>>(box_s, rec_s, n, occ, rank, ccnt, nullptr, nullptr);
    exclusive_scan_u32(ccnt, cbeg, n_occ, b.scan_tmp, st);
    RJB_CUDA(cudaMemsetAsync(ccnt, 0, (n_occ + 1) * sizeof(uint32_t), st));
    k_cell_lists<true><<<div_up(n, T), T, 0, st>>>(box_s, rec_s, n, occ, rank, ccnt, cbeg, clist);
    k_occ_dir<<<div_up(occ_words + 1, T), T, 0, st>>>(occ, rank, dir);
    RJB_CUDA(cudaGetLastError());
    b.have_cells = true;
  }
  b.built = true;
}

}  // namespace rjb
