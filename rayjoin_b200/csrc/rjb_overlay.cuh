// Polygon overlay: MapOverlay protocol on top of the LSI / PIP kernels, and the
// host writer of the output chains.
//
// Replaces MapOverlay{LBVH,Grid}::{Init,BuildIndex,IntersectEdge,
// LocateVerticesInOtherMap,ComputeOutputPolygons,WriteResult}
// (reference: src/app/map_overlay_lbvh.h:25-270, src/run_overlay.cu:196-226)
// and WriteOutputChain (src/app/output_chain.h:41-205).
//
// Included by rjb_api.cu after rjb_ctx and the do_* helpers are defined.
#pragma once
#include <stdio.h>

#include <algorithm>
#include <map>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace rjb {

// packed word: eid[im] (high 32 bits) | queue position (low 32 bits)
__global__ void k_ov_keys(const rjb_xsect* __restrict__ xs, uint32_t n, int im, uint64_t* __restrict__ key) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  key[i] = ((uint64_t) xs[i].eid[im] << 32) | i;
}

__global__ void k_ov_gather(const rjb_xsect* __restrict__ xs, const uint64_t* __restrict__ order,
                            uint32_t n, int im, rjb_xsect* __restrict__ out,
                            uint32_t* __restrict__ seg_flag) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rjb_xsect x = xs[(uint32_t) order[i]];
  out[i] = x;
  uint32_t prev = i ? (uint32_t) (order[i - 1] >> 32) : 0xFFFFFFFFu;
  seg_flag[i] = (i == 0 || prev != x.eid[im]) ? 1u : 0u;
}

__global__ void k_ov_seg_starts(const uint32_t* __restrict__ seg_flag,
                                const uint32_t* __restrict__ seg_scan, uint32_t n,
                                uint32_t* __restrict__ seg_start) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (seg_flag[i]) seg_start[seg_scan[i]] = i;
  if (i == n - 1) seg_start[seg_scan[i] + seg_flag[i]] = n;
}

// per intersected edge of map im: order its intersections along the edge
// (squared distance to p1 in exact integers, src/app/map_overlay_lbvh.h:204-214;
// ties by the other map's eid for determinism) and emit the integer mid-points
// of consecutive intersections (:216-228): trunc((x1 + x2) / 2).
__global__ void k_ov_sort_segments(MapView M, int im, rjb_xsect* __restrict__ xs,
                                   const uint32_t* __restrict__ seg_start, uint32_t n_segs,
                                   longlong2* __restrict__ mid_pts) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  uint32_t b = seg_start[s], e = seg_start[s + 1];
  uint32_t n = e - b;
  if (n < 2) return;
  uint32_t eid = xs[b].eid[im];
  longlong2 p1 = M.pts[eid + M.edge_chain[eid]];
  auto dist = [&](const rjb_xsect& x) {
    i128 dx = (i128) x.x - p1.x, dy = (i128) x.y - p1.y;
    return dx * dx + dy * dy;
  };
  // insertion sort (segments are short)
  for (uint32_t i = b + 1; i < e; i++) {
    rjb_xsect cur = xs[i];
    i128 dc = dist(cur);
    uint32_t j = i;
    while (j > b) {
      rjb_xsect pv = xs[j - 1];
      i128 dp = dist(pv);
      bool less = dc < dp || (dc == dp && cur.eid[1 - im] < pv.eid[1 - im]);
      if (!less) break;
      xs[j] = pv;
      j--;
    }
    xs[j] = cur;
  }
  for (uint32_t k = 0; k + 1 < n; k++) {
    const rjb_xsect& a = xs[b + k];
    const rjb_xsect& c = xs[b + k + 1];
    longlong2 m;
    m.x = (a.x + c.x) / 2;  // C division truncates toward zero like the
    m.y = (a.y + c.y) / 2;  // rational -> double -> int64 path of the reference
    mid_pts[b + k - s] = m;
  }
}

__global__ void k_ov_fill_mid_faces(rjb_xsect* __restrict__ xs, const uint32_t* __restrict__ seg_start,
                                    uint32_t n_segs, const int32_t* __restrict__ mid_face) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_segs) return;
  uint32_t b = seg_start[s], e = seg_start[s + 1];
  for (uint32_t k = 0; b + k + 1 < e; k++) xs[b + k].mid_point_polygon_id = mid_face[b + k - s];
}

// Results computed elsewhere (the multi-GPU driver: LSI and vertex location are
// sharded over the ranks and gathered on one) handed to overlay_run instead of
// running IntersectEdge / LocateVerticesInOtherMap here.
struct OverlayImport {
  bool device = false;  // the arrays below are DEVICE pointers (gathered over NCCL), else host
  const rjb_xsect* h_xsects = nullptr;
  uint64_t n_xsects = 0;
  const uint32_t* h_closest_eid[2] = {nullptr, nullptr};
  const int32_t* h_point_in_polygon[2] = {nullptr, nullptr};
};

static void overlay_run(rjb_ctx* c, int mode, uint32_t grid_size, double xsect_factor,
                        double* phase_ms, const OverlayImport* imp = nullptr) {
  RJB_REQUIRE(c->maps[0].loaded && c->maps[1].loaded, "rjb_overlay_run: load both maps first");
  RJB_REQUIRE(mode == RJB_MODE_LBVH || mode == RJB_MODE_GRID || mode == RJB_MODE_BRUTE,
              "rjb_overlay_run: unknown mode");
  OverlayState& ov = c->ov;
  ov.done = false;
  cudaStream_t st = c->stream;
  cudaEvent_t ev[7];
  for (auto& e : ev) RJB_CUDA(cudaEventCreate(&e));
  auto mark = [&](int i) { RJB_CUDA(cudaEventRecord(ev[i], st)); };
  mark(0);
  // BuildIndex: one index per map, both directions are queried.  (The multi-GPU finish on a
  // rank that ran the sharded phases reuses its resident maps and the indexes built for them.)
  for (int im = 0; im < 2; im++) {
    DeviceMap& m = c->maps[im];
    const bool have = mode == RJB_MODE_LBVH ? m.bvh.built
                                            : (mode == RJB_MODE_GRID ? m.grid.built && m.grid.gsize == grid_size : true);
    if (!(imp && have)) do_build_index(c, im, mode, grid_size, nullptr);
  }
  mark(1);
  uint64_t n = 0;
  if (imp) {
    // gathered results of the sharded phases
    n = imp->n_xsects;
    rjb_xsect* xs = c->xsects.ensure(n ? n : 1);
    const cudaMemcpyKind kind = imp->device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (n) RJB_CUDA(cudaMemcpyAsync(xs, imp->h_xsects, n * sizeof(rjb_xsect), kind, st));
    mark(2);
    for (int im = 0; im < 2; im++) {
      DeviceMap& Qm = c->maps[im];
      uint32_t* ce = ov.closest_eid[im].ensure(Qm.n_points ? Qm.n_points : 1);
      int32_t* pf = ov.point_in_polygon[im].ensure(Qm.n_points ? Qm.n_points : 1);
      RJB_CUDA(cudaMemcpyAsync(ce, imp->h_closest_eid[im], Qm.n_points * sizeof(uint32_t), kind, st));
      RJB_CUDA(cudaMemcpyAsync(pf, imp->h_point_in_polygon[im], Qm.n_points * sizeof(int32_t), kind, st));
      mark(3 + im);
    }
  } else {
    // IntersectEdge(0): map 0 is the query side (src/run_overlay.cu:206)
    n = do_lsi(c, 0, mode, xsect_factor, nullptr);
    mark(2);
    // LocateVerticesInOtherMap(im)
    for (int im = 0; im < 2; im++) {
      DeviceMap& Qm = c->maps[im];
      do_pip(c, im, mode, Qm.pts.p, Qm.n_points, nullptr);
      uint32_t* ce = ov.closest_eid[im].ensure(Qm.n_points ? Qm.n_points : 1);
      int32_t* pf = ov.point_in_polygon[im].ensure(Qm.n_points ? Qm.n_points : 1);
      RJB_CUDA(cudaMemcpyAsync(ce, c->pip_eid.p, Qm.n_points * sizeof(uint32_t),
                               cudaMemcpyDeviceToDevice, st));
      RJB_CUDA(cudaMemcpyAsync(pf, c->pip_face.p, Qm.n_points * sizeof(int32_t),
                               cudaMemcpyDeviceToDevice, st));
      mark(3 + im);
    }
  }
  ov.n_xsects = n;
  // ComputeOutputPolygons
  for (int im = 0; im < 2; im++) {
    rjb_xsect* sorted = ov.xsects_sorted[im].ensure(n ? n : 1);
    ov.n_segs[im] = 0;
    if (n == 0) continue;
    uint32_t n32 = (uint32_t) n;
    uint64_t* ka = ov.keys_a.ensure(n);
    uint64_t* kb = ov.keys_b.ensure(n);
    uint32_t* flag = ov.seg_flag.ensure(n + 1);
    uint32_t* scan = ov.seg_scan.ensure(n + 1);
    k_ov_keys<<<div_up(n, 256), 256, 0, st>>>(c->xsects.p, n32, im, ka);
    const uint64_t* order = sort_packed(ka, kb, n32, 0, 32, ov.sort_tmp, st);
    k_ov_gather<<<div_up(n, 256), 256, 0, st>>>(c->xsects.p, order, n32, im, sorted, flag);
    exclusive_scan_u32(flag, scan, n32, ov.scan_tmp, st);
    uint32_t n_segs = 0;
    RJB_CUDA(cudaMemcpyAsync(&n_segs, scan + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    RJB_CUDA(cudaStreamSynchronize(st));
    ov.n_segs[im] = n_segs;
    uint32_t* seg_start = ov.seg_start[im].ensure(n_segs + 1);
    k_ov_seg_starts<<<div_up(n, 256), 256, 0, st>>>(flag, scan, n32, seg_start);
    uint32_t n_mid = n32 - n_segs;
    longlong2* mid = ov.mid_pts.ensure(n_mid ? n_mid : 1);
    k_ov_sort_segments<<<div_up(n_segs, 128), 128, 0, st>>>(c->maps[im].view(), im, sorted,
                                                            seg_start, n_segs, mid);
    RJB_CUDA(cudaGetLastError());
    // mid-points are located in the OTHER map with query_map_id = im
    // (src/app/map_overlay_lbvh.h:233-236)
    do_pip(c, im, mode, mid, n_mid, nullptr);
    k_ov_fill_mid_faces<<<div_up(n_segs, 128), 128, 0, st>>>(sorted, seg_start, n_segs,
                                                             c->pip_face.p);
    RJB_CUDA(cudaGetLastError());
  }
  mark(5);
  RJB_CUDA(cudaStreamSynchronize(st));
  if (phase_ms) {
    float ms;
    for (int i = 0; i < 5; i++) {
      RJB_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      phase_ms[i] = ms;
    }
    RJB_CUDA(cudaEventElapsedTime(&ms, ev[0], ev[5]));
    phase_ms[5] = ms;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  ov.done = true;
}

// ---------------------------------------------------------------------------
// host writer (WriteOutputChain is a single-threaded
// unordered_map of every output point plus one ostream << per coordinate; for the
// County x Zipcode-scale result -- 17 M points, 285 MB -- that took 12.7 s).
// Same semantics, different mechanics: one flat point array instead of a vector
// per chain, an open-addressing table with prefetch for the first-appearance point
// ids, an exact fixed-point "%.6f" formatter, formatting on all host threads.
// ---------------------------------------------------------------------------
struct P2 {
  double x, y;
  bool operator==(const P2& o) const { return x == o.x && y == o.y; }
};

struct OutPiece {
  uint64_t begin, end;  // range in the flat point array
  int64_t left, right, other;
  uint32_t first_pid, last_pid;
};

// "%.6f" of a finite double, exactly as printf rounds it (round-half-even on the
// exact binary value): x = M * 2^E, N = round(M * 10^6 / 2^-E) in 128-bit integers.
static inline char* fmt_fixed6(double v, char* p) {
  uint64_t bits;
  memcpy(&bits, &v, 8);
  const int ebits = (int) ((bits >> 52) & 0x7FF);
  uint64_t M = bits & ((1ull << 52) - 1);
  if (ebits == 0x7FF || ebits >= 1023 + 43) return p + sprintf(p, "%.6f", v);  // inf, nan, >= 2^43
  if (bits >> 63) *p++ = '-';
  int E;
  if (ebits == 0) E = -1074; else { M |= 1ull << 52; E = ebits - 1075; }
  unsigned __int128 P = (unsigned __int128) M * 1000000u;  // < 2^73
  uint64_t q;
  if (E >= 0) {
    q = (uint64_t) (P << E);  // unreachable below 2^43 (E <= -10); kept for completeness
  } else {
    const int sh = -E;
    if (sh >= 100) q = 0;  // < 2^-27 * 10^6: rounds to zero
    else {
      unsigned __int128 qq = P >> sh, rem = P & (((unsigned __int128) 1 << sh) - 1);
      const unsigned __int128 half = (unsigned __int128) 1 << (sh - 1);
      if (rem > half || (rem == half && (qq & 1))) qq++;
      q = (uint64_t) qq;
    }
  }
  const uint64_t ip = q / 1000000u;
  uint32_t fp = (uint32_t) (q % 1000000u);
  char tmp[24];
  int n = 0;
  uint64_t t = ip;
  do { tmp[n++] = (char) ('0' + t % 10); t /= 10; } while (t);
  while (n) *p++ = tmp[--n];
  *p++ = '.';
  for (int i = 5; i >= 0; i--) { p[i] = (char) ('0' + fp % 10); fp /= 10; }
  return p + 6;
}

static inline uint64_t hash_p2(const P2& p) {
  uint64_t a, b;
  memcpy(&a, &p.x, 8);
  memcpy(&b, &p.y, 8);
  uint64_t h = a * 0x9E3779B97F4A7C15ull ^ (b + 0x7F4A7C15ull + (a << 6) + (a >> 2));
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  return h ^ (h >> 32);
}

static void overlay_write(rjb_ctx* c, const char* path) {
  OverlayState& ov = c->ov;
  RJB_REQUIRE(ov.done, "rjb_overlay_write: run rjb_overlay_run first");
  for (int im = 0; im < 2; im++)
    RJB_REQUIRE(c->maps[im].h_xy.size() == 2 * (size_t) c->maps[im].n_points,
                "rjb_overlay_write: maps were loaded with keep_host_graph=0");
  const rjb_scaling& sc = c->sc;
  std::vector<P2> pts;          // all output points, piece after piece
  std::vector<OutPiece> out;
  pts.reserve((size_t) c->maps[0].n_points + c->maps[1].n_points + 4 * ov.n_xsects + 16);
  uint64_t cur_begin = 0;
  int64_t cur_left = 0, cur_right = 0, cur_other = 0;
  // flush(): keep a piece iff one of its own faces and the other map's face are
  // both non-exterior; consecutive equal points collapse (output_chain.h:56-76)
  auto flush = [&]() {
    if (pts.size() == cur_begin) return;
    if (cur_left * cur_other != 0 || cur_right * cur_other != 0) {
      uint64_t w = cur_begin + 1;
      for (uint64_t r = cur_begin + 1; r < pts.size(); r++)
        if (!(pts[r] == pts[w - 1])) pts[w++] = pts[r];
      pts.resize(w);
      OutPiece pc = {cur_begin, w, cur_left, cur_right, cur_other, 0, 0};
      out.push_back(pc);
      cur_begin = w;
    } else {
      pts.resize(cur_begin);
    }
  };
  // host Unscale (src/map/scaling.h:100-106) without FMA contraction
  auto unscale = [&](const rjb_xsect& x) {
    volatile double tx = (double) x.x * sc.rrx, ty = (double) x.y * sc.rry;
    P2 p = {tx + sc.ddeltax, ty + sc.ddeltay};
    return p;
  };
  for (int im = 0; im < 2; im++) {
    DeviceMap& m = c->maps[im];
    uint64_t n = ov.n_xsects;
    std::vector<rjb_xsect> xs(n);
    std::vector<int32_t> pip(m.n_points);
    if (n)
      RJB_CUDA(cudaMemcpy(xs.data(), ov.xsects_sorted[im].p, n * sizeof(rjb_xsect),
                          cudaMemcpyDeviceToHost));
    if (m.n_points)
      RJB_CUDA(cudaMemcpy(pip.data(), ov.point_in_polygon[im].p, m.n_points * sizeof(int32_t),
                          cudaMemcpyDeviceToHost));
    // xs is sorted by eid[im]: CSR over the edges of this map
    std::vector<uint32_t> first(m.n_edges + 2, 0);
    for (uint64_t i = 0; i < n; i++) {
      RJB_REQUIRE(xs[i].eid[im] < m.n_edges, "rjb_overlay_write: result refers to an edge the map does not have");
      first[xs[i].eid[im] + 1]++;
    }
    for (uint32_t e = 0; e < m.n_edges; e++) first[e + 1] += first[e];
    for (uint32_t ic = 0; ic < m.n_chains; ic++) {
      uint32_t pb = m.h_row_index[ic], pe = m.h_row_index[ic + 1];
      cur_begin = pts.size();
      cur_left = m.h_left[ic];
      cur_right = m.h_right[ic];
      for (uint32_t pid = pb; pid < pe; pid++) {
        cur_other = pip[pid];
        P2 p = {m.h_xy[2 * (size_t) pid], m.h_xy[2 * (size_t) pid + 1]};
        pts.push_back(p);
        if (pid != pe - 1) {
          const uint32_t eid = pid - ic;
          const uint32_t b = first[eid], e = first[eid + 1];
          if (b != e) {
            pts.push_back(unscale(xs[b]));
            for (uint32_t k = b; k + 1 < e; k++) {
              flush();
              cur_other = xs[k].mid_point_polygon_id;
              pts.push_back(unscale(xs[k]));
              pts.push_back(unscale(xs[k + 1]));
            }
            flush();
            pts.push_back(unscale(xs[e - 1]));
          }
        }
      }
      flush();
    }
  }
  // face pairs by first appearance (output_chain.h:141-174)
  std::map<std::pair<int64_t, int64_t>, size_t> face_ids;
  auto create_polygon = [&](int64_t a, int64_t b) -> size_t {
    if (a == 0 || b == 0) return 0;
    auto k = std::make_pair(a, b);
    auto it = face_ids.find(k);
    if (it == face_ids.end()) {
      size_t id = face_ids.size() + 1;
      face_ids[k] = id;
      return id;
    }
    return it->second;
  };
  for (auto& ch : out) {
    ch.left = ch.left < ch.other ? (int64_t) create_polygon(ch.left, ch.other)
                                 : (int64_t) create_polygon(ch.other, ch.left);
    ch.right = ch.right < ch.other ? (int64_t) create_polygon(ch.right, ch.other)
                                   : (int64_t) create_polygon(ch.other, ch.right);
  }
  // point ids by first appearance of the exact coordinates (output_chain.h:176-182):
  // open addressing, slot = {32-bit tag, 32-bit id}, keys compared through the id's
  // representative point; the probe targets of a block are prefetched ahead
  {
    const uint64_t np = pts.size();
    RJB_REQUIRE(np < 0xFFFFFFF0ull, "rjb_overlay_write: too many output points");
    uint64_t cap = 64;
    while (cap < 2 * np) cap <<= 1;
    std::vector<uint64_t> table(cap, ~0ull);
    std::vector<uint32_t> rep;  // id -> index of its first occurrence
    rep.reserve(np / 2 + 16);
    std::vector<uint32_t> pid(np);
    const uint64_t mask = cap - 1;
    const uint64_t B = 32;
    uint64_t hs[B];
    for (uint64_t i0 = 0; i0 < np; i0 += B) {
      const uint64_t n = std::min(B, np - i0);
      for (uint64_t k = 0; k < n; k++) {
        hs[k] = hash_p2(pts[i0 + k]);
        __builtin_prefetch(&table[hs[k] & mask]);
      }
      for (uint64_t k = 0; k < n; k++) {
        const P2& p = pts[i0 + k];
        const uint64_t tag = hs[k] >> 32;
        uint64_t s = hs[k] & mask;
        while (true) {
          const uint64_t v = table[s];
          if (v == ~0ull) {
            const uint32_t id = (uint32_t) rep.size();
            rep.push_back((uint32_t) (i0 + k));
            table[s] = (tag << 32) | id;
            pid[i0 + k] = id;
            break;
          }
          if ((v >> 32) == tag && pts[rep[(uint32_t) v]] == p) {
            pid[i0 + k] = (uint32_t) v;
            break;
          }
          s = (s + 1) & mask;
        }
      }
    }
    for (auto& ch : out) {
      ch.first_pid = pid[ch.begin];
      ch.last_pid = pid[ch.end - 1];
    }
  }
  // format on all host threads, write in order
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 32) nt = 32;
  if (out.size() < 4096) nt = 1;
  std::vector<std::string> text(nt);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
      th.emplace_back([&, t] {
        const size_t c0 = out.size() * t / nt, c1 = out.size() * (t + 1) / nt;
        size_t npts = 0;
        for (size_t i = c0; i < c1; i++) npts += out[i].end - out[i].begin;
        std::string& s = text[t];
        s.resize(npts * 64 + (c1 - c0) * 128 + 64);
        char* p = &s[0];
        for (size_t i = c0; i < c1; i++) {
          const OutPiece& ch = out[i];
          p += sprintf(p, "%zu %llu %u %u %lld %lld\n", i + 1, (unsigned long long) (ch.end - ch.begin),
                       ch.first_pid, ch.last_pid, (long long) ch.left, (long long) ch.right);
          for (uint64_t k = ch.begin; k < ch.end; k++) {
            p = fmt_fixed6(pts[k].x, p);
            *p++ = ' ';
            p = fmt_fixed6(pts[k].y, p);
            *p++ = '\n';
          }
        }
        s.resize((size_t) (p - &s[0]));
      });
    for (auto& t : th) t.join();
  }
  FILE* f = fopen(path, "w");
  if (!f) throw Error(RJB_ERR_IO, std::string("Cannot open ") + path);
  bool ok = true;
  for (auto& s : text) ok = ok && (s.empty() || fwrite(s.data(), 1, s.size(), f) == s.size());
  ok = (fclose(f) == 0) && ok;
  if (!ok) throw Error(RJB_ERR_IO, std::string("Write failed: ") + path);
}

}  // namespace rjb
