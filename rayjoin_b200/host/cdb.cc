// CDB chain files: text parser, .bin cache, load_from.  Host C++ only.
//
// Behaviour-compatible with the reference loader
// (src/map/planar_graph.h:41-126 read_pgraph, :128-167 serialize_pgraph,
//  :169-220 deserialize_pgraph, :222-252 load_from):
//   * lines that are empty or start with '#' or '%' are skipped (:58);
//   * a chain header is `id np first last left right`, followed by np lines
//     `x y`; np < 2 and two consecutive identical points are errors (:71,85);
//   * row_index gets one entry per chain plus a final one unless the graph is
//     empty (:102-104); the bounding box covers every point (:88-91);
//   * <prefix>/<path with '/' -> '-'>.bin is read when readable, else the text
//     is parsed and the .bin written when the prefix is writable (:222-252);
//   * .bin layout: u64 0xabcdabcd, n_chains, n_row_index, n_points; per chain
//     5 x i64; row_index as u32; points as double x,y; bbox min_x,min_y,max_x,
//     max_y; u64 0xabcdabcd (:129-167).
// The parser itself is new (the reference's
// istringstream-per-line loop takes 0.7-2.4 s on the real maps and dominates the
// end-to-end wall time).  Fast path: the file is split at line boundaries over the
// host threads; a line with two tokens is a vertex, a line with six is a chain
// header, numbers go through std::from_chars (correctly rounded, like strtod); the
// pieces are stitched and validated (np, duplicates).  Anything unusual -- other
// token counts, '+' signs, hex floats, a count that does not add up -- falls back
// to the sequential strtod/strtoll parser, which reproduces the reference's
// error behaviour line by line.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <charconv>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "rjb200.h"

namespace {

// vector<double> whose resize() does not touch the memory: the pages of a 150 MB vertex
// array are first touched by the threads that fill it, not zero-filled by one thread first
template <class T>
struct NoInitAlloc : std::allocator<T> {
  template <class U>
  struct rebind {
    using other = NoInitAlloc<U>;
  };
  template <class U>
  void construct(U* p) noexcept {
    ::new (static_cast<void*>(p)) U;  // default-init: nothing for double
  }
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
  }
};
using DVec = std::vector<double, NoInitAlloc<double>>;

struct GraphOwner {
  std::vector<int64_t> chain_id, first_point, last_point, left, right;
  std::vector<uint32_t> row_index;
  DVec xy;
  double min_x = std::numeric_limits<double>::max();
  double min_y = std::numeric_limits<double>::max();
  double max_x = -std::numeric_limits<double>::max();
  double max_y = -std::numeric_limits<double>::max();
};

}  // namespace
extern "C" void rjb__set_error(const char* msg);  // rjb_api.cu
namespace {

void publish(GraphOwner* o, rjb_graph* g) {
  g->n_chains = o->chain_id.size();
  g->n_points = o->xy.size() / 2;
  g->chain_id = o->chain_id.data();
  g->first_point = o->first_point.data();
  g->last_point = o->last_point.data();
  g->left = o->left.data();
  g->right = o->right.data();
  g->row_index = o->row_index.data();
  g->xy = o->xy.data();
  g->min_x = o->min_x;
  g->min_y = o->min_y;
  g->max_x = o->max_x;
  g->max_y = o->max_y;
  g->_owner = o;
}

int fail(const std::string& msg) {
  rjb__set_error(msg.c_str());
  return RJB_ERR_IO;
}

// ---- parallel fast path -----------------------------------------------------
struct Piece {
  DVec xy;
  std::vector<int64_t> hdr;       // 6 per header
  std::vector<uint64_t> hdr_pos;  // vertices of this piece seen before the header
  double min_x = std::numeric_limits<double>::max(), min_y = min_x;
  double max_x = -std::numeric_limits<double>::max(), max_y = max_x;
  bool ok = true;
};

inline const char* skip_ws(const char* p, const char* e) {
  while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
  return p;
}

void parse_piece(const char* p, const char* end, Piece* out) {
  out->xy.reserve((size_t) (end - p) / 12);
  while (p < end) {
    const char* eol = (const char*) memchr(p, '\n', (size_t) (end - p));
    if (!eol) eol = end;
    const char* q = p;
    p = eol + 1;
    if (q == eol || *q == '#' || *q == '%') continue;  // like the reference: first char only
    // count tokens
    int ntok = 0;
    const char* t = q;
    const char* tok[7];
    while (true) {
      t = skip_ws(t, eol);
      if (t >= eol) break;
      if (ntok < 7) tok[ntok] = t;
      ntok++;
      while (t < eol && *t != ' ' && *t != '\t' && *t != '\r') t++;
    }
    if (ntok == 2) {
      double x, y;
      const char* e0 = tok[0];
      while (e0 < eol && *e0 != ' ' && *e0 != '\t' && *e0 != '\r') e0++;
      const char* e1 = tok[1];
      while (e1 < eol && *e1 != ' ' && *e1 != '\t' && *e1 != '\r') e1++;
      auto r0 = std::from_chars(tok[0], e0, x);
      auto r1 = std::from_chars(tok[1], e1, y);
      if (r0.ec != std::errc() || r0.ptr != e0 || r1.ec != std::errc() || r1.ptr != e1) {
        out->ok = false;
        return;
      }
      out->xy.push_back(x);
      out->xy.push_back(y);
      if (x < out->min_x) out->min_x = x;
      if (x > out->max_x) out->max_x = x;
      if (y < out->min_y) out->min_y = y;
      if (y > out->max_y) out->max_y = y;
    } else if (ntok == 6) {
      for (int i = 0; i < 6; i++) {
        const char* e0 = tok[i];
        while (e0 < eol && *e0 != ' ' && *e0 != '\t' && *e0 != '\r') e0++;
        long long v;
        auto r = std::from_chars(tok[i], e0, v);
        if (r.ec != std::errc() || r.ptr != e0) {
          out->ok = false;
          return;
        }
        out->hdr.push_back(v);
      }
      out->hdr_pos.push_back(out->xy.size() / 2);
    } else {
      out->ok = false;  // blank-but-not-empty line, extra tokens, ...: let the slow path decide
      return;
    }
  }
}

// returns true when the fast path produced a valid graph in *o
bool read_text_parallel(const char* buf, size_t len, GraphOwner* o) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 32) nt = 32;
  if (len < (1u << 20)) nt = 1;
  std::vector<size_t> cut(nt + 1, len);
  cut[0] = 0;
  for (unsigned i = 1; i < nt; i++) {
    size_t c = len / nt * i;
    if (c < cut[i - 1]) c = cut[i - 1];
    const char* nl = (const char*) memchr(buf + c, '\n', len - c);
    cut[i] = nl ? (size_t) (nl - buf) + 1 : len;
  }
  std::vector<Piece> pieces(nt);
  std::vector<std::thread> th;
  for (unsigned i = 0; i < nt; i++)
    th.emplace_back(parse_piece, buf + cut[i], buf + cut[i + 1], &pieces[i]);
  for (auto& t : th) t.join();
  size_t n_pts = 0, n_hdr = 0;
  for (auto& pc : pieces) {
    if (!pc.ok) return false;
    n_pts += pc.xy.size() / 2;
    n_hdr += pc.hdr_pos.size();
  }
  if (n_pts >= 0xFFFFFFF0ull) return false;
  o->xy.resize(2 * n_pts);
  o->chain_id.resize(n_hdr);
  o->first_point.resize(n_hdr);
  o->last_point.resize(n_hdr);
  o->left.resize(n_hdr);
  o->right.resize(n_hdr);
  o->row_index.resize(n_hdr ? n_hdr + 1 : 0);
  std::vector<int64_t> np(n_hdr);
  size_t pbase = 0, hbase = 0;
  std::vector<std::thread> cp;
  for (auto& pc : pieces) {
    cp.emplace_back([&o, &pc, pbase] { memcpy(o->xy.data() + 2 * pbase, pc.xy.data(), pc.xy.size() * 8); });
    for (size_t h = 0; h < pc.hdr_pos.size(); h++) {
      const int64_t* f = &pc.hdr[6 * h];
      o->chain_id[hbase + h] = f[0];
      np[hbase + h] = f[1];
      o->first_point[hbase + h] = f[2];
      o->last_point[hbase + h] = f[3];
      o->left[hbase + h] = f[4];
      o->right[hbase + h] = f[5];
      o->row_index[hbase + h] = (uint32_t) (pbase + pc.hdr_pos[h]);
    }
    pbase += pc.xy.size() / 2;
    hbase += pc.hdr_pos.size();
    if (pc.min_x < o->min_x) o->min_x = pc.min_x;
    if (pc.max_x > o->max_x) o->max_x = pc.max_x;
    if (pc.min_y < o->min_y) o->min_y = pc.min_y;
    if (pc.max_y > o->max_y) o->max_y = pc.max_y;
  }
  for (auto& t : cp) t.join();
  if (n_hdr) o->row_index[n_hdr] = (uint32_t) n_pts;
  // validation: the first line must be a header, every chain has exactly np >= 2
  // vertices, no two consecutive vertices of a chain coincide
  if (n_pts && (n_hdr == 0 || o->row_index[0] != 0)) return false;
  for (size_t h = 0; h < n_hdr; h++) {
    const uint32_t b = o->row_index[h], e = o->row_index[h + 1];
    if (np[h] < 2 || (int64_t) (e - b) != np[h]) return false;
  }
  bool dup = false;
  {
    std::vector<std::thread> chk;
    std::vector<char> bad(nt, 0);
    for (unsigned i = 0; i < nt; i++)
      chk.emplace_back([&, i] {
        const size_t h0 = n_hdr * i / nt, h1 = n_hdr * (i + 1) / nt;
        for (size_t h = h0; h < h1 && !bad[i]; h++)
          for (uint32_t p = o->row_index[h] + 1; p < o->row_index[h + 1]; p++)
            if (o->xy[2 * (size_t) p] == o->xy[2 * (size_t) p - 2] &&
                o->xy[2 * (size_t) p + 1] == o->xy[2 * (size_t) p - 1]) {
              bad[i] = 1;
              break;
            }
      });
    for (auto& t : chk) t.join();
    for (char b : bad) dup |= (b != 0);
  }
  return !dup;
}

int read_text(const char* path, GraphOwner* o) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(std::string("Cannot open file ") + path);
  struct stat sb;
  if (fstat(fd, &sb) != 0) {
    close(fd);
    return fail(std::string("Cannot stat file ") + path);
  }
  size_t len = (size_t) sb.st_size;
  // Fast path straight from the page cache: no 250 MB buffer to allocate, zero and fill.
  if (len > 0 && getenv("RJB_CDB_SEQUENTIAL") == nullptr) {
    void* m = mmap(nullptr, len, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (m != MAP_FAILED) {
      const bool ok = read_text_parallel((const char*) m, len, o);
      munmap(m, len);
      if (ok) {
        close(fd);
        return RJB_OK;
      }
      *o = GraphOwner();  // anything unusual: the sequential parser decides and reports
    }
  }
  std::vector<char> buf(len + 1);
  size_t got = 0;
  while (got < len) {
    ssize_t r = read(fd, buf.data() + got, len - got);
    if (r <= 0) break;
    got += (size_t) r;
  }
  close(fd);
  if (got != len) return fail(std::string("Short read on ") + path);
  buf[len] = '\0';

  int64_t np = 0;
  bool have_last = false;
  double last_x = 0, last_y = 0;
  size_t lno = 0;
  char* p = buf.data();
  char* end = p + len;
  while (p < end) {
    char* eol = (char*) memchr(p, '\n', (size_t) (end - p));
    if (!eol) eol = end;
    lno++;
    char saved = *eol;
    *eol = '\0';
    char* line = p;
    p = eol + 1;
    size_t ll = (size_t) (eol - line);
    if (ll > 0 && line[ll - 1] == '\r') line[ll - 1] = '\0';
    if (line[0] == '\0' || line[0] == '#' || line[0] == '%') {
      *eol = saved;
      continue;
    }
    bool bad = false;
    char* q = line;
    if (np == 0) {
      long long v[6];
      for (int i = 0; i < 6 && !bad; i++) {
        char* e;
        errno = 0;
        v[i] = strtoll(q, &e, 10);
        if (e == q) bad = true;
        q = e;
      }
      if (!bad) {
        np = v[1];
        bad |= np < 2;
        o->chain_id.push_back(v[0]);
        o->first_point.push_back(v[2]);
        o->last_point.push_back(v[3]);
        o->left.push_back(v[4]);
        o->right.push_back(v[5]);
        o->row_index.push_back((uint32_t) (o->xy.size() / 2));
        have_last = false;
      }
    } else {
      char* e;
      double x = strtod(q, &e);
      if (e == q) bad = true;
      q = e;
      double y = strtod(q, &e);
      if (e == q) bad = true;
      if (!bad) {
        if (have_last) bad |= (x == last_x && y == last_y);
        if (x < o->min_x) o->min_x = x;
        if (x > o->max_x) o->max_x = x;
        if (y < o->min_y) o->min_y = y;
        if (y > o->max_y) o->max_y = y;
        o->xy.push_back(x);
        o->xy.push_back(y);
        last_x = x;
        last_y = y;
        have_last = true;
        np--;
      }
    }
    if (bad) {
      std::string l(line);
      return fail(std::string("Bad line. Check your dataset! ") + path + "[" + std::to_string(lno) +
                  "]: " + l);
    }
    *eol = saved;
  }
  if (!o->xy.empty()) o->row_index.push_back((uint32_t) (o->xy.size() / 2));
  if (np != 0) return fail(std::string("Truncated chain at end of ") + path);
  return RJB_OK;
}

template <typename T>
bool rd(FILE* f, T& v) {
  return fread(&v, sizeof(T), 1, f) == 1;
}

int read_bin(const char* path, GraphOwner* o) {
  FILE* f = fopen(path, "rb");
  if (!f) return fail(std::string("Cannot open file ") + path);
  uint64_t magic = 0, n_chains = 0, n_row = 0, n_points = 0;
  bool ok = rd(f, magic) && magic == 0xabcdabcdull && rd(f, n_chains) && rd(f, n_row) &&
            rd(f, n_points);
  if (ok) {
    // the counts come straight from the file: check them against each other and against the
    // file size before any vector is sized from them (a truncated or corrupt cache must give
    // RJB_ERR_IO, not std::bad_alloc)
    struct stat sb;
    ok = fstat(fileno(f), &sb) == 0;
    const uint64_t lim = 1ull << 40;
    ok = ok && n_chains < lim && n_row < lim && n_points < lim;
    ok = ok && n_row == (n_points ? n_chains + 1 : 0) && n_points < 0xFFFFFFF0ull;
    ok = ok && (uint64_t) sb.st_size == 8 * 4 + n_chains * 40 + n_row * 4 + n_points * 16 + 4 * 8 + 8;
  }
  if (ok) {
    o->chain_id.resize(n_chains);
    o->first_point.resize(n_chains);
    o->last_point.resize(n_chains);
    o->left.resize(n_chains);
    o->right.resize(n_chains);
    o->row_index.resize(n_row);
    o->xy.resize(2 * n_points);
    for (uint64_t i = 0; i < n_chains && ok; i++)
      ok = rd(f, o->chain_id[i]) && rd(f, o->first_point[i]) && rd(f, o->last_point[i]) &&
           rd(f, o->left[i]) && rd(f, o->right[i]);
    if (ok && n_row) ok = fread(o->row_index.data(), sizeof(uint32_t), n_row, f) == n_row;
    if (ok && n_points) ok = fread(o->xy.data(), 2 * sizeof(double), n_points, f) == n_points;
    ok = ok && rd(f, o->min_x) && rd(f, o->min_y) && rd(f, o->max_x) && rd(f, o->max_y);
    ok = ok && rd(f, magic) && magic == 0xabcdabcdull;
  }
  fclose(f);
  if (!ok) return fail(std::string("Corrupt .bin graph ") + path);
  return RJB_OK;
}

// no exception leaves the C ABI: allocation failures while parsing become RJB_ERR_IO
template <typename F>
int read_guarded(const char* path, rjb_graph* out, F&& reader) {
  if (!path || !out) return RJB_ERR_INVALID;
  GraphOwner* o = nullptr;
  try {
    o = new GraphOwner();
    int rc = reader(path, o);
    if (rc != RJB_OK) {
      delete o;
      return rc;
    }
    publish(o, out);
    return RJB_OK;
  } catch (const std::exception& e) {
    delete o;
    return fail(std::string("Cannot load ") + path + ": " + e.what());
  }
}

}  // namespace

extern "C" {

int rjb_graph_read_text(const char* path, rjb_graph* out) {
  return read_guarded(path, out, [](const char* p, GraphOwner* o) { return read_text(p, o); });
}

int rjb_graph_read_bin(const char* path, rjb_graph* out) {
  return read_guarded(path, out, [](const char* p, GraphOwner* o) { return read_bin(p, o); });
}

int rjb_graph_write_bin(const rjb_graph* g, const char* path) {
  if (!g || !path) return RJB_ERR_INVALID;
  FILE* f = fopen(path, "wb");
  if (!f) return fail(std::string("Cannot open ") + path);
  uint64_t magic = 0xabcdabcdull, n_chains = g->n_chains, n_points = g->n_points;
  uint64_t n_row = n_points ? n_chains + 1 : 0;
  fwrite(&magic, 8, 1, f);
  fwrite(&n_chains, 8, 1, f);
  fwrite(&n_row, 8, 1, f);
  fwrite(&n_points, 8, 1, f);
  for (uint64_t i = 0; i < n_chains; i++) {
    fwrite(&g->chain_id[i], 8, 1, f);
    fwrite(&g->first_point[i], 8, 1, f);
    fwrite(&g->last_point[i], 8, 1, f);
    fwrite(&g->left[i], 8, 1, f);
    fwrite(&g->right[i], 8, 1, f);
  }
  if (n_row) fwrite(g->row_index, sizeof(uint32_t), n_row, f);
  if (n_points) fwrite(g->xy, 2 * sizeof(double), n_points, f);
  fwrite(&g->min_x, 8, 1, f);
  fwrite(&g->min_y, 8, 1, f);
  fwrite(&g->max_x, 8, 1, f);
  fwrite(&g->max_y, 8, 1, f);
  fwrite(&magic, 8, 1, f);
  bool ok = ferror(f) == 0;
  ok = (fclose(f) == 0) && ok;
  return ok ? RJB_OK : fail(std::string("Write failed: ") + path);
}

static int graph_load(const char* path, const char* serialize_prefix, rjb_graph* out) {
  std::string prefix = serialize_prefix ? serialize_prefix : "";
  std::string escaped(path);
  for (char& ch : escaped)
    if (ch == '/') ch = '-';
  if (!prefix.empty()) {
    struct stat sb;
    if (stat(prefix.c_str(), &sb) != 0) {
      if (errno == ENOENT) {
        if (mkdir(prefix.c_str(), 0755) != 0) return fail("Cannot create dir " + prefix);
      } else {
        return fail("Cannot open dir " + prefix);
      }
    }
  }
  std::string ser = prefix + "/" + escaped + ".bin";
  if (access(ser.c_str(), R_OK) == 0) return rjb_graph_read_bin(ser.c_str(), out);
  int rc = rjb_graph_read_text(path, out);
  if (rc != RJB_OK) return rc;
  if (!prefix.empty() && access(prefix.c_str(), W_OK) == 0) rjb_graph_write_bin(out, ser.c_str());
  return RJB_OK;
}

int rjb_graph_load(const char* path, const char* serialize_prefix, rjb_graph* out) {
  if (!path || !out) return RJB_ERR_INVALID;
  try {
    return graph_load(path, serialize_prefix, out);
  } catch (const std::exception& e) {
    return fail(std::string("Cannot load ") + path + ": " + e.what());
  }
}

void rjb_graph_free(rjb_graph* g) {
  if (!g || !g->_owner) return;
  delete (GraphOwner*) g->_owner;
  memset(g, 0, sizeof(*g));
}

}  // extern "C"
