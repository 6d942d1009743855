// polyover_exec: flag-compatible stand-in for RayJoin's polyover_exec
// (reference src/overlay.cc:8-43 -> RunOverlay, src/run_overlay.cu:143-228).
//
//   polyover_exec -poly1 A.cdb -poly2 B.cdb -mode=lbvh|grid [-xsect_factor f]
//                 [-grid_size n] [-serialize dir] [-check] [-output out.cdb]
#include <algorithm>

#include "cli_common.h"

using namespace cli;

int main(int argc, char** argv) {
  Flags f;
  std::string err;
  if (argc == 1) {
    std::cerr << "Usage: -poly1 <cdb> -poly2 <cdb> -mode=grid|lbvh [-output file]" << std::endl;
    return 1;
  }
  if (!f.parse(argc, argv, &err)) die(err);
  Timer tm;
  rjb_graph g0{}, g1{};
  tm.next("Read map 0");
  std::cerr << "Reading map 0 from " << f.s("poly1") << std::endl;
  load_graph(f.s("poly1"), f.s("serialize"), &g0);
  tm.next("Read map 1");
  std::cerr << "Reading map 1 from " << f.s("poly2") << std::endl;
  load_graph(f.s("poly2"), f.s("serialize"), &g1);
  tm.next("Create App");
  int mode = parse_mode(f.s("mode"));
  rjb_ctx* ctx = nullptr;
  ok(rjb_create(f.i("device"), &ctx), "rjb_create");
  set_leaf_options(ctx, f);
  tm.next("Load Data");
  set_maps(ctx, &g0, &g1);
  tm.next("Overlay (device)");
  double ms[6];
  ok(rjb_overlay_run(ctx, mode, f.i("grid_size"), f.d("xsect_factor"), ms), "rjb_overlay_run");
  const rjb_xsect* d = nullptr;
  uint64_t n = 0;
  ok(rjb_overlay_results(ctx, 0, &d, &n, nullptr, nullptr), "rjb_overlay_results");
  std::cerr << "Intersections: " << n << std::endl;
  std::vector<uint32_t> eids[2];
  std::vector<rjb_xsect> xs;
  if (f.b("check") && mode != RJB_MODE_GRID) {
    for (int im = 0; im < 2; im++) {
      uint64_t info[3];
      const uint32_t* de = nullptr;
      ok(rjb_map_info(ctx, im, info), "rjb_map_info");
      ok(rjb_overlay_results(ctx, im, nullptr, nullptr, &de, nullptr), "rjb_overlay_results");
      eids[im].resize(info[0]);
      ok(rjb_copy_to_host(ctx, de, eids[im].data(), info[0] * 4), "copy");
    }
  }
  if (!f.s("output").empty()) {
    tm.next("Write to file");
    ok(rjb_overlay_write(ctx, f.s("output").c_str()), "rjb_overlay_write");
  }
  if (f.b("check") && mode != RJB_MODE_GRID) {
    // CheckResult, src/run_overlay.cu:17-141: rebuild with the grid backend and
    // compare the LSI count and the located edges (by end points)
    tm.next("Check result");
    std::cerr << "Checking LSI Results" << std::endl;
    double ms2[6];
    ok(rjb_overlay_run(ctx, RJB_MODE_GRID, f.i("grid_size"), f.d("xsect_factor"), ms2), "overlay(grid)");
    uint64_t n2 = 0;
    ok(rjb_overlay_results(ctx, 0, nullptr, &n2, nullptr, nullptr), "rjb_overlay_results");
    if (n2 != n) std::cerr << "LSI  xsects (Answer): " << n2 << " xsects (Result): " << n << std::endl;
    else std::cerr << "LSI passed check" << std::endl;
    for (int im = 0; im < 2; im++) {
      std::cerr << "Checking point in polygon" << std::endl;
      uint64_t info[3], binfo[3];
      const uint32_t* de = nullptr;
      ok(rjb_map_info(ctx, im, info), "rjb_map_info");
      ok(rjb_map_info(ctx, 1 - im, binfo), "rjb_map_info");
      ok(rjb_overlay_results(ctx, im, nullptr, nullptr, &de, nullptr), "rjb_overlay_results");
      std::vector<uint32_t> ans(info[0]);
      ok(rjb_copy_to_host(ctx, de, ans.data(), info[0] * 4), "copy");
      const int64_t* d_pts = nullptr;
      const uint32_t* d_chain = nullptr;
      ok(rjb_map_device_views(ctx, 1 - im, &d_pts, &d_chain), "views");
      std::vector<int64_t> pts(2 * binfo[0]);
      std::vector<uint32_t> chain(binfo[1]);
      ok(rjb_copy_to_host(ctx, d_pts, pts.data(), pts.size() * 8), "copy");
      ok(rjb_copy_to_host(ctx, d_chain, chain.data(), chain.size() * 4), "copy");
      size_t n_diff = 0;
      for (uint64_t i = 0; i < info[0]; i++) {
        uint32_t a = ans[i], r = eids[im][i];
        if (a == r) continue;
        bool diff = a == RJB_NO_HIT || r == RJB_NO_HIT;
        if (!diff) diff = memcmp(&pts[2 * (uint64_t) (a + chain[a])], &pts[2 * (uint64_t) (r + chain[r])], 32) != 0;
        n_diff += diff;
      }
      if (n_diff) std::cerr << "Map: " << im << " Total points: " << info[0] << " n diff: " << n_diff << std::endl;
      else std::cerr << "Map: " << im << " PIP passed check" << std::endl;
    }
  }
  tm.end();
  std::cerr << "Device phases (CUDA events):" << std::endl
            << " - Build Index: " << ms[0] << " ms" << std::endl
            << " - Intersection edges: " << ms[1] << " ms" << std::endl
            << " - Map 0: Locate vertices in other map: " << ms[2] << " ms" << std::endl
            << " - Map 1: Locate vertices in other map: " << ms[3] << " ms" << std::endl
            << " - Computer output polygons: " << ms[4] << " ms" << std::endl;
  rjb_destroy(ctx);
  rjb_graph_free(&g0);
  rjb_graph_free(&g1);
  return 0;
}
