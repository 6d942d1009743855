/*
 * ref_driver.cu -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * A replacement for RayJoin's run_query.cu / run_overlay.cu that drives the
 * REFERENCE's own, unmodified grid and lbvh backends (LSIGrid, LSILBVH,
 * PIPGrid, PIPLBVH, MapOverlayGrid, MapOverlayLBVH, Context, the CDB loader),
 * included from the read-only /root/reference tree.  The reference's drivers
 * cannot be used because they include the OptiX RT headers (run_query.cu:7-12).
 * Built by oracle/Makefile (target ref_exec) into oracle/_ref/ref_exec.
 *
 * Phase order and timers follow src/run_query.cu:169-314 (LSI), :316-463 (PIP)
 * and src/run_overlay.cu:143-228 (overlay).  Output: one JSON object on stdout;
 * optional dumps of the results for parity checks.
 *
 *   ref_exec lsi|pip|overlay lbvh|grid MAP0 MAP1 [-grid_size N] [-xsect_factor F]
 *            [-warmup W] [-repeat K] [-dump FILE] [-output FILE] [-v]
 *   ref_exec lsi|pip rjb|rjbgrid MAP0 MAP1 ...   the same driver loop calling librjb200 through
 *            LSIRJB / PIPRJB (oracle/rjb_binding.h): the binding, compiled and run
 *   MAPx: a RayJoin .bin graph (planar_graph.h:128-167) or a CDB text file.
 */
#include <array>
#include <chrono>
#include <fstream>
#include <functional>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include <cuda_runtime.h>

namespace std {
template <>
struct equal_to<double2> {
  bool operator()(const double2& a, const double2& b) const { return a.x == b.x && a.y == b.y; }
};
}  // namespace std

#include "glog/logging.h"
namespace rjb_glog_shim {
int verbose = 0;
}

#include "app/lsi_grid.h"
#include "app/lsi_lbvh.h"
#include "app/map_overlay_grid.h"
#include "app/map_overlay_lbvh.h"
#include "app/pip_grid.h"
#include "app/pip_lbvh.h"
#include "map/planar_graph.h"
#include "tree/primtive.h"
// the reference-side binding of librjb200, driven like any other backend
#include "rjb_binding.h"

using namespace rayjoin;
using context_t = Context<coord_t, coefficient_t>;
using xsect_t = dev::Intersection<typename context_t::internal_coord_t>;
using point_t = typename context_t::map_t::point_t;
using bvh_t = lbvh::bvh<float, segment, aabb_getter>;

static double now_ms() {
  return std::chrono::duration<double, std::milli>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

static std::shared_ptr<PlanarGraph<coord_t>> load_map(const std::string& path) {
  if (path.size() > 4 && path.substr(path.size() - 4) == ".bin")
    return deserialize_pgraph<coord_t>(path.c_str());
  return read_pgraph<coord_t>(path.c_str());
}

struct Args {
  std::string query, mode, map0, map1, dump, output;
  unsigned grid_size = 2048;
  float xsect_factor = 0.2f;
  int warmup = 5, repeat = 5;
};

static std::shared_ptr<bvh_t> build_bvh(context_t& ctx, int map_id) {
  auto& stream = ctx.get_stream();
  auto d_map = ctx.get_map(map_id)->DeviceObject();
  thrust::device_vector<segment> primitives;
  auto bvh = std::make_shared<bvh_t>();
  FillPrimitivesLBVH(stream, d_map, ctx.get_scaling(), primitives);
  stream.Sync();
  bvh->assign(primitives);
  bvh->construct(false);
  cudaDeviceSynchronize();
  return bvh;
}

static int run_lsi(const Args& a) {
  double t0 = now_ms();
  auto g0 = load_map(a.map0), g1 = load_map(a.map1);
  double t_read = now_ms();
  context_t ctx({g0, g1});
  Stream& stream = ctx.get_stream();
  int base_map_id = 0, query_map_id = 1;
  ctx.LoadToDevice();
  cudaDeviceSynchronize();
  double t_load = now_ms();
  auto d_base = ctx.get_map(base_map_id)->DeviceObject();
  auto d_query = ctx.get_map(query_map_id)->DeviceObject();
  size_t queue_cap = (d_base.get_edges_num() + d_query.get_edges_num()) * a.xsect_factor;
  std::shared_ptr<LSI<context_t>> lsi;
  std::shared_ptr<bvh_t> bvh;
  double t_init0 = now_ms(), t_build0, t_build1;
  if (a.mode == "rjb" || a.mode == "rjbgrid") {
    auto p = std::make_shared<LSIRJB<context_t>>(ctx, a.mode == "rjb" ? RJB_MODE_LBVH : RJB_MODE_GRID,
                                                  a.grid_size);
    p->Init(queue_cap);
    t_build0 = now_ms();
    p->BuildIndex(base_map_id);
    t_build1 = now_ms();
    lsi = p;
  } else if (a.mode == "grid") {
    auto grid = std::make_shared<UniformGrid>(a.grid_size);
    auto p = std::make_shared<LSIGrid<context_t>>(ctx, grid);
    QueryConfigGrid qc;
    qc.grid_size = a.grid_size;
    p->set_config(qc);
    p->Init(queue_cap);
    t_build0 = now_ms();
    grid->AddMapsToGrid(ctx, false);
    cudaDeviceSynchronize();
    t_build1 = now_ms();
    lsi = p;
  } else {
    auto p = std::make_shared<LSILBVH<context_t>>(ctx);
    p->Init(queue_cap);
    t_build0 = now_ms();
    bvh = build_bvh(ctx, base_map_id);
    t_build1 = now_ms();
    QueryConfigLBVH qc;
    qc.lbvh = bvh;
    p->set_config(qc);
    lsi = p;
  }
  for (int i = 0; i < a.warmup; i++) lsi->Query(stream, query_map_id);
  cudaDeviceSynchronize();
  double tq0 = now_ms();
  for (int i = 0; i < a.repeat; i++) lsi->Query(stream, query_map_id);
  cudaDeviceSynchronize();
  double tq1 = now_ms();
  auto n = lsi->get_xsects().size();
  if (!a.dump.empty()) {
    thrust::host_vector<xsect_t> xs;
    lsi->CopyTo(xs);
    std::vector<size_t> order(xs.size());
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](size_t i, size_t j) {
      if (xs[i].eid[1] != xs[j].eid[1]) return xs[i].eid[1] < xs[j].eid[1];
      return xs[i].eid[0] < xs[j].eid[0];
    });
    std::ofstream ofs(a.dump);
    for (auto i : order)
      ofs << xs[i].eid[1] << " " << xs[i].eid[0] << " " << (long long) xs[i].x.num() << " "
          << (long long) xs[i].x.denom() << " " << (long long) xs[i].y.num() << " "
          << (long long) xs[i].y.denom() << "\n";
  }
  printf("{\"query\": \"lsi\", \"mode\": \"%s\", \"edges\": [%zu, %zu], \"intersections\": %zu, "
         "\"queue_cap\": %zu, \"phases\": {\"read_ms\": %.3f, \"load_data_ms\": %.3f, "
         "\"init_ms\": %.3f, \"build_index_ms\": %.3f, \"query_ms\": %.4f}, \"warmup\": %d, \"repeat\": %d}\n",
         a.mode.c_str(), d_base.get_edges_num(), d_query.get_edges_num(), (size_t) n, queue_cap,
         t_read - t0, t_load - t_read, t_build0 - t_init0, t_build1 - t_build0,
         (tq1 - tq0) / a.repeat, a.warmup, a.repeat);
  return 0;
}

static int run_pip(const Args& a) {
  auto g0 = load_map(a.map0), g1 = load_map(a.map1);
  context_t ctx({g0, g1});
  Stream& stream = ctx.get_stream();
  int base_map_id = 0, query_map_id = 1;
  double t_load0 = now_ms();
  ctx.LoadToDevice();
  cudaDeviceSynchronize();
  double t_load1 = now_ms();
  thrust::device_vector<point_t> query_points = ctx.get_map(1)->get_points();
  ArrayView<point_t> d_query_points(query_points);
  std::shared_ptr<PIP<context_t>> pip;
  std::shared_ptr<bvh_t> bvh;
  double t_build0, t_build1;
  std::function<void()> query;
  if (a.mode == "rjb" || a.mode == "rjbgrid") {
    auto p = std::make_shared<PIPRJB<context_t>>(ctx, a.mode == "rjb" ? RJB_MODE_LBVH : RJB_MODE_GRID,
                                                  a.grid_size);
    p->Init(query_points.size());
    t_build0 = now_ms();
    p->BuildIndex(base_map_id);
    t_build1 = now_ms();
    query = [=, &stream]() mutable { p->Query(stream, query_map_id, d_query_points); };
    pip = p;
  } else if (a.mode == "grid") {
    auto grid = std::make_shared<UniformGrid>(a.grid_size);
    auto p = std::make_shared<PIPGrid<context_t>>(ctx, grid);
    p->Init(query_points.size());
    t_build0 = now_ms();
    grid->AddMapToGrid(ctx, 0, false);
    cudaDeviceSynchronize();
    t_build1 = now_ms();
    query = [=, &stream]() mutable { p->Query(stream, query_map_id, d_query_points); };
    pip = p;
  } else {
    auto p = std::make_shared<PIPLBVH<context_t>>(ctx);
    p->Init(query_points.size());
    t_build0 = now_ms();
    bvh = build_bvh(ctx, base_map_id);
    t_build1 = now_ms();
    QueryConfigLBVH qc;
    qc.lbvh = bvh;
    p->set_config(qc);
    query = [=, &stream]() mutable { p->Query(stream, query_map_id, d_query_points); };
    pip = p;
  }
  for (int i = 0; i < a.warmup; i++) query();
  cudaDeviceSynchronize();
  double tq0 = now_ms();
  for (int i = 0; i < a.repeat; i++) query();
  cudaDeviceSynchronize();
  double tq1 = now_ms();
  if (!a.dump.empty()) {
    thrust::host_vector<index_t> eids = pip->get_closest_eids();
    std::ofstream ofs(a.dump, std::ios::binary);
    ofs.write(reinterpret_cast<const char*>(eids.data()), eids.size() * sizeof(index_t));
  }
  printf("{\"query\": \"pip\", \"mode\": \"%s\", \"points\": %zu, \"phases\": {\"load_data_ms\": %.3f, "
         "\"build_index_ms\": %.3f, \"query_ms\": %.4f}, \"warmup\": %d, \"repeat\": %d}\n",
         a.mode.c_str(), (size_t) query_points.size(), t_load1 - t_load0, t_build1 - t_build0,
         (tq1 - tq0) / a.repeat, a.warmup, a.repeat);
  return 0;
}

static int run_overlay(const Args& a) {
  auto g0 = load_map(a.map0), g1 = load_map(a.map1);
  context_t ctx({g0, g1});
  std::shared_ptr<MapOverlay<context_t>> overlay;
  if (a.mode == "grid") {
    auto o = std::make_shared<MapOverlayGrid<context_t>>(ctx);
    QueryConfigGrid qc;
    qc.grid_size = a.grid_size;
    qc.xsect_factor = a.xsect_factor;
    o->set_config(qc);
    overlay = o;
  } else {
    auto o = std::make_shared<MapOverlayLBVH<context_t>>(ctx);
    QueryConfigLBVH qc;
    qc.xsect_factor = a.xsect_factor;
    o->set_config(qc);
    overlay = o;
  }
  double t[8];
  t[0] = now_ms();
  ctx.LoadToDevice();
  overlay->Init();
  cudaDeviceSynchronize();
  t[1] = now_ms();
  overlay->BuildIndex();
  cudaDeviceSynchronize();
  t[2] = now_ms();
  overlay->IntersectEdge(0);
  cudaDeviceSynchronize();
  t[3] = now_ms();
  overlay->LocateVerticesInOtherMap(0);
  cudaDeviceSynchronize();
  t[4] = now_ms();
  overlay->LocateVerticesInOtherMap(1);
  cudaDeviceSynchronize();
  t[5] = now_ms();
  overlay->ComputeOutputPolygons();
  cudaDeviceSynchronize();
  t[6] = now_ms();
  size_t n = overlay->get_xsect_edges().size();
  if (!a.output.empty()) overlay->WriteResult(a.output.c_str());
  t[7] = now_ms();
  if (!a.dump.empty()) {
    FOR2 {
      auto eids = overlay->get_closet_eids(im);
      auto pip = overlay->get_point_in_polygon(im);
      std::ofstream ofs(a.dump + ".pip" + std::to_string(im), std::ios::binary);
      ofs.write(reinterpret_cast<const char*>(eids.data()), eids.size() * sizeof(index_t));
      ofs.write(reinterpret_cast<const char*>(pip.data()), pip.size() * sizeof(index_t));
    }
  }
  printf("{\"query\": \"overlay\", \"mode\": \"%s\", \"intersections\": %zu, \"phases\": {"
         "\"load_init_ms\": %.3f, \"build_index_ms\": %.3f, \"lsi_ms\": %.3f, \"pip0_ms\": %.3f, "
         "\"pip1_ms\": %.3f, \"polygons_ms\": %.3f, \"write_ms\": %.3f}}\n",
         a.mode.c_str(), n, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4],
         t[6] - t[5], t[7] - t[6]);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: ref_exec lsi|pip|overlay lbvh|grid MAP0 MAP1 [flags]\n");
    return 2;
  }
  Args a;
  a.query = argv[1];
  a.mode = argv[2];
  a.map0 = argv[3];
  a.map1 = argv[4];
  for (int i = 5; i < argc; i++) {
    std::string f = argv[i];
    auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
    if (f == "-grid_size") a.grid_size = std::stoul(next());
    else if (f == "-xsect_factor") a.xsect_factor = std::stof(next());
    else if (f == "-warmup") a.warmup = std::stoi(next());
    else if (f == "-repeat") a.repeat = std::stoi(next());
    else if (f == "-dump") a.dump = next();
    else if (f == "-output") a.output = next();
    else if (f == "-v") rjb_glog_shim::verbose = 1;
    else { fprintf(stderr, "unknown flag %s\n", f.c_str()); return 2; }
  }
  const bool rjb_mode = a.mode == "rjb" || a.mode == "rjbgrid";
  if (a.mode != "grid" && a.mode != "lbvh" && !(rjb_mode && a.query != "overlay")) {
    fprintf(stderr, "mode must be grid|lbvh (lsi / pip also: rjb|rjbgrid)\n");
    return 2;
  }
  try {
    if (a.query == "lsi") return run_lsi(a);
    if (a.query == "pip") return run_pip(a);
    if (a.query == "overlay") return run_overlay(a);
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_exec: %s\n", e.what());
    return 1;
  }
  fprintf(stderr, "unknown query %s\n", a.query.c_str());
  return 2;
}
