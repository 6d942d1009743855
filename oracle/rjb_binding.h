/*
 * rjb_binding.h -- the reference-side binding of librjb200: what a RayJoin maintainer adds
 * under src/app/ (INTEGRATION.md section 2), kept here so that it is COMPILED and RUN: it is
 * included by oracle/ref_driver.cu after the reference's own headers and driven from the
 * reference's own Context / PlanarGraph / Stream through its own LSI<CTX> / PIP<CTX>
 * interfaces (`ref_exec lsi rjb`, `ref_exec pip rjb`).  Test infrastructure: nothing in the
 * product includes this file.
 *
 *   LSIRJB<CTX> : LSI<CTX>   (reference src/app/lsi.h:7-43;  sibling of src/app/lsi_lbvh.h)
 *   PIPRJB<CTX> : PIP<CTX>   (reference src/app/pip.h:8-38;  sibling of src/app/pip_lbvh.h)
 *
 * Everything goes through the C ABI of include/rjb200.h.
 */
#ifndef RJB_BINDING_H
#define RJB_BINDING_H
#include <stdexcept>
#include <string>
#include <vector>

#include "app/lsi.h"
#include "app/pip.h"
#include "rjb200.h"

namespace rayjoin {

namespace rjb_detail {
inline void check(int rc, const char* what) {
  if (rc != RJB_OK) throw std::runtime_error(std::string(what) + ": " + rjb_last_error());
}

// One rjb_ctx per RayJoin Context: same bounding box (-> same Scaling), both planar graphs
template <typename CONTEXT_T>
inline rjb_ctx* create_from(CONTEXT_T& ctx) {
  rjb_ctx* h = nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  check(rjb_create(dev, &h), "rjb_create");
  const auto& bb = ctx.get_bounding_box();  // src/context.h:112
  check(rjb_set_bounding_box(h, bb.min_x, bb.min_y, bb.max_x, bb.max_y), "rjb_set_bounding_box");
  for (int im = 0; im < 2; im++) {  // PlanarGraph -> host SoA (src/map/planar_graph.h:24-39)
    auto g = ctx.get_planar_graph(im);
    if (g == nullptr) continue;
    std::vector<int64_t> l, r;
    for (auto& c : g->chains) {
      l.push_back(c.left_polygon_id);
      r.push_back(c.right_polygon_id);
    }
    static_assert(sizeof(g->points[0]) == 2 * sizeof(double), "double2 points");
    // (pinned_vector: thrust pointers -> raw host pointers)
    check(rjb_set_map(h, im, reinterpret_cast<const double*>(thrust::raw_pointer_cast(g->points.data())),
                      g->points.size(), thrust::raw_pointer_cast(g->row_index.data()), l.data(), r.data(),
                      g->chains.size()),
          "rjb_set_map");
  }
  return h;
}

// rjb_xsect {x, y, eid[2], mid_point_polygon_id} -> dev::Intersection<int64_t> {x/1, y/1, ...}:
// the reference always stores denominators 1 (src/util/rational.h:190-192)
template <typename XSECT_T, typename QUEUE_T>
__global__ void convert_xsects(const rjb_xsect* in, uint32_t n, QUEUE_T q) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XSECT_T x;
  x.x = tcb::rational<int64_t>(in[i].x);
  x.y = tcb::rational<int64_t>(in[i].y);
  x.eid[0] = in[i].eid[0];
  x.eid[1] = in[i].eid[1];
  x.mid_point_polygon_id = in[i].mid_point_polygon_id;
  q.AppendWarp(x);
}
}  // namespace rjb_detail

template <typename CONTEXT_T>
class LSIRJB : public LSI<CONTEXT_T> {
  using xsect_t = typename LSI<CONTEXT_T>::xsect_t;

 public:
  // mode: RJB_MODE_LBVH or RJB_MODE_GRID
  LSIRJB(CONTEXT_T& ctx, int mode, unsigned grid_size = 2048)
      : LSI<CONTEXT_T>(ctx), mode_(mode), grid_size_(grid_size) {
    h_ = rjb_detail::create_from(ctx);
  }
  ~LSIRJB() override { rjb_destroy(h_); }

  void Init(size_t max_n_xsects) override {
    LSI<CONTEXT_T>::Init(max_n_xsects);  // the reference-side queue get_xsects() hands out
    // RayJoin sizes the queue as (|E0| + |E1|) * xsect_factor (src/run_query.cu:226-228); the
    // ABI takes the factor
    uint64_t info0[3], info1[3];
    rjb_detail::check(rjb_map_info(h_, 0, info0), "rjb_map_info");
    rjb_detail::check(rjb_map_info(h_, 1, info1), "rjb_map_info");
    xsect_factor_ = (double) (max_n_xsects + 1) / (double) (info0[1] + info1[1]);
  }

  double BuildIndex(int base_map_id) {
    double ms = 0;
    rjb_detail::check(rjb_build_index(h_, base_map_id, mode_, grid_size_, &ms), "rjb_build_index");
    return ms;
  }

  void Query(Stream& stream, int query_map_id) override {
    rjb_detail::check(rjb_set_stream(h_, stream.cuda_stream()), "rjb_set_stream");  // RayJoin's stream
    const rjb_xsect* d = nullptr;
    uint64_t n = 0;
    rjb_detail::check(rjb_lsi(h_, query_map_id, mode_, xsect_factor_, &d, &n, nullptr), "rjb_lsi");
    this->xsect_queue_.Clear(stream);
    if (n) {
      auto q = this->xsect_queue_.DeviceObject();
      rjb_detail::convert_xsects<xsect_t><<<(unsigned) ((n + 255) / 256), 256, 0, stream.cuda_stream()>>>(
          d, (uint32_t) n, q);
    }
    stream.Sync();  // like every reference backend (src/app/lsi_lbvh.h:89)
  }

 private:
  rjb_ctx* h_ = nullptr;
  int mode_;
  unsigned grid_size_;
  double xsect_factor_ = 0.2;
};

template <typename CONTEXT_T>
class PIPRJB : public PIP<CONTEXT_T> {
  using point_t = typename CONTEXT_T::map_t::point_t;

 public:
  PIPRJB(CONTEXT_T& ctx, int mode, unsigned grid_size = 2048)
      : PIP<CONTEXT_T>(ctx), mode_(mode), grid_size_(grid_size) {
    h_ = rjb_detail::create_from(ctx);
  }
  ~PIPRJB() override { rjb_destroy(h_); }

  void Init(size_t n_points) override { this->closest_eids_.resize(n_points); }

  double BuildIndex(int base_map_id) {
    double ms = 0;
    rjb_detail::check(rjb_build_index(h_, base_map_id, mode_, grid_size_, &ms), "rjb_build_index");
    return ms;
  }

  void Query(Stream& stream, int query_map_id, ArrayView<point_t> query_points) override {
    rjb_detail::check(rjb_set_stream(h_, stream.cuda_stream()), "rjb_set_stream");
    static_assert(sizeof(point_t) == 2 * sizeof(int64_t), "long2 points");
    const uint32_t* d_eid = nullptr;
    // RayJoin's device points are scaled int64 (x, y) pairs: handed over as they are
    rjb_detail::check(rjb_pip(h_, query_map_id, mode_, reinterpret_cast<const int64_t*>(query_points.data()),
                              query_points.size(), &d_eid, nullptr, nullptr),
                      "rjb_pip");
    this->closest_eids_.resize(query_points.size());
    cudaMemcpyAsync(thrust::raw_pointer_cast(this->closest_eids_.data()), d_eid,
                    query_points.size() * sizeof(index_t), cudaMemcpyDeviceToDevice, stream.cuda_stream());
    stream.Sync();
  }

 private:
  rjb_ctx* h_ = nullptr;
  int mode_;
  unsigned grid_size_;
};

}  // namespace rayjoin
#endif  // RJB_BINDING_H
