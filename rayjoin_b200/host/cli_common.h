// Shared pieces of query_exec / polyover_exec: a gflags-compatible flag parser
// for RayJoin's flag set (reference src/flags.cc:3-33) and its phase timer
// (reference src/util/timer.h:43-80, same stderr format).
#pragma once
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "rjb200.h"

namespace cli {

struct Flags {
  // name -> value; defaults are those of src/flags.cc
  std::map<std::string, std::string> v = {
      {"poly1", ""},      {"poly2", ""},        {"output", ""},       {"grid_size", "2048"},
      {"xsect_factor", "0.2"}, {"mode", ""},    {"box", "false"},     {"check", "true"},
      {"fau", "false"},   {"warmup", "5"},      {"repeat", "5"},      {"serialize", ""},
      {"ag", "1"},        {"ag_iter", "5"},     {"win", "32"},        {"enlarge", "5"},
      {"sample_map_id", "-1"}, {"sample", ""},  {"query", ""},        {"sample_rate", "1"},
      {"seed", "0"},      {"gen_t", "0.1"},     {"gen_n", "10000"},   {"histo", "false"},
      {"profile", "false"}, {"v", "0"},         {"device", "0"},      {"lbvh_leaf_size", "4"},
      {"lbvh_ag", "-1"}};
  static bool is_bool(const std::string& n) {
    return n == "box" || n == "check" || n == "fau" || n == "histo" || n == "profile";
  }
  // accepts -flag value, -flag=value, --flag=value, -boolflag, -noboolflag
  bool parse(int argc, char** argv, std::string* err) {
    for (int i = 1; i < argc; i++) {
      std::string a = argv[i];
      if (a.size() < 2 || a[0] != '-') { *err = "unexpected argument " + a; return false; }
      a = a.substr(a[1] == '-' ? 2 : 1);
      std::string name = a, val;
      bool has_val = false;
      size_t eq = a.find('=');
      if (eq != std::string::npos) { name = a.substr(0, eq); val = a.substr(eq + 1); has_val = true; }
      if (!v.count(name)) {
        if (name.rfind("no", 0) == 0 && v.count(name.substr(2)) && is_bool(name.substr(2))) {
          v[name.substr(2)] = "false";
          continue;
        }
        *err = "unknown command line flag '" + name + "'";
        return false;
      }
      if (!has_val) {
        if (is_bool(name)) val = "true";
        else if (i + 1 < argc) val = argv[++i];
        else { *err = "flag '" + name + "' is missing its argument"; return false; }
      }
      v[name] = val;
    }
    return true;
  }
  std::string s(const std::string& n) const { return v.at(n); }
  int i(const std::string& n) const { return atoi(v.at(n).c_str()); }
  double d(const std::string& n) const { return atof(v.at(n).c_str()); }
  bool b(const std::string& n) const {
    const std::string& x = v.at(n);
    return x == "true" || x == "1" || x == "yes" || x == "t" || x == "y";
  }
};

struct Timer {
  std::vector<std::tuple<std::string, double, int>> t;
  static double now() {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + tv.tv_usec / 1000000.0;
  }
  void next(const std::string& name, int repeat = 1) { t.emplace_back(name, now(), repeat); }
  void end() {
    next("end");
    std::cerr << "Timing results:" << std::endl;
    for (size_t k = 0; k + 1 < t.size(); k++) {
      double dt = std::get<1>(t[k + 1]) - std::get<1>(t[k]);
      std::cerr << " - " << std::get<0>(t[k]) << ": " << dt * 1000 / std::get<2>(t[k]) << " ms"
                << std::endl;
      std::cerr << std::endl;
    }
    t.clear();
  }
};

inline void die(const std::string& msg) {
  std::cerr << "FATAL: " << msg << std::endl;
  exit(1);
}

inline void ok(int rc, const char* what) {
  if (rc != RJB_OK) die(std::string(what) + ": " + rjb_last_error());
}

inline int parse_mode(const std::string& mode) {
  if (mode == "grid") return RJB_MODE_GRID;
  if (mode == "lbvh") return RJB_MODE_LBVH;
  if (mode == "brute") return RJB_MODE_BRUTE;
  if (mode == "rt") {
    std::cerr << "WARNING: -mode=rt needs RT cores, which B200 does not have; using the "
                 "CUDA BVH traversal (-mode=lbvh) instead" << std::endl;
    return RJB_MODE_LBVH;
  }
  die("Illegal mode: " + mode);
  return -1;
}

// Leaf shape of the LBVH.  RayJoin's Adaptive Grouping flags (-ag -ag_iter -enlarge,
// src/flags.cc:20-24) belong to its RT backend (src/run_query.cu:237-271); -mode=rt runs the
// CUDA BVH traversal here, so there they size the LBVH leaves with the same merge rule.
// -mode=lbvh keeps fixed runs of -lbvh_leaf_size edges unless -lbvh_ag=1 asks for grouping.
inline void set_leaf_options(rjb_ctx* ctx, const Flags& f) {
  ok(rjb_set_option(ctx, "lbvh_leaf_size", f.i("lbvh_leaf_size")), "rjb_set_option");
  int ag = f.i("lbvh_ag");
  if (ag < 0) ag = f.s("mode") == "rt" ? f.i("ag") : 0;
  ok(rjb_set_option(ctx, "lbvh_ag", ag), "rjb_set_option");
  if (ag) {
    ok(rjb_set_option(ctx, "lbvh_ag_iter", std::max(1, f.i("ag_iter"))), "rjb_set_option");
    ok(rjb_set_option(ctx, "lbvh_enlarge_x1000", (long long) (std::max(1.0, f.d("enlarge")) * 1000.0)),
       "rjb_set_option");
    std::cerr << "Adaptive leaf grouping: enlarge limit " << f.d("enlarge") << " max iter " << f.i("ag_iter")
              << std::endl;
  }
}

inline void load_graph(const std::string& path, const std::string& prefix, rjb_graph* g) {
  ok(rjb_graph_load(path.c_str(), prefix.c_str(), g), "load_from");
  std::cerr << "Map " << path << " is loaded, chains: " << g->n_chains << " points: " << g->n_points
            << " edges: " << g->n_points - g->n_chains << std::endl;
}

inline void set_maps(rjb_ctx* ctx, const rjb_graph* g0, const rjb_graph* g1) {
  double mnx = g0->min_x, mny = g0->min_y, mxx = g0->max_x, mxy = g0->max_y;
  if (g1 && g1->n_points) {
    mnx = std::min(mnx, g1->min_x); mny = std::min(mny, g1->min_y);
    mxx = std::max(mxx, g1->max_x); mxy = std::max(mxy, g1->max_y);
  }
  std::cerr << "Bounding Box, Bottom-left: (" << mnx << ", " << mny << "), Top-right: (" << mxx
            << ", " << mxy << ")" << std::endl;
  ok(rjb_set_bounding_box(ctx, mnx, mny, mxx, mxy), "rjb_set_bounding_box");
  ok(rjb_set_map(ctx, 0, g0->xy, g0->n_points, g0->row_index, g0->left, g0->right, g0->n_chains),
     "rjb_set_map(0)");
  if (g1)
    ok(rjb_set_map(ctx, 1, g1->xy, g1->n_points, g1->row_index, g1->left, g1->right, g1->n_chains),
       "rjb_set_map(1)");
}

}  // namespace cli
