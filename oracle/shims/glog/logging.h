// Minimal glog stand-in for building the reference's headers (test infrastructure).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace rjb_glog_shim {
struct Sink {
  bool fatal, live;
  std::ostringstream ss;
  Sink(bool f, bool l) : fatal(f), live(l) {}
  ~Sink() {
    if (live) std::cerr << ss.str() << std::endl;
    if (fatal) std::abort();
  }
  template <typename T>
  Sink& operator<<(const T& v) {
    if (live) ss << v;
    return *this;
  }
};
struct Voidify {
  void operator&(const Sink&) {}
};
extern int verbose;
}  // namespace rjb_glog_shim
#define RJB_SEV_INFO false
#define RJB_SEV_WARNING false
#define RJB_SEV_ERROR false
#define RJB_SEV_FATAL true
#define LOG(sev) ::rjb_glog_shim::Sink(RJB_SEV_##sev, RJB_SEV_##sev || ::rjb_glog_shim::verbose > 0)
#define VLOG(n) ::rjb_glog_shim::Sink(false, ::rjb_glog_shim::verbose >= (n))
#define CHECK(cond) \
  (cond) ? (void) 0 : ::rjb_glog_shim::Voidify() & ::rjb_glog_shim::Sink(true, true) << "Check failed: " #cond " "
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
