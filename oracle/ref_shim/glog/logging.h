// TEST INFRASTRUCTURE — glog stand-in for the oracle/_ref build: LOG(x) / VLOG(x) swallow their stream, CHECK* abort with a
// message.  (glog is not in this image; the reference only logs progress with it.)
#pragma once
#include <cstdlib>
#include <iostream>
namespace ref_glog {
struct Null { template <typename T> Null& operator<<(const T&) { return *this; } Null& operator<<(std::ostream& (*)(std::ostream&)) { return *this; } };
struct Fatal {
  ~Fatal() { std::cerr << std::endl; std::abort(); }
  template <typename T> Fatal& operator<<(const T& v) { std::cerr << v; return *this; }
};
}  // namespace ref_glog
#define LOG(severity) ref_glog::Null()
#define VLOG(level) ref_glog::Null()
#define DLOG(severity) ref_glog::Null()
#define LOG_IF(severity, cond) ref_glog::Null()
#define CHECK(cond) if (cond) {} else ref_glog::Fatal() << "CHECK failed: " #cond " "
#define CHECK_NOTNULL(p) (p)
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
