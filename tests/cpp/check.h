// Minimal self-registering unit-test harness for the cusp:: drop-in headers
// (the reference has its own in testing/unittest/; same idea: DECLARE a test,
// the driver runs every registered one and prints PASS/FAIL lines).
//   ./cusp_api_tests [--host-only] [--list] [name-substring ...]
// Tests registered with space == "device" need a GPU and are skipped under --host-only.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace check {

struct failure {
  std::string msg;
};

struct test_case {
  std::string name;
  bool needs_gpu;
  std::function<void()> fn;
};

inline std::vector<test_case> &registry() {
  static std::vector<test_case> r;
  return r;
}
struct registrar {
  registrar(const std::string &name, bool needs_gpu, std::function<void()> fn) {
    registry().push_back({name, needs_gpu, fn});
  }
};

#define CHECK_FAIL(text)                                                   \
  do {                                                                     \
    std::ostringstream os_;                                                \
    os_ << __FILE__ << ":" << __LINE__ << ": " << text;                    \
    throw check::failure{os_.str()};                                       \
  } while (0)
#define ASSERT_TRUE(c) \
  do {                 \
    if (!(c)) CHECK_FAIL("ASSERT_TRUE(" #c ")"); \
  } while (0)
#define ASSERT_EQUAL(a, b)                                                 \
  do {                                                                     \
    if (!((a) == (b))) CHECK_FAIL("ASSERT_EQUAL(" #a ", " #b ")");       \
  } while (0)
#define ASSERT_NEAR(a, b, tol)                                             \
  do {                                                                     \
    if (!(std::fabs((double)(a) - (double)(b)) <= (tol)))                  \
      CHECK_FAIL("ASSERT_NEAR(" #a ", " #b "): " << (a) << " vs " << (b)); \
  } while (0)
#define ASSERT_THROWS(expr, exc)                                           \
  do {                                                                     \
    bool thrown_ = false;                                                  \
    try {                                                                  \
      expr;                                                                \
    } catch (const exc &) {                                                \
      thrown_ = true;                                                      \
    }                                                                      \
    if (!thrown_) CHECK_FAIL(#expr " did not throw " #exc);                \
  } while (0)

// TEST(name) { ... }                      host only
// SPACE_TEST(name) template<Space> ...    registered for host_memory and device_memory
#define TEST_HOST(fn) static check::registrar reg_##fn(#fn, false, fn);
#define TEST_DEVICE(fn) static check::registrar reg_##fn(#fn, true, fn);
#define TEST_HOST_DEVICE(fn)                                                               \
  static check::registrar reg_h_##fn(#fn "<host_memory>", false, fn<cusp::host_memory>);   \
  static check::registrar reg_d_##fn(#fn "<device_memory>", true, fn<cusp::device_memory>);

inline int run(int argc, char **argv) {
  bool host_only = false, list = false;
  std::vector<std::string> filters;
  for (int i = 1; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--host-only")) host_only = true;
    else if (!std::strcmp(argv[i], "--list")) list = true;
    else filters.push_back(argv[i]);
  }
  int failed = 0, passed = 0, skipped = 0;
  for (const test_case &t : registry()) {
    bool selected = filters.empty();
    for (const std::string &f : filters) selected = selected || t.name.find(f) != std::string::npos;
    if (!selected) continue;
    if (list) {
      std::printf("%s%s\n", t.name.c_str(), t.needs_gpu ? " [gpu]" : "");
      continue;
    }
    if (host_only && t.needs_gpu) {
      ++skipped;
      continue;
    }
    try {
      t.fn();
      ++passed;
      std::printf("PASS %s\n", t.name.c_str());
    } catch (const failure &f) {
      ++failed;
      std::printf("FAIL %s\n     %s\n", t.name.c_str(), f.msg.c_str());
    } catch (const std::exception &e) {
      ++failed;
      std::printf("FAIL %s\n     exception: %s\n", t.name.c_str(), e.what());
    }
    std::fflush(stdout);
  }
  if (!list) std::printf("SUMMARY passed=%d failed=%d skipped=%d\n", passed, failed, skipped);
  return failed ? 1 : 0;
}

}  // namespace check
