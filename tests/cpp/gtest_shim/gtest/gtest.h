// Minimal stand-in for the parts of GoogleTest that the reference's tests use (TEST, TEST_P + TestWithParam +
// INSTANTIATE_TEST_SUITE_P + Values, EXPECT_*, InitGoogleTest / RUN_ALL_TESTS, --gtest_filter with ':'-separated
// patterns, '*' wildcards and a '-' negative section).  GoogleTest is not installed in this image; with this header
// the reference's own test sources (tests/cals/test_cals.cpp, tests/als/test_als.cpp) compile unchanged against this
// repository's cals:: library, and the restated tests under tests/cpp use the same macros.
#ifndef CALS_B200_GTEST_SHIM_H
#define CALS_B200_GTEST_SHIM_H

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <exception>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

namespace testing {

namespace internal {
struct Case {
  std::string name;
  std::function<void()> run;
};
inline std::vector<Case> &cases() {
  static std::vector<Case> v;
  return v;
}
inline std::vector<std::function<void()>> &expanders() {
  static std::vector<std::function<void()>> v;
  return v;
}
inline int &failures_of_current() {
  static int n = 0;
  return n;
}
inline std::string &filter() {
  static std::string f = "*";
  return f;
}
inline bool glob(const char *p, const char *s) {
  if (!*p)
    return !*s;
  if (*p == '*')
    return glob(p + 1, s) || (*s && glob(p, s + 1));
  return *s && (*p == '?' || *p == *s) && glob(p + 1, s + 1);
}
inline bool any_match(const std::string &patterns, const std::string &name) {
  std::stringstream ss(patterns);
  std::string p;
  while (std::getline(ss, p, ':'))
    if (!p.empty() && glob(p.c_str(), name.c_str()))
      return true;
  return false;
}
inline bool selected(const std::string &name) {
  const std::string &f = filter();
  const size_t dash = f.find('-');
  const std::string pos = dash == std::string::npos ? f : f.substr(0, dash);
  const std::string neg = dash == std::string::npos ? "" : f.substr(dash + 1);
  return any_match(pos.empty() ? "*" : pos, name) && !any_match(neg, name);
}
inline void report_failure(const char *file, int line, const std::string &what) {
  failures_of_current()++;
  std::cout << file << ":" << line << ": Failure\n" << what << std::endl;
}
struct Registrar {
  Registrar(const std::string &name, std::function<void()> fn) { cases().push_back({name, std::move(fn)}); }
};
struct ExpanderRegistrar {
  explicit ExpanderRegistrar(std::function<void()> fn) { expanders().push_back(std::move(fn)); }
};
template <class Fixture> struct ParamTests {
  static std::vector<std::pair<std::string, std::function<Fixture *()>>> &list() {
    static std::vector<std::pair<std::string, std::function<Fixture *()>>> v;
    return v;
  }
};
} // namespace internal

class Test {
public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};

template <class T> class TestWithParam : public Test {
public:
  using ParamType = T;
  const T &GetParam() const { return *param_; }
  void set_param_(const T *p) { param_ = p; }

private:
  const T *param_{nullptr};
};

template <class... A> std::tuple<A...> Values(A... a) { return std::tuple<A...>(a...); }

inline void InitGoogleTest(int *argc, char **argv) {
  for (int i = 1; i < *argc; i++)
    if (std::strncmp(argv[i], "--gtest_filter=", 15) == 0)
      internal::filter() = argv[i] + 15;
}

inline int RunAllTests() {
  for (auto &e : internal::expanders())
    e();
  int ran = 0, failed = 0;
  std::vector<std::string> failed_names;
  for (auto &c : internal::cases()) {
    if (!internal::selected(c.name))
      continue;
    std::cout << "[ RUN      ] " << c.name << std::endl;
    internal::failures_of_current() = 0;
    const auto t0 = std::chrono::steady_clock::now();
    try {
      c.run();
    } catch (const std::exception &ex) {
      internal::report_failure("<exception>", 0, std::string("uncaught exception: ") + ex.what());
    } catch (const std::string &s) {
      internal::report_failure("<exception>", 0, "uncaught std::string: " + s);
    } catch (...) {
      internal::report_failure("<exception>", 0, "uncaught exception of unknown type");
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ran++;
    if (internal::failures_of_current()) {
      failed++;
      failed_names.push_back(c.name);
      std::cout << "[  FAILED  ] " << c.name << " (" << (long)ms << " ms)" << std::endl;
    } else
      std::cout << "[       OK ] " << c.name << " (" << (long)ms << " ms)" << std::endl;
  }
  std::cout << "[==========] " << ran << " tests ran." << std::endl;
  std::cout << "[  PASSED  ] " << ran - failed << " tests." << std::endl;
  if (failed) {
    std::cout << "[  FAILED  ] " << failed << " tests, listed below:" << std::endl;
    for (auto &n : failed_names)
      std::cout << "[  FAILED  ] " << n << std::endl;
  }
  return failed ? 1 : 0;
}

} // namespace testing

#define RUN_ALL_TESTS() ::testing::RunAllTests()

#define GTEST_SHIM_CAT_(a, b) a##b
#define GTEST_SHIM_CAT(a, b) GTEST_SHIM_CAT_(a, b)

#define TEST(suite, name)                                                                                              \
  class suite##_##name##_Test : public ::testing::Test {                                                               \
  public:                                                                                                              \
    void TestBody() override;                                                                                          \
  };                                                                                                                   \
  static ::testing::internal::Registrar suite##_##name##_registrar(#suite "." #name, [] {                              \
    suite##_##name##_Test t;                                                                                           \
    t.SetUp();                                                                                                         \
    t.TestBody();                                                                                                      \
    t.TearDown();                                                                                                      \
  });                                                                                                                  \
  void suite##_##name##_Test::TestBody()

#define TEST_P(fixture, name)                                                                                          \
  class fixture##_##name##_Test : public fixture {                                                                     \
  public:                                                                                                              \
    void TestBody() override;                                                                                          \
  };                                                                                                                   \
  static int fixture##_##name##_registered =                                                                           \
      (::testing::internal::ParamTests<fixture>::list().push_back(                                                     \
           {#fixture "." #name, []() -> fixture * { return new fixture##_##name##_Test; }}),                           \
       0);                                                                                                             \
  void fixture##_##name##_Test::TestBody()

#define INSTANTIATE_TEST_SUITE_P(prefix, fixture, values)                                                              \
  static ::testing::internal::ExpanderRegistrar GTEST_SHIM_CAT(prefix##_##fixture##_expander_, __LINE__)([] {          \
    auto vals = values;                                                                                                \
    auto *params = new std::deque<fixture::ParamType>;                                                                \
    std::apply([&](auto... v) { (params->push_back(static_cast<fixture::ParamType>(v)), ...); }, vals);                \
    for (size_t i = 0; i < params->size(); i++)                                                                        \
      for (auto &t : ::testing::internal::ParamTests<fixture>::list()) {                                               \
        auto make = t.second;                                                                                          \
        const fixture::ParamType *p = &(*params)[i];                                                                   \
        ::testing::internal::cases().push_back({std::string(#prefix "/") + t.first + "/" + std::to_string(i), [make, p] { \
                                                  fixture *f = make();                                                 \
                                                  f->set_param_(p);                                                    \
                                                  f->SetUp();                                                          \
                                                  f->TestBody();                                                       \
                                                  f->TearDown();                                                       \
                                                  delete f;                                                            \
                                                }});                                                                   \
      }                                                                                                                \
  })
#define INSTANTIATE_TEST_CASE_P INSTANTIATE_TEST_SUITE_P

#define GTEST_SHIM_EXPECT(cond, text)                                                                                  \
  do {                                                                                                                 \
    if (!(cond)) {                                                                                                     \
      std::ostringstream os_;                                                                                          \
      os_.precision(17);                                                                                               \
      os_ << text;                                                                                                     \
      ::testing::internal::report_failure(__FILE__, __LINE__, os_.str());                                              \
    }                                                                                                                  \
  } while (0)

#define EXPECT_TRUE(c) GTEST_SHIM_EXPECT((c), "Expected true: " #c)
#define EXPECT_FALSE(c) GTEST_SHIM_EXPECT(!(c), "Expected false: " #c)
#define EXPECT_EQ(a, b) GTEST_SHIM_EXPECT((a) == (b), "Expected equality of " #a " (" << (a) << ") and " #b " (" << (b) << ")")
#define EXPECT_NE(a, b) GTEST_SHIM_EXPECT((a) != (b), "Expected " #a " != " #b)
#define EXPECT_LT(a, b) GTEST_SHIM_EXPECT((a) < (b), "Expected " #a " (" << (a) << ") < " #b " (" << (b) << ")")
#define EXPECT_LE(a, b) GTEST_SHIM_EXPECT((a) <= (b), "Expected " #a " (" << (a) << ") <= " #b " (" << (b) << ")")
#define EXPECT_GT(a, b) GTEST_SHIM_EXPECT((a) > (b), "Expected " #a " (" << (a) << ") > " #b " (" << (b) << ")")
#define EXPECT_GE(a, b) GTEST_SHIM_EXPECT((a) >= (b), "Expected " #a " (" << (a) << ") >= " #b " (" << (b) << ")")
#define EXPECT_NEAR(a, b, tol)                                                                                         \
  GTEST_SHIM_EXPECT(std::fabs((double)(a) - (double)(b)) <= (tol),                                                     \
                    "The difference between " #a " (" << (a) << ") and " #b " (" << (b) << ") exceeds " #tol " ("      \
                                                      << (tol) << ")")
#define ASSERT_TRUE(c)                                                                                                 \
  do {                                                                                                                 \
    EXPECT_TRUE(c);                                                                                                    \
    if (!(c))                                                                                                          \
      return;                                                                                                          \
  } while (0)
#define ASSERT_EQ(a, b)                                                                                                \
  do {                                                                                                                 \
    EXPECT_EQ(a, b);                                                                                                   \
    if (!((a) == (b)))                                                                                                 \
      return;                                                                                                          \
  } while (0)

#endif
