// gtest/gtest.h -- a minimal stand-in for GoogleTest (absent from this image), just large enough to run the REFERENCE's own test
// files (tests/warp_test.cpp, tests/utils/test_quaternion.cc, tests/utils/test_dual_quaternion.cc) unchanged: TEST(), the
// ASSERT_/EXPECT_ comparisons they use, and a runner that prints GoogleTest-style lines.  TEST INFRASTRUCTURE.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace testing {

struct TestInfo { const char *suite, *name; void (*body)(); };
inline std::vector<TestInfo> &registry() { static std::vector<TestInfo> r; return r; }
inline bool &current_failed() { static bool f = false; return f; }
struct Registrar { Registrar(const char *s, const char *n, void (*b)()) { TestInfo t = {s, n, b}; registry().push_back(t); } };
inline void InitGoogleTest(int *, char **) {}

inline void report(const char *file, int line, const std::string &what)
{
    current_failed() = true;
    std::printf("%s:%d: Failure\n%s\n", file, line, what.c_str());
}

// GoogleTest's AlmostEquals: within 4 units in the last place
inline bool float_almost_equal(float a, float b)
{
    if (std::isnan(a) || std::isnan(b)) return false;
    int32_t ia, ib;
    std::memcpy(&ia, &a, 4); std::memcpy(&ib, &b, 4);
    const uint32_t ua = ia < 0 ? 0x80000000u - (uint32_t)(ia & 0x7fffffff) : 0x80000000u + (uint32_t)ia;
    const uint32_t ub = ib < 0 ? 0x80000000u - (uint32_t)(ib & 0x7fffffff) : 0x80000000u + (uint32_t)ib;
    return (ua > ub ? ua - ub : ub - ua) <= 4u;
}

inline int run_all_tests()
{
    int failed = 0;
    std::printf("[==========] Running %zu tests.\n", registry().size());
    for (size_t i = 0; i < registry().size(); ++i) {
        const TestInfo &t = registry()[i];
        std::printf("[ RUN      ] %s.%s\n", t.suite, t.name);
        current_failed() = false;
        t.body();
        std::printf(current_failed() ? "[  FAILED  ] %s.%s\n" : "[       OK ] %s.%s\n", t.suite, t.name);
        failed += current_failed();
    }
    std::printf("[==========] %zu tests ran.\n[  PASSED  ] %zu tests.\n", registry().size(), registry().size() - (size_t)failed);
    if (failed) std::printf("[  FAILED  ] %d tests.\n", failed);
    std::fflush(stdout);
    return failed ? 1 : 0;
}

}  // namespace testing

#define RUN_ALL_TESTS() ::testing::run_all_tests()

#define TEST(suite, name)                                                                          \
    static void suite##_##name##_body();                                                           \
    static ::testing::Registrar suite##_##name##_registrar(#suite, #name, &suite##_##name##_body); \
    static void suite##_##name##_body()

#define DFGT_MSG_(expr_text, a, b, extra)                                                                                        \
    do { std::ostringstream os_; os_ << expr_text << "\n  first: " << (a) << "\n second: " << (b) << extra; ::testing::report(__FILE__, __LINE__, os_.str()); } while (0)

#define DFGT_NEAR_(a, b, tol, on_fail)                                                                                            \
    do { const double a_ = (double)(a), b_ = (double)(b), t_ = (double)(tol);                                                     \
         if (!(std::fabs(a_ - b_) <= t_)) { DFGT_MSG_("The difference between " #a " and " #b " exceeds " #tol, a_, b_, ""); on_fail; } } while (0)
#define ASSERT_NEAR(a, b, tol) DFGT_NEAR_(a, b, tol, return)
#define EXPECT_NEAR(a, b, tol) DFGT_NEAR_(a, b, tol, (void)0)

#define DFGT_CMP_(a, b, op, text, on_fail)                                                                                        \
    do { const auto &a_ = (a); const auto &b_ = (b); if (!(a_ op b_)) { DFGT_MSG_("Expected: (" #a ") " text " (" #b ")", a_, b_, ""); on_fail; } } while (0)
#define EXPECT_EQ(a, b) DFGT_CMP_(a, b, ==, "==", (void)0)
#define EXPECT_NE(a, b) DFGT_CMP_(a, b, !=, "!=", (void)0)
#define ASSERT_EQ(a, b) DFGT_CMP_(a, b, ==, "==", return)
#define ASSERT_NE(a, b) DFGT_CMP_(a, b, !=, "!=", return)
#define EXPECT_TRUE(c) do { if (!(c)) ::testing::report(__FILE__, __LINE__, "Value of: " #c "\n  Actual: false"); } while (0)
#define ASSERT_TRUE(c) do { if (!(c)) { ::testing::report(__FILE__, __LINE__, "Value of: " #c "\n  Actual: false"); return; } } while (0)

#define DFGT_FLOAT_EQ_(a, b, on_fail)                                                                                             \
    do { const float a_ = (float)(a), b_ = (float)(b); if (!::testing::float_almost_equal(a_, b_)) { DFGT_MSG_("Expected equality (4 ULP) of " #a " and " #b, a_, b_, ""); on_fail; } } while (0)
#define ASSERT_FLOAT_EQ(a, b) DFGT_FLOAT_EQ_(a, b, return)
#define EXPECT_FLOAT_EQ(a, b) DFGT_FLOAT_EQ_(a, b, (void)0)
