// what libgtest_main provides: the reference's test files have no main() of their own
#include <gtest/gtest.h>
int main(int argc, char **argv) { ::testing::InitGoogleTest(&argc, argv); return RUN_ALL_TESTS(); }
