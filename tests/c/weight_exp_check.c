/* Test infrastructure: exhaustive check that the polynomial of node_weighting() (dynamicfusion_b200/csrc/warp_common.cuh) rounds to the
 * same float as (float)exp((double)x) of the host libm for every float x in [-0.5, -0].  Same constants and operation order as the CUDA
 * code (fused multiply-adds in double).  Prints "n <count> mismatches <count>". */
#include <math.h>
#include <stdint.h>
#include <stdio.h>

static inline float poly_weight(float xf)
{
    const double r = (double)xf + 0.25;
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return (float)(0.77880078307140486825 * p);
}

int main(void)
{
    long long mism = 0, n = 0;
    union { float f; uint32_t u; } hi, v;
    hi.f = 0.5f;
    for (uint32_t u = 0; u <= hi.u; ++u) {
        v.u = u | 0x80000000u;
        const float a = poly_weight(v.f), b = (float)exp((double)v.f);
        if (a != b && mism++ < 5) printf("x %a polynomial %a libm %a\n", v.f, a, b);
        ++n;
    }
    printf("n %lld mismatches %lld\n", n, mism);
    return mism != 0;
}
