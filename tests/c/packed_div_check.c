/* CPU model of integrate_kernel_v5's arithmetic (dynamicfusion_b200/csrc/tsdf.cu): the operation sequences the kernel
 * writes out for x / z, sqrtf(n) and the running average's division -- a reciprocal / reciprocal-square-root SEED followed by fused
 * multiply-add corrections (ptxas's own fast paths).  The seed here is the correctly rounded reciprocal displaced by k = -2 .. +2 ulp,
 * the corrections are fmaf() (exactly rounded), the reference is the C operator (float division and square root through double are
 * correctly rounded: 53 >= 2 * 24 + 2).  What the model establishes:
 *   - with a correctly rounded seed (k = 0) every sequence returns the IEEE result on the kernel's whole checked domain;
 *   - the tiny-numerator argument (|x| < 2^-80: fma(fx, q, cx) == cx) and the running average hold for every seed;
 *   - the division is NOT seed-independent in the classic hard case -- divisor mantissa all ones -- and the square root fails for a
 *     handful of operands at +-2 ulp: bit-exactness with the '/' operator on the GPU rests on executing the SAME sequence on the SAME
 *     hardware seed as ptxas's expansion of '/', which the on-device self-test (df_integrate_selftest: every divisor mantissa)
 *     checks; this model only shows where that matters.
 * Output: one line per part, "<part> n <cases> k0 <mismatches with k = 0> perturbed <mismatches with k != 0> outside_hard_case <...>". */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint64_t rnd64(void) { uint64_t x = rng_state; x ^= x << 13; x ^= x >> 7; x ^= x << 17; return rng_state = x; }
static double unif(void) { return (double)(rnd64() >> 11) * (1.0 / 9007199254740992.0); }
static float f_from_bits(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static float nudge(float f, int k) { for (; k > 0; --k) f = nextafterf(f, INFINITY); for (; k < 0; ++k) f = nextafterf(f, -INFINITY); return f; }
/* magnitude 2^[lo, hi) with a random mantissa, random sign if sgn */
static float rnd_mag(int lo, int hi, int sgn)
{
    const int e = lo + (int)(rnd64() % (uint64_t)(hi - lo));
    uint32_t bits = ((uint32_t)(e + 127) << 23) | (uint32_t)(rnd64() & 0x7fffffu);
    const uint64_t r = rnd64() % 16;
    if (r == 0) bits &= 0xff800000u;                 /* power of two */
    if (r == 1) bits |= 0x007fffffu;                 /* mantissa all ones */
    if (r == 2) bits = (bits & 0xff800000u) | 1u;    /* just above a power of two */
    if (sgn && (rnd64() & 1)) bits |= 0x80000000u;
    return f_from_bits(bits);
}

static float div_seq(float x, float z, float r0)
{
    const float e = fmaf(r0, -z, 1.f);
    const float r1 = fmaf(r0, e, r0);
    const float q0 = r1 * x;
    const float rem = fmaf(q0, -z, x);
    return fmaf(r1, rem, q0);
}

int main(void)
{
    long long bad = 0, n = 0, bad0 = 0, soft = 0, fail = 0;
    /* 1. projection: z in [6 mm, 64 m), |x| in [2^-80, 64 m) */
    for (long long it = 0; it < 6000000; ++it) {
        const float z = rnd_mag(-8, 6, 0);
        if (z < 6e-3f) continue;
        const float x = (it & 3) ? rnd_mag(-20, 6, 1) : rnd_mag(-80, 6, 1);
        const float ref = x / z;
        const float rt = (float)(1.0 / (double)z);
        uint32_t zb; memcpy(&zb, &z, 4);
        for (int k = -2; k <= 2; ++k) {
            ++n;
            if (div_seq(x, z, nudge(rt, k)) != ref) {
                if (k == 0) ++bad0; else ++bad;
                if ((zb & 0x7fffffu) != 0x7fffffu) ++soft;        /* a mismatch outside the all-ones-divisor hard case */
            }
        }
    }
    printf("division n %lld k0 %lld perturbed %lld outside_hard_case %lld\n", n, bad0, bad, soft);
    fail += bad0 + soft;
    /* 2. |x| < 2^-80 (+-0 and denormals included): the exact and the computed quotient both leave fma(fx, q, cx) == cx for |cx| >= 1 */
    bad = 0; n = 0; bad0 = 0;
    for (long long it = 0; it < 2000000; ++it) {
        const float z = rnd_mag(-8, 6, 0);
        if (z < 6e-3f) continue;
        float x;
        switch (it % 4) { case 0: x = 0.f; break; case 1: x = -0.f; break; case 2: x = f_from_bits((uint32_t)(rnd64() & 0x7fffffu) | ((rnd64() & 1) ? 0x80000000u : 0u)); break;
                          default: x = rnd_mag(-126, -80, 1); }
        const float fx = (float)(1.0 + unif() * 999999.0), cx = (float)((1.0 + unif() * 99999.0) * ((rnd64() & 1) ? 1.0 : -1.0));
        const float want = fmaf(fx, x / z, cx);
        const float rt = (float)(1.0 / (double)z);
        for (int k = -2; k <= 2; ++k) {
            ++n;
            const float got = fmaf(fx, div_seq(x, z, nudge(rt, k)), cx);
            if (got != want || want != cx) { if (k == 0) ++bad0; else ++bad; }
        }
    }
    printf("tiny_numerators n %lld k0 %lld perturbed %lld outside_hard_case %lld\n", n, bad0, bad, bad);
    fail += bad0 + bad;
    /* 3. sqrt: n in [2^-15, 2^62) */
    bad = 0; n = 0; bad0 = 0; soft = 0;
    for (long long it = 0; it < 6000000; ++it) {
        const float v = rnd_mag(-15, 62, 0);
        const float ref = sqrtf(v);
        const float rt = (float)(1.0 / sqrt((double)v));
        for (int k = -2; k <= 2; ++k) {
            const float rs = nudge(rt, k);
            const float s = v * rs, h = rs * 0.5f;
            const float e = fmaf(-s, s, v);
            const float s1 = fmaf(e, h, s);
            ++n;
            if (s1 != ref) { if (k == 0) ++bad0; else if (k == 1 || k == -1) ++soft, ++bad; else ++bad; }
        }
    }
    printf("square_root n %lld k0 %lld perturbed %lld at_one_ulp %lld\n", n, bad0, bad, soft);
    fail += bad0;
    /* 4. running average: |num| in [1e-30, 1e30], den = 1 .. 65536 */
    bad = 0; n = 0; bad0 = 0;
    for (long long it = 0; it < 6000000; ++it) {
        const float num = rnd_mag(-99, 99, 1);
        if (!(fabsf(num) >= 1e-30f && fabsf(num) <= 1e30f)) continue;
        const float den = (float)(1 + (int)(rnd64() % 65536));
        const float ref = num / den;
        const float rt = (float)(1.0 / (double)den);
        for (int k = -2; k <= 2; ++k) {
            ++n;
            if (div_seq(num, den, nudge(rt, k)) != ref) { if (k == 0) ++bad0; else ++bad; }
        }
    }
    printf("running_average n %lld k0 %lld perturbed %lld outside_hard_case %lld\n", n, bad0, bad, bad);
    fail += bad0 + bad;
    /* 5. every divisor mantissa (the sweep df_integrate_selftest runs on the device), exact seed: powers of two, all-ones and a few odd
     *    dividends; and every mantissa of the square root's operand at both exponent parities */
    bad0 = 0; n = 0;
    {
        const float xs[8] = {1.f, 0x1p-11f, -0x1p3f, 0x1.fffffep-4f, -0x1.fffffep2f, 0x1.555556p-1f, 0x1.000002p0f, -0x1.b6db6ep1f};
        for (uint32_t m = 0; m < (1u << 23); ++m) {
            const float z = f_from_bits((127u << 23) | m);
            const float rt = (float)(1.0 / (double)z);
            for (int j = 0; j < 8; ++j) { ++n; if (div_seq(xs[j], z, rt) != xs[j] / z) ++bad0; }
            for (int par = 0; par < 2; ++par) {
                const float v = f_from_bits(((127u + (uint32_t)par) << 23) | m);
                const float rs = (float)(1.0 / sqrt((double)v));
                const float sq = v * rs, h = rs * 0.5f;
                ++n;
                if (fmaf(fmaf(-sq, sq, v), h, sq) != sqrtf(v)) ++bad0;
            }
        }
    }
    printf("every_mantissa n %lld k0 %lld perturbed 0 outside_hard_case 0\n", n, bad0);
    fail += bad0;
    return fail ? 1 : 0;
}
