// div_zero.cpp -- rt_div_zero_fast(a, s) must be bit for bit `a / s` (raytracer-2022_b200/csrc/rt_div.h): special values crossed
// with each other, and random operands.  Build: g++ -O2 -std=c++17 -o div_zero div_zero.cpp
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include "../../raytracer-2022_b200/csrc/rt_div.h"

static uint64_t bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static bool same(double x, double y) { return (x != x && y != y) || bits(x) == bits(y); }

int main() {
    const double sp[] = {0., -0., 1., -1., 4.9e-324, -4.9e-324, 2.2e-308, 1e-300, 1e300, 1.7976931348623157e308, -1.7976931348623157e308,
                         INFINITY, -INFINITY, NAN, 0.5, 3., -7.25};
    long long n = 0;
    for (double a : sp)
        for (double s : sp) {
            volatile double va = a, vs = s;
            double want = va / vs, got = rt_div_zero_fast(a, s);
            if (!same(want, got)) { printf("mismatch %g / %g: %g vs %g\n", a, s, want, got); return 1; }
            ++n;
        }
    std::mt19937_64 rng(5);
    for (int k = 0; k < 2000000; ++k) {
        uint64_t ua = rng(), us = rng();
        if ((k & 3) == 0) ua &= 0x8000000000000000ull; /* a signed zero */
        double a, s;
        memcpy(&a, &ua, 8);
        memcpy(&s, &us, 8);
        volatile double va = a, vs = s;
        double want = va / vs, got = rt_div_zero_fast(a, s);
        if (!same(want, got)) { printf("mismatch %a / %a\n", a, s); return 1; }
        ++n;
    }
    printf("div_zero ok: %lld quotients\n", n);
    return 0;
}
