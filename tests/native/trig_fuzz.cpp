// trig_fuzz.cpp -- pins raytracer-2022_b200/csrc/rt_trig.h (sin and cos of phi in [0, 2 pi] for the direction sampling of
// basic/pdf.rs:12-21 and basic/vec.rs:108-117) against long double sinl / cosl: maximum error in ulps of the double result.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -o trig_fuzz trig_fuzz.cpp ; run: ./trig_fuzz [millions] [seed]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../raytracer-2022_b200/csrc/rt_trig.h"

static double ulps(double got, long double ref) {
    double r = (double)ref;
    if (r == 0.0) return got == 0.0 ? 0.0 : fabs(got) / 4.9e-324;
    int e;
    frexp(r, &e);
    long double ulp = ldexpl(1.0L, e - 53);
    return (double)(fabsl((long double)got - ref) / ulp);
}

int main(int argc, char **argv) {
    long long n = (argc > 1 ? atoll(argv[1]) : 20) * 1000000LL;
    std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 12345);
    double worst_s = 0., worst_c = 0., at_s = 0., at_c = 0.;
    long long refused = 0;
    auto probe = [&](double phi) {
        double s, c;
        if (!rt_sincos_2pi(phi, s, c)) { ++refused; return; }
        double es = ulps(s, sinl((long double)phi)), ec = ulps(c, cosl((long double)phi));
        if (es > worst_s) { worst_s = es; at_s = phi; }
        if (ec > worst_c) { worst_c = ec; at_c = phi; }
    };
    const double two_pi = 2. * 3.141592653589793;
    for (long long i = 0; i < n; ++i) {
        double r1 = (double)(rng() >> 11) * (1.0 / 9007199254740992.0); /* the 53-bit uniform of rt_rng::gen() */
        probe(two_pi * r1);                                            /* phi exactly as the callers form it */
    }
    /* the quadrant boundaries and the ends of the interval, a few thousand neighbours of each */
    for (int k = 0; k <= 8; ++k) {
        double x = k * 0.7853981633974483;
        double lo = x, hi = x;
        for (int j = 0; j < 4000; ++j) {
            probe(lo); probe(hi);
            lo = nextafter(lo, -1.); hi = nextafter(hi, 10.);
        }
    }
    for (double x : {0.0, 4.9e-324, 1e-300, 1e-20, 1e-9, 6.283185307179586, 6.2831853071795862, 6.39}) probe(x);
    /* outside the interval the routine must refuse, not extrapolate */
    double s = 7., c = 7.;
    bool bad = rt_sincos_2pi(6.5, s, c) || rt_sincos_2pi(-0.2, s, c) || rt_sincos_2pi(NAN, s, c) || rt_sincos_2pi(INFINITY, s, c) || s != 7. || c != 7.;
    printf("max error sin %.3f ulp at %.17g, cos %.3f ulp at %.17g, refused %lld, outside %s\n", worst_s, at_s, worst_c, at_c, refused,
           bad ? "ACCEPTED" : "refused");
    return (worst_s <= 1.0 && worst_c <= 1.0 && refused == 0 && !bad) ? 0 : 1;
}
