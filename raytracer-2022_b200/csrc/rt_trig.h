/*
 * rt_trig.h -- sin and cos of an angle in [0, 2 pi], together, in ~70 instructions.
 *
 * The reference draws a direction as (cos(phi) sqrt(r2), sin(phi) sqrt(r2), sqrt(1 - r2)) with phi = 2 pi r1
 * (random_cosine_direction, basic/pdf.rs:12-21; random_to_sphere, basic/vec.rs:108-117) through Rust's f64::sin / f64::cos,
 * i.e. the platform libm.  No two libms agree to the last bit (sin/cos/acos/log are compared to 1e-9, not bitwise),
 * so the only requirement on the device side is an error well under 1 ulp-and-a-bit, like any libm.  CUDA's sincos() is that,
 * but it is a general-argument routine and, under this file's -fmad=false, costs ~310 instructions per call: a fifth of the
 * Lambertian shade kernel.  phi never leaves [0, 2 pi), so:
 *
 *   q = rint(phi * 2/pi) in 0..4,  r = phi - q * pi/2 in two pieces (Cody-Waite, pi/2 = HI + LO; exact to 2^-105 for q <= 4,
 *   the rounding error of r kept as a tail),  then the classic minimax kernels on |r| <= pi/4 (the coefficients and the
 *   evaluation order of Sun's fdlibm __kernel_sin / __kernel_cos, error < 1 ulp),  then the quadrant swap.
 *
 * Outside [-0.1, 6.4] (never reached from the two callers; NaN included) the caller's fallback is used.
 * tests/native/trig_fuzz.cpp pins the error against long double sinl/cosl (max 0.8 ulp over 2 * 10^7 arguments).
 */
#ifndef RT_TRIG_H
#define RT_TRIG_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RT_TRIG_HD __host__ __device__ __forceinline__
#else
#define RT_TRIG_HD inline
#endif

/* x with the low word cleared and 2 taken off the exponent: fdlibm's `qx` (x/4 truncated), for |x| in [0.3, 0.78125] */
RT_TRIG_HD double rt_trig_quarter_hi(double ax) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(ax) - 0x00200000, 0);
#else
    uint64_t u;
    memcpy(&u, &ax, 8);
    u = ((u >> 32) - 0x00200000ull) << 32;
    double r;
    memcpy(&r, &u, 8);
    return r;
#endif
}

/* sin(x + y), cos(x + y) for |x| <= ~pi/4, |y| << |x| (tail of the argument reduction) */
RT_TRIG_HD void rt_trig_kernels(double x, double y, double &s, double &c) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x;
    {
        const double v = z * x;
        const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
        s = x - ((z * (0.5 * y - v * r) - y) - v * S1);
    }
    {
        const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
        const double ax = fabs(x);
        if (ax < 0.3) {
            c = 1.0 - (0.5 * z - (z * r - x * y));
        } else {
            const double qx = ax > 0.78125 ? 0.28125 : rt_trig_quarter_hi(ax);
            const double hz = 0.5 * z - qx;
            const double a = 1.0 - qx;
            c = a - (hz - (z * r - x * y));
        }
    }
}

/* returns false (s, c untouched) when phi is outside the supported interval */
RT_TRIG_HD bool rt_sincos_2pi(double phi, double &s, double &c) {
    if (!(phi >= -0.1 && phi <= 6.4)) return false;
    const double TWO_OVER_PI = 6.36619772367581382433e-01;
    const double PIO2_HI = 1.57079632679489655800e+00; /* 0x3FF921FB54442D18 */
    const double PIO2_LO = 6.12323399573676603587e-17; /* 0x3C91A62633145C07: pi/2 - PIO2_HI */
    const double q = rint(phi * TWO_OVER_PI);
    const double t = fma(-q, PIO2_HI, phi); /* exact up to one rounding of a value that the cancellation makes (nearly) exact */
    const double w = q * PIO2_LO;
    const double r = t - w;
    const double tail = (t - r) - w;
    double ks, kc;
    rt_trig_kernels(r, tail, ks, kc);
    const int n = (int)q & 3;
    s = (n & 1) ? kc : ks;
    c = (n & 1) ? ks : kc;
    if (n & 2) s = -s;
    if ((n + 1) & 2) c = -c;
    return true;
}

#endif /* RT_TRIG_H */
