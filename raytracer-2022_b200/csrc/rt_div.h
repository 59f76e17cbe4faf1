/*
 * rt_div.h -- the IEEE quotient a / s with a short cut for a zero numerator.
 *
 * div.rn.f64 is a dozen instructions on its fast path and a ~100-instruction subroutine otherwise, and a ZERO numerator is
 * "otherwise" (the fast path needs the numerator's exponent above 2^-969).  Axis-aligned normals and directions, and a zero
 * scattering pdf, divide zeros all the time: a fifth of the Lambertian kernel's instructions were in that subroutine.
 * +-0 / s for a non-zero, non-NaN s (infinities included) is +-0 with the xor of the signs; everything else takes the division,
 * so the result is bit for bit what `a / s` gives (tests/native/div_zero.cpp).
 */
#ifndef RT_DIV_H
#define RT_DIV_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define RT_DIV_HD __host__ __device__ __forceinline__
#else
#define RT_DIV_HD inline
#endif

RT_DIV_HD double rt_div_zero_fast(double a, double s) {
    if (a == 0. && s == s && s != 0.) {
#if defined(__CUDA_ARCH__)
        return __hiloint2double((__double2hiint(a) ^ __double2hiint(s)) & (int)0x80000000, 0);
#else
        uint64_t ua, us;
        memcpy(&ua, &a, 8);
        memcpy(&us, &s, 8);
        ua = (ua ^ us) & 0x8000000000000000ull;
        double r;
        memcpy(&r, &ua, 8);
        return r;
#endif
    }
    return a / s;
}

#endif /* RT_DIV_H */
