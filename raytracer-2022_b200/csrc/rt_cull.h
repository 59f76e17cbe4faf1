/*
 * rt_cull.h -- single-precision pre-test of AABB::hit (hittable/bvh/aabb.rs:15-32) with a proven error margin.
 *
 * The reference decides every BvhNode box in f64:  t0/t1 = (plane - origin) * (1/dir) per axis, swapped when
 * 1/dir < 0, t_min = max(t_min, t0), t_max = min(t_max, t1), miss iff t_max <= t_min.  Hit/miss must be
 * bit-identical to that, but 99.99 % of the boxes a ray meets are nowhere near the decision boundary.  cull_box()
 * evaluates the same slabs in f32 from f32 copies of the box and of the ray and returns
 *
 *      RT_CULL_MISS / RT_CULL_HIT   only when the f64 test is CERTAIN to give that answer,
 *      RT_CULL_UNSURE               otherwise -- the caller then runs the f64 test (hit_aabb) on the f64 box.
 *
 * Error bound (all roundings to nearest, 2^-24 relative; fmaf rounds once).  With b~ = f32(b), inv~ = f32(1/d),
 * noi~ = f32(-o*(1/d)) the f32 slab value  t~ = fmaf(b~, inv~, noi~)  differs from the value the reference
 * computes, t = fl(fl(b - o) * fl(1/d)), by at most
 *
 *      |t~ - t| <= 2^-24 * (3|b| + 2|o|) * |1/d| * (1 + 2^-20)            [b~, inv~: 2|b||inv|; noi~: |o||inv|;
 *                                                                          fmaf: |t~| <= (|b|+|o|)|inv|; f64: 2^-50]
 * so with B = the largest |coordinate| of any BvhNode box of the scene (at least 16, which also covers f32
 * subnormal results) every slab value of a ray is within  E = 2^-24 (3B + 2 max|o_a|) max|1/d_a|  of the f64 one,
 * max/min chains are 1-Lipschitz, and f32(t_min) is within 2^-24 |t_min| of t_min.  The running best enters as
 * bestL <= best <= bestU (directed roundings), hence
 *
 *      min(t1~.., bestL) - max(t0~.., t_min~) >  2E + 2^-23|t_min|   =>  f64 t_max > t_min   (hit),
 *      min(t1~.., bestU) - max(t0~.., t_min~) < -2E - 2^-23|t_min|   =>  f64 t_max < t_min   (miss).
 *
 * The margin e2 below is that bound times 1.09 (which also absorbs the rounding of the f32 subtraction and of e2
 * itself).  A ray whose f32 copies could overflow, underflow or turn NaN (|1/d| outside [1e-30, 1e30], slab values
 * beyond 1e37, a zero direction component, NaN anywhere) gets e2 = +inf: every box is then RT_CULL_UNSURE and the
 * ray is decided in f64 throughout.  Zero-thickness node boxes (Q15) give t0~ == t1~, a gap <= 0 that is never
 * below -e2 unless another axis already misses: they too end in the f64 test.
 *
 * Shared by the CUDA kernels and by tests/cull_fuzz.cpp (host build), which pins it against the f64 test on
 * random and adversarial (grazing) ray/box pairs.
 */
#ifndef RT_CULL_H
#define RT_CULL_H

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_CULL_HD __host__ __device__ __forceinline__
#else
#define RT_CULL_HD inline
#endif

enum { RT_CULL_MISS = 0, RT_CULL_HIT = 1, RT_CULL_UNSURE = 2 };

#define RT_CULL_MIN_EXTENT 16.0f /* lower bound of B: keeps e2 above the f32 subnormal granularity (see above) */

/* f32 view of a ray for the pre-test (scalar members: an array indexed in a loop would push the enclosing
 * traversal state into local memory on the device) */
struct CullRay {
    float ix, iy, iz; /* f32(1/d) */
    float nx, ny, nz; /* f32(-o * (1/d)) */
    float e2;         /* decision margin; +inf = decide everything in f64 */
    float tmin;       /* f32(t_min) */
};

RT_CULL_HD float rt_f32_down(double x) { /* largest float <= x */
#if defined(__CUDA_ARCH__)
    return __double2float_rd(x);
#else
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return f;
#endif
}
RT_CULL_HD float rt_f32_up(double x) { /* smallest float >= x */
#if defined(__CUDA_ARCH__)
    return __double2float_ru(x);
#else
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
#endif
}

/* one axis of cull_ray_setup: returns (3B + 2|o|) |1/d|, clears ok when the f32 copies could misbehave */
RT_CULL_HD double cull_axis_setup(double o, double inv, double extent, float &fi, float &fn, bool &ok) {
    double ai = fabs(inv);
    fi = (float)inv;
    fn = (float)(-(o * inv));
    double e = (3. * extent + 2. * fabs(o)) * ai;
    ok = ok && (ai >= 1e-30) && (ai <= 1e30) && (e <= 1e37); /* NaN fails every comparison */
    return e;
}

/* o, inv = 1/d as the reference computes it (f64 division), t_min of the query, extent = B of the scene */
RT_CULL_HD void cull_ray_setup(CullRay &c, double ox, double oy, double oz, double ivx, double ivy, double ivz, double tmin,
                               float extent) {
    bool ok = tmin == tmin;
    double ex = cull_axis_setup(ox, ivx, (double)extent, c.ix, c.nx, ok);
    double ey = cull_axis_setup(oy, ivy, (double)extent, c.iy, c.ny, ok);
    double ez = cull_axis_setup(oz, ivz, (double)extent, c.iz, c.nz, ok);
    double emax = ex > ey ? ex : ey;
    emax = ez > emax ? ez : emax;
    /* 2 * 2^-24 * 1.09 = 1.2994e-7;  2^-23 * 1.09 * 2 = 2.6e-7 */
    double e2 = 1.3e-7 * emax + (fabs(tmin) <= 1.7e308 ? 2.6e-7 * fabs(tmin) : 0.);
    c.e2 = ok ? rt_f32_up(e2) : INFINITY;
    c.tmin = (float)tmin;
}

/* The same ray queried with another t_min (a ConstantMedium asks its boundary from -inf, then from t1 + 0.0001): the slab terms of
 * e2 do not depend on t_min, so adding the new t_min's share to the old margin -- which still holds the old t_min's -- gives a
 * valid, slightly larger margin without the three reciprocals of cull_ray_setup.  Larger only means more RT_CULL_UNSURE. */
RT_CULL_HD void cull_ray_retime(CullRay &c, double tmin) {
    double e2 = (double)c.e2 + (fabs(tmin) <= 1.7e308 ? 2.6e-7 * fabs(tmin) : 0.);
    c.e2 = tmin == tmin ? rt_f32_up(e2) : INFINITY; /* (inf stays inf) */
    c.tmin = (float)tmin;
}

/* box = f32(min.xyz), f32(max.xyz) of the node;  bestL <= running best <= bestU */
RT_CULL_HD int cull_box(const CullRay &c, float b0, float b1, float b2, float b3, float b4, float b5, float bestL, float bestU) {
    float lx = fmaf(b0, c.ix, c.nx), hx = fmaf(b3, c.ix, c.nx);
    float ly = fmaf(b1, c.iy, c.ny), hy = fmaf(b4, c.iy, c.ny);
    float lz = fmaf(b2, c.iz, c.nz), hz = fmaf(b5, c.iz, c.nz);
    float m0 = fmaxf(fmaxf(fminf(lx, hx), fminf(ly, hy)), fmaxf(fminf(lz, hz), c.tmin));
    float m1 = fminf(fminf(fmaxf(lx, hx), fmaxf(ly, hy)), fmaxf(lz, hz));
    float gl = fminf(m1, bestL) - m0;
    float gu = fminf(m1, bestU) - m0;
    if (gl > c.e2) return RT_CULL_HIT;
    if (gu < -c.e2) return RT_CULL_MISS;
    return RT_CULL_UNSURE;
}

#endif /* RT_CULL_H */
