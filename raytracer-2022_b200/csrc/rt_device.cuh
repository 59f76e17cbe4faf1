/*
 * rt_device.cuh -- device-side restatement of the reference's geometry, materials, textures
 * and pdfs for sm_100a.  Everything that decides something is f64 with the reference's operation order,
 * compiled with -fmad=false (rustc never contracts a*b+c), so hit/miss decisions and
 * primitive ids are bit-identical to BvhNode::hit (hittable/bvh/mod.rs:86-101); the one f32 computation, the
 * pre-test of BvhNode boxes, only answers when the f64 answer is certain (rt_cull.h).
 *
 * Layout of the scene in HBM: see include/rtgpu.h for what the caller hands over.  On the device the traversal
 * program is re-packed into 32-byte UNITS (rtgpu_scene_create, rt_kernels.cu):
 *     BvhNode   1 unit : kind | ordinal << 8, skip, f32 box (min.xyz, max.xyz); its f64 box lives in boxes64[ordinal]
 *     any other 2 units: the caller's 64-byte op (two-slot ops: 4 units), skip re-mapped to units
 * A thread fetches units i and i+1 with four 128-bit loads: a whole leaf op, or a BvhNode together with whatever
 * follows it (RT_NODE_PAIR = 1 also tests that unit when it is a BvhNode too: measured slower, off).  Boxes are decided
 * by the f32 pre-test of rt_cull.h and, when that is not certain, by the f64 test on the f64 box.
 */
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/rt_rng.h"
#include "../../include/rtgpu.h"
#include "rt_cull.h"
#include "rt_div.h"
#include "rt_trig.h"

#define RT_NO_HIT 0xffffffffu
#ifndef RT_THREADS
#define RT_THREADS 128 /* threads per block of the plain grid-stride kernels */
#endif
#define RT_MAX_PUSH 4   /* nesting of PUSH ops (instances inside instances) */
#define RT_MAX_CHAIN 8  /* wrappers around one primitive */
#define RT_CLASS_ISOTROPIC 5u  /* shade classes: see k_classify in rt_kernels.cu */
#define RT_KIND_PASS_END 13u /* in registers only: the cursor of a lane that walks a medium's boundary in the warp loop reached its end */
#define RT_KIND_UNSURE_NODE 12u /* in registers only: a fetched NODE whose box the f32 pre-test could not decide */
#define RT_PI 3.14159265358979323846264338327950288
#define RT_F64_MAX 1.7976931348623157e308

struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 v3(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, double s) { return V3{rt_div_zero_fast(a.x, s), rt_div_zero_fast(a.y, s), rt_div_zero_fast(a.z, s)}; } /* rt_div.h */
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ double length_sqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double length(V3 a) { return sqrt(length_sqr(a)); }
__device__ __forceinline__ V3 to_unit(V3 a) { return a / length(a); }

struct Ray {
    V3 o, d;
    double tm;
};
__device__ __forceinline__ V3 ray_at(const Ray &r, double t) { return r.o + r.d * t; }

/* device view of a scene (pointers into HBM) */
struct DScene {
    const uint4 *prog;      /* 32-byte units, two uint4 each */
    const double2 *boxes64; /* f64 box of BvhNode `ordinal`: three double2 */
    const rtgpu_xform *xforms;
    const uint32_t *chain;
    const rtgpu_ctx *ctxs;
    const rtgpu_material *mats;
    const rtgpu_texture *texs;
    const rtgpu_perlin *perlins;
    const rtgpu_image *images;
    const uint8_t *texels;
    const rtgpu_light *lights;
    uint32_t n_units, n_lights, n_light_roots;
    float extent; /* largest |coordinate| of any BvhNode box (>= RT_CULL_MIN_EXTENT): the B of rt_cull.h */
    rtgpu_camera cam;
    double background[3];
};

/* What the out-of-line helpers need of the scene, BY VALUE: a `const DScene &` of the kernel's parameter block makes the compiler
 * copy the whole block (camera included) into local memory just to have an address to pass. */
struct TexTables {
    const rtgpu_texture *texs;
    const rtgpu_perlin *perlins;
    const rtgpu_image *images;
    const uint8_t *texels;
};
__device__ __forceinline__ TexTables tex_of(const DScene &sc) { return TexTables{sc.texs, sc.perlins, sc.images, sc.texels}; }
struct LightTable {
    const rtgpu_light *lights;
    uint32_t n_lights, n_light_roots;
};
__device__ __forceinline__ LightTable lights_of(const DScene &sc) { return LightTable{sc.lights, sc.n_lights, sc.n_light_roots}; }

/* ---- 128-bit loads of an op ------------------------------------------------------------- */
__device__ __forceinline__ const uint4 *unit_ptr(const uint4 *prog, uint32_t unit) { return prog + 2 * (size_t)unit; }
__device__ __forceinline__ uint4 load_op_header(const uint4 *op) { return __ldg(op); }
__device__ __forceinline__ void load_op_payload(const uint4 *op, double d[6]) {
    const double2 *p = reinterpret_cast<const double2 *>(op) + 1;
    double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y; d[4] = c.x; d[5] = c.y;
}
__device__ __forceinline__ void load_slot8(const uint4 *op, double d[8]) { /* second slot of a 2-slot op */
    const double2 *p = reinterpret_cast<const double2 *>(op);
    double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), e = __ldg(p + 3);
    d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y; d[4] = c.x; d[5] = c.y; d[6] = e.x; d[7] = e.y;
}

/* ---- wrappers: ray into object space (hittable/mod.rs:166,240-248,322) -------------------- */
__device__ __forceinline__ void xform_ray(const rtgpu_xform &xf, V3 &o, V3 &d) {
    switch (xf.kind) {
    case RTGPU_XF_TRANSLATE:
        o = o - v3(xf.p[0], xf.p[1], xf.p[2]);
        break;
    case RTGPU_XF_ROTATE_Y: {
        double s = xf.p[0], c = xf.p[1];
        double ox = c * o.x - s * o.z, oz = s * o.x + c * o.z;
        double dx = c * d.x - s * d.z, dz = s * d.x + c * d.z;
        o.x = ox; o.z = oz;
        d.x = dx; d.z = dz;
        break;
    }
    case RTGPU_XF_ZOOM:
        o = o / xf.p[0]; /* direction left unscaled (Q5) */
        break;
    default: /* FLIP: the ray is untouched */
        break;
    }
}
/* the three arrays traversal needs, passed by value (kept in registers / the param bank) */
struct TravScene {
    const uint4 *prog;
    const double2 *boxes64;
    const rtgpu_xform *xforms;
    const uint32_t *chain;
    float extent;
    uint32_t end; /* the walk stops at this unit: n_units, or the end of a ConstantMedium's boundary (kept here, i.e. in the parameter
                     bank, rather than in a register of every lane) */
};
__device__ __forceinline__ TravScene trav_of(const DScene &sc) { return TravScene{sc.prog, sc.boxes64, sc.xforms, sc.chain, sc.extent, sc.n_units}; }
__device__ __forceinline__ rtgpu_xform load_xform(const TravScene &sc, uint32_t chain_idx) {
    const rtgpu_xform *x = sc.xforms + __ldg(sc.chain + chain_idx);
    const double2 *p = reinterpret_cast<const double2 *>(x);
    rtgpu_xform r;
    double2 a = __ldg(p), b = __ldg(p + 1);
    uint2 hdr = *reinterpret_cast<uint2 *>(&a.x);
    r.kind = hdr.x;
    r.gid = hdr.y;
    r.p[0] = a.y; r.p[1] = b.x; r.p[2] = b.y;
    return r;
}

#ifndef RT_LEAF_INLINE
#define RT_LEAF_INLINE __forceinline__
#endif
/* ---- primitive tests: return true and shrink `best` on a hit ------------------------------- */
/* Sphere::hit, sphere.rs:39-56 (decision part; the record is rebuilt later for the winner) */
/* The two queries ConstantMedium::hit makes of a Sphere boundary (constantmedium.rs:50-58): hit(r, -inf, inf) -> t1, then
 * hit(r, t1 + 0.0001, inf) -> t2.  Both evaluate the same discriminant, square root and first root from the same operands, hence the
 * same values: they are formed once.  The comparisons are the ones Sphere::hit makes, in its order (a NaN root passes them, as there). */
__device__ __forceinline__ bool hit_sphere_twice(V3 o, V3 d, double a, V3 c, double radius, double &t1, bool &h2, double &t2) {
    h2 = false;
    V3 oc = o - c;
    double half_b = dot(oc, d);
    double cc = length_sqr(oc) - radius * radius;
    double disc = half_b * half_b - a * cc;
    if (disc < 0.) return false;
    double sqrtd = sqrt(disc);
    const double lo = (-half_b - sqrtd) / a;
    /* first query: t_min = -inf, t_max = +inf */
    double root = lo;
    if (root < -CUDART_INF || CUDART_INF < root) {
        root = (-half_b + sqrtd) / a;
        if (root < -CUDART_INF || CUDART_INF < root) return false;
    }
    t1 = root;
    /* second query: t_min = t1 + 0.0001, t_max = +inf */
    const double tmin = t1 + 0.0001;
    root = lo;
    if (root < tmin || CUDART_INF < root) {
        root = (-half_b + sqrtd) / a;
        if (root < tmin || CUDART_INF < root) return true;
    }
    t2 = root;
    h2 = true;
    return true;
}
__device__ __forceinline__ bool hit_sphere(V3 o, V3 d, double a, V3 c, double radius, double tmin, double &best) {
    V3 oc = o - c;
    double half_b = dot(oc, d);
    double cc = length_sqr(oc) - radius * radius;
    double disc = half_b * half_b - a * cc;
    if (disc < 0.) return false;
    double sqrtd = sqrt(disc);
    double root = (-half_b - sqrtd) / a;
    if (root < tmin || best < root) {
        root = (-half_b + sqrtd) / a;
        if (root < tmin || best < root) return false;
    }
    best = root;
    return true;
}
/* XYRect/XZRect/YZRect::hit, aarect.rs:46-56,129-139,212-222.  AX = thin axis; (A,B) in-plane.
 * A NaN t passes the range test, as upstream (Q2). */
template <int AX>
__device__ __forceinline__ bool hit_rect(V3 o, V3 d, double a0, double a1, double b0, double b1, double k, double tmin,
                                         double &best) {
    double oa, ob, ok, da, db, dk;
    if (AX == 2) { oa = o.x; ob = o.y; ok = o.z; da = d.x; db = d.y; dk = d.z; }
    else if (AX == 1) { oa = o.x; ob = o.z; ok = o.y; da = d.x; db = d.z; dk = d.y; }
    else { oa = o.y; ob = o.z; ok = o.x; da = d.y; db = d.z; dk = d.x; }
    double t = (k - ok) / dk;
    if (t < tmin || t > best) return false;
    double a = oa + t * da;
    double b = ob + t * db;
    if (a < a0 || a > a1 || b < b0 || b > b1) return false;
    best = t;
    return true;
}
/* Ring::hit, ring.rs:34-46: annulus in the plane y = 0 */
__device__ __forceinline__ bool hit_ring(V3 o, V3 d, double dis_min, double dis_max, double tmin, double &best) {
    double t = -o.y / d.y;
    if (t != t || t < tmin || t > best) return false;
    V3 p = o + d * t;
    double dd = p.x * p.x + p.z * p.z;
    if (dd < dis_min || dd > dis_max) return false;
    best = t;
    return true;
}
/* Boxes::hit = HittableList::hit over six rects (boxes.rs:24-69,82-84; hittable/mod.rs:90-100) */
__device__ RT_LEAF_INLINE bool hit_box(V3 o, V3 d, const double p[6], double tmin, double &best, uint32_t &face) {
    bool any = false;
    if (hit_rect<2>(o, d, p[0], p[3], p[1], p[4], p[5], tmin, best)) { any = true; face = 0; }
    if (hit_rect<2>(o, d, p[0], p[3], p[1], p[4], p[2], tmin, best)) { any = true; face = 1; }
    if (hit_rect<1>(o, d, p[0], p[3], p[2], p[5], p[4], tmin, best)) { any = true; face = 2; }
    if (hit_rect<1>(o, d, p[0], p[3], p[2], p[5], p[1], tmin, best)) { any = true; face = 3; }
    if (hit_rect<0>(o, d, p[1], p[4], p[2], p[5], p[3], tmin, best)) { any = true; face = 4; }
    if (hit_rect<0>(o, d, p[1], p[4], p[2], p[5], p[0], tmin, best)) { any = true; face = 5; }
    return any;
}
/* Triangle::hit, triangle.rs:51-64 with inside() of :33-46; n is get_normal(), precomputed */
__device__ RT_LEAF_INLINE bool hit_triangle(V3 o, V3 d, V3 a, V3 b, V3 c, V3 n, double tmin, double &best) {
    double t = dot(d, n);
    t = dot(a - o, n) / t;
    if (t != t || t < tmin || t > best) return false;
    V3 p = o + d * t;
    if (!(dot(cross(c - a, p - a), cross(c - a, b - a)) >= 0.)) return false;
    if (!(dot(cross(a - b, p - b), cross(a - b, c - b)) >= 0.)) return false;
    if (!(dot(cross(b - c, p - c), cross(b - c, a - c)) >= 0.)) return false;
    best = t;
    return true;
}
/* AABB::hit, aabb.rs:15-32.  inv = 1/dir is what upstream recomputes per node.  Upstream
 * computes t0 from min, t1 from max and swaps them when inv < 0; picking the near/far plane by
 * the sign first yields the same two values without the swap.  One final `t_max <= t_min` test
 * is equivalent to the per-axis early exits (both bounds are monotone).  The comparisons keep
 * upstream's polarity, so NaN (0 * inf) leaves a bound unchanged. */
struct RaySigns {
    uint32_t m; /* bit 0/1/2: inv.x/y/z < 0 */
    __device__ __forceinline__ bool x() const { return m & 1u; }
    __device__ __forceinline__ bool y() const { return m & 2u; }
    __device__ __forceinline__ bool z() const { return m & 4u; }
};
__device__ __forceinline__ bool hit_aabb(V3 o, V3 inv, RaySigns sg, const double b[6], double tmin, double tmax) {
    double t0x = ((sg.x() ? b[3] : b[0]) - o.x) * inv.x, t1x = ((sg.x() ? b[0] : b[3]) - o.x) * inv.x;
    double t0y = ((sg.y() ? b[4] : b[1]) - o.y) * inv.y, t1y = ((sg.y() ? b[1] : b[4]) - o.y) * inv.y;
    double t0z = ((sg.z() ? b[5] : b[2]) - o.z) * inv.z, t1z = ((sg.z() ? b[2] : b[5]) - o.z) * inv.z;
    /* (a variant with 15 independent pairwise comparisons instead of the running max/min was tried:
     *  shorter dependency chain, but 16 DSETP instead of 7 on the half-rate FP64 pipe -- 9 % slower) */
    tmin = t0x > tmin ? t0x : tmin;
    tmax = t1x < tmax ? t1x : tmax;
    tmin = t0y > tmin ? t0y : tmin;
    tmax = t1y < tmax ? t1y : tmax;
    tmin = t0z > tmin ? t0z : tmin;
    tmax = t1z < tmax ? t1z : tmax;
    return !(tmax <= tmin);
}

static_assert(RT_MAX_PUSH >= 1, "push stack");
/* ---- closest hit over a program range ------------------------------------------------------ */
struct Closest {
    double t;      /* running best (= t_max when nothing was hit) */
    uint32_t op;   /* unit index of the winning op, RT_NO_HIT if none */
    uint32_t face; /* box side of the winner */
    uint32_t cls;  /* shade class of the winner's material: word 1 of a leaf op's header on the device (rt_kernels.cu, repack_program) */
};

/* Walks units [begin, end) in order -- the reference's left-then-right recursion unrolled
 * (SURVEY.md 2.5).  The loop is "while-while": a tight run over consecutive NODE units, then one
 * leaf op, so that the lanes of a warp spend most iterations in the same code.
 * LEVEL 0 is the walk of the world; LEVEL 1 the walk of a ConstantMedium's boundary, LEVEL 2 the boundary of a ConstantMedium that
 * sits inside another one's boundary (upstream's ConstantMedium<H> takes any Hittable, constantmedium.rs:14-43): a medium met on
 * LEVEL 1 draws from the same stream, in visiting order; deeper nesting is refused when the scene is created.
 * `rng` draws the medium's free path (constantmedium.rs:60) in visiting order.
 *
 * Traversal<RayRegs> is the single-ray walk with everything in registers: it runs the boundary of a ConstantMedium
 * that is not one sphere (closest_hit_boundary).  The main walk of k_extend is extend_cta_loop below, which keeps its
 * rays in shared memory and uses the same tests phase by phase. */
#define RT_MAX_MEDIUM_NESTING 2 /* a ConstantMedium inside the boundary of a ConstantMedium, and no deeper (kernels: NEST 1, 2) */
/* CONST_TMIN: t_min is the 0.001 of main.rs:243 for every ray (k_extend), which saves its two registers */
template <bool CONST_TMIN>
struct RayRegsT {
    V3 o_, d_;
    double tm_, tmin_;
    __device__ __forceinline__ V3 o() const { return o_; }
    __device__ __forceinline__ V3 d() const { return d_; }
    __device__ __forceinline__ double tm() const { return tm_; }
    __device__ __forceinline__ double tmin() const { return CONST_TMIN ? 0.001 : tmin_; }
    __device__ __forceinline__ void set_od(V3 o, V3 d) { o_ = o; d_ = d; }
    __device__ __forceinline__ void set_times(double tm, double tmin) {
        tm_ = tm;
        if (!CONST_TMIN) tmin_ = tmin;
    }
    __device__ __forceinline__ void set_tmin(double tmin) {
        if (!CONST_TMIN) tmin_ = tmin;
    }
};
typedef RayRegsT<false> RayRegs;
#ifndef RT_NODE_FETCH32
#define RT_NODE_FETCH32 0 /* 1: fetch 32 bytes for a NODE and the other 32 only for leaf ops -- fewer L1 data-pipe cycles (50 % -> 35 %)
                             but the dependent second fetch lengthens every leaf: 8 % slower on the B200 */
#endif
#ifndef RT_NODE_PAIR
#define RT_NODE_PAIR 0 /* with 64-byte fetches: also test the NODE that follows a NODE under the same fetch (2 % slower: the
                          second test is wasted whenever the first box is missed) */
#endif

/* Where a walk draws its uniforms from: a VALUE type with `double operator()()` and a stream position `pos`.  It is handed to
 * the out-of-line boundary walk by value and the position comes back in the result: taking the address of the caller's state
 * would push it into local memory for the whole kernel. */
/* result of a boundary walk: t and op as usual; `face` carries the stream position after the walk (a boundary's box side is
 * never needed), so that the result stays a plain register-returned Closest */
typedef Closest BoundaryHit;
/* a fully resolved Philox slot stream (include/rt_rng.h): 24 bytes, cheap to pass by value */
struct StreamDraw {
    uint64_t seed;
    uint32_t s0, s1, domain, pos;
    __device__ __forceinline__ double operator()() {
        rt_rng g;
        g.init(seed, s0, s1, domain, pos);
        double u = g.gen();
        pos = g.pos;
        return u;
    }
    __device__ __forceinline__ StreamDraw stream() const { return *this; }
};
/* The out-of-line boundary walk of the SINGLE-RAY walkers (k_tail, and itself for a medium inside a medium's boundary): NEST = how
 * many levels of boundaries the instantiation walks (LEVEL < NEST calls LEVEL + 1).  The warp loop of k_extend does not use it: a
 * call there shapes the register allocation of the whole kernel, so its lanes walk boundaries themselves
 * (Traversal<.., WALK_IN_LOOP>). */
#ifndef RT_BOUNDARY_CALL
#define RT_BOUNDARY_CALL __noinline__
#endif
template <int LEVEL, int NEST>
__device__ RT_BOUNDARY_CALL BoundaryHit closest_hit_boundary(const TravScene sc, uint32_t begin, uint32_t end, V3 o, V3 d, double tm,
                                                         double tmin, double tmax, StreamDraw ds);

/* phases of the warp scheduler: ops whose code is the same run together */
enum { PH_NODE = 0, PH_SPHERE = 1, PH_BOX = 2, PH_TRI = 3, PH_MISC = 4, PH_EXACT = 5, PH_NONE = 7 };
__device__ __forceinline__ uint32_t phase_of_kind(uint32_t kind) {
    /* NODE SPHERE MSPHERE RECT_XY RECT_XZ RECT_YZ BOX TRI MEDIUM PUSH POP RING UNSURE_NODE PASS_END, 3 bits each */
    return (uint32_t)((045444432444110ull >> (3 * kind)) & 7ull);
}

struct SavedRays {
    double v[RT_MAX_PUSH][6];
    float c[RT_MAX_PUSH][8]; /* the outer ray's CullRay: cheaper to keep than three divisions at the POP */
};
/* What a lane of the warp loop puts away while it walks the boundary of a ConstantMedium (Traversal<RS, true>): the outer walk's
 * running best, t_min and range end, and what ConstantMedium::hit needs once both boundary hits are known.  Indexed by the
 * nesting depth, hence in local memory like SavedRays; touched twice per walked medium. */
struct MediumFrames {
    double best_t[RT_MAX_MEDIUM_NESTING], tmin[RT_MAX_MEDIUM_NESTING], t1[RT_MAX_MEDIUM_NESTING], neg_inv_density[RT_MAX_MEDIUM_NESTING];
    float cull_e2[RT_MAX_MEDIUM_NESTING], cull_tmin[RT_MAX_MEDIUM_NESTING]; /* the outer walk's margin and t_min of the box pre-test */
    uint32_t best_op[RT_MAX_MEDIUM_NESTING], best_face_cls[RT_MAX_MEDIUM_NESTING], lim[RT_MAX_MEDIUM_NESTING], op[RT_MAX_MEDIUM_NESTING],
        skip[RT_MAX_MEDIUM_NESTING], pass[RT_MAX_MEDIUM_NESTING];
};

__device__ __forceinline__ float lo_f32(double x) { return __int_as_float(__double2loint(x)); }
__device__ __forceinline__ float hi_f32(double x) { return __int_as_float(__double2hiint(x)); }

/* Resumable in-order walk of units [begin, end).  Units i and i+1 (64 B = four 128-bit loads issued
 * together) are fetched as soon as the index is known; exec_*() consume them.
 * One call of exec handles one op, so a warp can run "all lanes that sit on a NODE" together,
 * then "all lanes that sit on a sphere", ... (see extend_warp_loop). */
/* WALK_IN_LOOP (the warp loop of scenes with walked media): a ConstantMedium whose boundary is not one Sphere does not call
 * closest_hit_boundary; the lane puts its outer best / t_min away (MediumFrames), walks the boundary's ops through the very same
 * NODE and leaf rounds -- once from -inf, once from t1 + 0.0001 (constantmedium.rs:50-58) -- and picks the outer walk up again.
 * No call in the loop (what a call costs there: DESIGN.md 5), and the boundary's boxes, rects and wrappers run warp-wide. */
template <typename RS, bool WALK_IN_LOOP = false>
struct Traversal {
    RS ray;             /* f64 origin, direction, shutter time, t_min */
    CullRay c;          /* f32 view for the box pre-test (rt_cull.h) */
    float bestL, bestU; /* bestL <= best.t <= bestU */
    Closest best;
    uint32_t i;
    int sp;
    uint32_t lim; /* WALK_IN_LOOP: end of the boundary being walked (0xffffffff outside of one) */
    int md;       /* WALK_IN_LOOP: boundaries being walked (depth of MediumFrames) */
    uint4 h;     /* bytes 0..15 of unit i: header of a leaf op / kind|ordinal, skip, box min.xy of a NODE */
    double p[6]; /* bytes 16..63: payload of a leaf op / rest of the NODE box, then unit i+1 */
    /* the PUSH/POP stack lives OUTSIDE this struct (SavedRays): it is indexed dynamically, and an
     * aggregate with a dynamically indexed member is kept in local memory as a whole */

    /* fetched as soon as the index is known: the L1/L2 latency of the next op overlaps with the warp
     * vote and with the other lanes' work */
    __device__ __forceinline__ void fetch_op(const TravScene sc) {
        if constexpr (WALK_IN_LOOP) {
            if (i == lim) { /* ops of a boundary never jump past its end: the pass is over */
                h.x = RT_KIND_PASS_END;
                return;
            }
        }
        if (i < sc.end) {
            const uint4 *op = unit_ptr(sc.prog, i);
            h = load_op_header(op);
            const double2 *q = reinterpret_cast<const double2 *>(op) + 1;
            double2 a = __ldg(q);
            p[0] = a.x;
            p[1] = a.y;
#if RT_NODE_FETCH32
            /* a NODE is one 32-byte unit: two 128-bit loads.  Only a leaf op needs units i and i+1; its second half is
             * asked for once the kind is known (every 128-bit load of a warp costs four L1 data-pipe cycles whatever
             * the number of lanes, and ~25 of the ~31 ops of a ray are boxes) */
            if ((h.x & 0xffu) != RTGPU_OP_NODE)
#endif
            {
                double2 b = __ldg(q + 1), c2 = __ldg(q + 2);
                p[2] = b.x; p[3] = b.y; p[4] = c2.x; p[5] = c2.y;
            }
        }
    }
    __device__ __forceinline__ void set_best_bounds() {
        bestL = rt_f32_down(best.t);
        bestU = rt_f32_up(best.t);
    }
    __device__ __forceinline__ void derive(V3 o, V3 d, double tmin, float extent) {
        /* 1/d: what AABB::hit divides per node (aabb.rs:17) */
        cull_ray_setup(c, o.x, o.y, o.z, 1. / d.x, 1. / d.y, 1. / d.z, tmin, extent);
    }
    __device__ __forceinline__ void start(const TravScene sc, uint32_t begin, V3 o, V3 d, double tm_, double tmin_, double tmax) {
        best.t = tmax;
        best.op = RT_NO_HIT;
        best.face = 0;
        best.cls = 0;
        ray.set_od(o, d);
        ray.set_times(tm_, tmin_);
        derive(o, d, tmin_, sc.extent);
        set_best_bounds();
        i = begin;
        sp = 0;
        lim = 0xffffffffu;
        md = 0;
        fetch_op(sc);
    }
    __device__ __forceinline__ bool done(const TravScene sc) const {
        if constexpr (WALK_IN_LOOP) return i >= sc.end && md == 0; /* a boundary may end where the program ends */
        else return i >= sc.end;
    }
    /* WALK_IN_LOOP: (re)starts a pass over the boundary [begin, lim) with a fresh best and the given t_min */
    __device__ __forceinline__ void start_pass(const TravScene sc, uint32_t begin, double tmin) {
        best.t = CUDART_INF;
        best.op = RT_NO_HIT;
        best.face = 0;
        best.cls = 0;
        ray.set_tmin(tmin);
        cull_ray_retime(c, tmin);
        set_best_bounds();
        i = begin;
        fetch_op(sc);
    }
    /* the MEDIUM op under the cursor has a boundary to walk: put the outer walk away and start the first pass */
    __device__ __forceinline__ void begin_boundary(const TravScene sc, MediumFrames &mf) {
        mf.best_t[md] = best.t;
        mf.best_op[md] = best.op;
        mf.best_face_cls[md] = best.face | (best.cls << 8);
        mf.tmin[md] = ray.tmin();
        mf.cull_e2[md] = c.e2;
        mf.cull_tmin[md] = c.tmin;
        mf.lim[md] = lim;
        mf.op[md] = i;
        mf.skip[md] = h.y;
        mf.neg_inv_density[md] = p[0];
        mf.pass[md] = 1;
        ++md;
        lim = h.y;
        start_pass(sc, i + 2, -CUDART_INF);
    }
    /* the cursor reached the end of the boundary (kind PASS_END) */
    template <typename DS>
    __device__ __forceinline__ void end_pass(const TravScene sc, MediumFrames &mf, DS &draw) {
        const int k = md - 1;
        const bool first = mf.pass[k] == 1;
        if (first && best.op != RT_NO_HIT) { /* rec1 found: look for rec2 from rec1.t + 0.0001 */
            mf.t1[k] = best.t;
            mf.pass[k] = 2;
            start_pass(sc, mf.op[k] + 2, best.t + 0.0001);
            return;
        }
        const bool h1 = !first, h2 = !first && best.op != RT_NO_HIT;
        const double t1 = h1 ? mf.t1[k] : CUDART_INF, t2 = best.t;
        /* back to the outer walk */
        best.t = mf.best_t[k];
        best.op = mf.best_op[k];
        best.face = mf.best_face_cls[k] & 0xffu;
        best.cls = mf.best_face_cls[k] >> 8;
        ray.set_tmin(mf.tmin[k]);
        c.e2 = mf.cull_e2[k];
        c.tmin = mf.cull_tmin[k];
        lim = mf.lim[k];
        md = k;
        i = mf.op[k];
        p[0] = mf.neg_inv_density[k];
        const bool hit = medium_decide(length_sqr(ray.d()), ray.tmin(), h1, t1, h2, t2, draw);
        i = mf.skip[k];
        (void)hit;
        set_best_bounds();
        fetch_op(sc);
    }
    __device__ __forceinline__ uint32_t phase() const { return phase_of_kind(h.x & 0xffu); }
    __device__ __forceinline__ bool at_node() const { return (h.x & 0xffu) == RTGPU_OP_NODE; } /* = phase() == PH_NODE, two instructions */

    /* BvhNode::hit's box test (bvh/mod.rs:87) for the NODE at unit i and -- same fetch -- for the NODE that
     * follows it, if any: that is its left child, tested next anyway when the first box is hit.  Both answers
     * come from the f32 pre-test; a box it cannot decide is left for exec_node_exact (the lane changes phase,
     * so the f64 code exists once, outside the NODE rounds). */
    __device__ __forceinline__ void exec_node(const TravScene sc) {
        int ra = cull_box(c, __uint_as_float(h.z), __uint_as_float(h.w), lo_f32(p[0]), hi_f32(p[0]), lo_f32(p[1]), hi_f32(p[1]),
                          bestL, bestU);
        uint32_t bw = (uint32_t)__double2loint(p[2]), bskip = (uint32_t)__double2hiint(p[2]);
        int rb = cull_box(c, lo_f32(p[3]), hi_f32(p[3]), lo_f32(p[4]), hi_f32(p[4]), lo_f32(p[5]), hi_f32(p[5]), bestL, bestU);
        if (ra == RT_CULL_UNSURE) {
            h.x |= RT_KIND_UNSURE_NODE; /* (kind NODE is 0) the fetched unit now reads as "a NODE that needs the f64 test": phase PH_EXACT */
            return;
        }
        uint32_t nxt = i + 1; /* first box hit; what follows is a leaf, or a NODE the pre-test could not decide */
        if (ra == RT_CULL_MISS) nxt = h.y;
#if RT_NODE_PAIR
        else if ((bw & 0xffu) == RTGPU_OP_NODE && rb != RT_CULL_UNSURE) nxt = rb == RT_CULL_MISS ? bskip : i + 2;
#endif
        i = nxt;
        fetch_op(sc);
    }
    /* AABB::hit (aabb.rs:15-32) in f64 on the f64 box of the NODE at unit i */
    __device__ __forceinline__ void exec_node_exact(const TravScene sc) {
        const double2 *q = sc.boxes64 + 3 * (size_t)(h.x >> 8);
        double2 b0 = __ldg(q), b1 = __ldg(q + 1), b2 = __ldg(q + 2);
        double b[6] = {b0.x, b0.y, b1.x, b1.y, b2.x, b2.y};
        V3 d = ray.d();
        V3 inv = v3(1. / d.x, 1. / d.y, 1. / d.z);
        RaySigns sg;
        sg.m = (inv.x < 0. ? 1u : 0u) | (inv.y < 0. ? 2u : 0u) | (inv.z < 0. ? 4u : 0u);
        i = hit_aabb(ray.o(), inv, sg, b, ray.tmin(), best.t) ? i + 1 : h.y;
        fetch_op(sc);
    }
    __device__ __forceinline__ void exec_sphere(const TravScene sc) {
        V3 c0 = v3(p[0], p[1], p[2]);
        uint32_t width = 2;
        if ((h.x & 0xffu) == RTGPU_OP_MSPHERE) { /* MovingSphere::center, sphere.rs:124-127 */
            double q[8];
            load_slot8(unit_ptr(sc.prog, i + 2), q);
            c0 = c0 + (v3(q[0], q[1], q[2]) - c0) * ((ray.tm() - p[4]) / (p[5] - p[4]));
            width = 4;
        }
        V3 d = ray.d();
        if (hit_sphere(ray.o(), d, length_sqr(d), c0, p[3], ray.tmin(), best.t)) {
            best.op = i;
            best.cls = h.y;
            set_best_bounds();
        }
        i += width;
        fetch_op(sc);
    }
    __device__ __forceinline__ void exec_box(const TravScene sc) {
        if (hit_box(ray.o(), ray.d(), p, ray.tmin(), best.t, best.face)) {
            best.op = i;
            best.cls = h.y;
            set_best_bounds();
        }
        i += 2;
        fetch_op(sc);
    }
    __device__ __forceinline__ void exec_tri(const TravScene sc) {
        double q[8];
        load_slot8(unit_ptr(sc.prog, i + 2), q);
        if (hit_triangle(ray.o(), ray.d(), v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]),
                         ray.tmin(), best.t)) {
            best.op = i;
            best.cls = h.y;
            set_best_bounds();
        }
        i += 4;
        fetch_op(sc);
    }
    /* ConstantMedium::hit once the two boundary hits are known (constantmedium.rs:60-83); a = |d|^2, p[0] = -1/density, i = the op */
    template <typename DS>
    __device__ __forceinline__ bool medium_decide(double a, double tmin, bool h1, double t1, bool h2, double t2, DS &draw) {
        if (h1 && h2) {
            t1 = fmax(t1, tmin);
            t2 = fmin(t2, best.t);
            if (!(t1 >= t2)) {
                t1 = fmax(t1, 0.);
                double ray_length = sqrt(a);
                double distance_inside_boundary = (t2 - t1) * ray_length;
                double rnd = draw();
                double hit_distance = p[0] * log(rnd); /* neg_inv_density * rnd.log(E) */
                if (!(hit_distance > distance_inside_boundary)) {
                    best.t = t1 + hit_distance / ray_length;
                    best.op = i;
                    best.cls = RT_CLASS_ISOTROPIC; /* a ConstantMedium scatters with its Isotropic material (constantmedium.rs:77) */
                    return true;
                }
            }
        }
        return false;
    }
    /* rects, ConstantMedium, PUSH, POP.  DRAW() = next uniform of the ray's stream */
    /* MESH = false: the program holds no Triangle and no Ring (checked at rtgpu_scene_create): their code is left out */
    template <int LEVEL, int NEST, typename DS, bool MESH = true>
    __device__ __forceinline__ void exec_misc(const TravScene sc, SavedRays &stk, DS &draw, MediumFrames *mf = nullptr) {
        bool hit = false;
        const uint32_t kind = h.x & 0xffu;
        if constexpr (WALK_IN_LOOP) {
            if (kind == RT_KIND_PASS_END) {
                end_pass(sc, *mf, draw);
                return;
            }
        }
        if (kind == RTGPU_OP_POP) {
            --sp;
            ray.set_od(v3(stk.v[sp][0], stk.v[sp][1], stk.v[sp][2]), v3(stk.v[sp][3], stk.v[sp][4], stk.v[sp][5]));
            c.ix = stk.c[sp][0]; c.iy = stk.c[sp][1]; c.iz = stk.c[sp][2];
            c.nx = stk.c[sp][3]; c.ny = stk.c[sp][4]; c.nz = stk.c[sp][5];
            c.e2 = stk.c[sp][6]; c.tmin = stk.c[sp][7];
            i += 2;
            fetch_op(sc);
            return;
        }
        V3 o = ray.o(), d = ray.d();
        const double tmin = ray.tmin();
        switch (kind) {
        case RTGPU_OP_RECT_XY:
            hit = hit_rect<2>(o, d, p[0], p[1], p[2], p[3], p[4], tmin, best.t);
            if (hit) { best.op = i; best.cls = h.y; }
            i += 2;
            break;
        case RTGPU_OP_RECT_XZ:
            hit = hit_rect<1>(o, d, p[0], p[1], p[2], p[3], p[4], tmin, best.t);
            if (hit) { best.op = i; best.cls = h.y; }
            i += 2;
            break;
        case RTGPU_OP_RECT_YZ:
            hit = hit_rect<0>(o, d, p[0], p[1], p[2], p[3], p[4], tmin, best.t);
            if (hit) { best.op = i; best.cls = h.y; }
            i += 2;
            break;
        case RTGPU_OP_RING:
            if constexpr (MESH) {
                hit = hit_ring(o, d, p[0], p[1], tmin, best.t);
                if (hit) { best.op = i; best.cls = h.y; }
            }
            i += 2;
            break;
        case RTGPU_OP_MEDIUM:
            if constexpr (LEVEL < NEST || LEVEL == 0) { /* ConstantMedium::hit, constantmedium.rs:49-83 */
                double t1 = CUDART_INF, t2 = CUDART_INF;
                bool h1, h2 = false;
                const double a = length_sqr(d);
                bool single_sphere = h.y == i + 4;
                if (single_sphere) single_sphere = (load_op_header(unit_ptr(sc.prog, i + 2)).x & 0xffu) == RTGPU_OP_SPHERE;
                if (single_sphere) { /* boundary = one Sphere (both media of final_scene) */
                    double bp[6];
                    load_op_payload(unit_ptr(sc.prog, i + 2), bp);
                    V3 c0 = v3(bp[0], bp[1], bp[2]);
                    h1 = hit_sphere_twice(o, d, a, c0, bp[3], t1, h2, t2);
                } else if constexpr (WALK_IN_LOOP) {
                    begin_boundary(sc, *mf);
                    return;
                } else if constexpr (LEVEL < NEST) {
                    const double tm = ray.tm();
                    BoundaryHit r1 = closest_hit_boundary<LEVEL + 1, NEST>(sc, i + 2, h.y, o, d, tm, -CUDART_INF, CUDART_INF, draw.stream());
                    draw.pos = r1.face;
                    h1 = r1.op != RT_NO_HIT;
                    t1 = r1.t;
                    if (h1) {
                        BoundaryHit r2 = closest_hit_boundary<LEVEL + 1, NEST>(sc, i + 2, h.y, o, d, tm, t1 + 0.0001, CUDART_INF, draw.stream());
                        draw.pos = r2.face;
                        h2 = r2.op != RT_NO_HIT;
                        t2 = r2.t;
                    }
                } else {
                    h1 = false; /* NEST 0: rtgpu_scene_create found every boundary to be one Sphere */
                }
                hit = medium_decide(a, tmin, h1, t1, h2, t2, draw);
            }
            i = h.y;
            break;
        default: { /* PUSH */
            stk.v[sp][0] = o.x; stk.v[sp][1] = o.y; stk.v[sp][2] = o.z;
            stk.v[sp][3] = d.x; stk.v[sp][4] = d.y; stk.v[sp][5] = d.z;
            stk.c[sp][0] = c.ix; stk.c[sp][1] = c.iy; stk.c[sp][2] = c.iz;
            stk.c[sp][3] = c.nx; stk.c[sp][4] = c.ny; stk.c[sp][5] = c.nz;
            stk.c[sp][6] = c.e2; stk.c[sp][7] = c.tmin;
            ++sp;
            for (uint32_t j = 0; j < h.w; ++j) {
                rtgpu_xform xf = load_xform(sc, h.z + j);
                xform_ray(xf, o, d);
            }
            ray.set_od(o, d);
            derive(o, d, tmin, (float)p[0]); /* p[0]: the extent of the boxes under this wrapper, in the child's space (repack_program) */
            i += 2;
            break;
        }
        }
        if (hit) set_best_bounds();
        fetch_op(sc);
    }
    /* one op, whatever it is (single-ray loops) */
    template <int LEVEL, int NEST, typename DS>
    __device__ __forceinline__ void step(const TravScene sc, SavedRays &stk, DS &draw) {
        switch (phase()) {
        case PH_NODE: exec_node(sc); break;
        case PH_EXACT: exec_node_exact(sc); break;
        case PH_SPHERE: exec_sphere(sc); break;
        case PH_BOX: exec_box(sc); break;
        case PH_TRI: exec_tri(sc); break;
        default: exec_misc<LEVEL, NEST>(sc, stk, draw); break;
        }
    }
};

/* boundary walk of a medium whose boundary is not a single sphere: rare, kept out of line */
template <int LEVEL, int NEST>
__device__ RT_BOUNDARY_CALL BoundaryHit closest_hit_boundary(const TravScene sc, uint32_t begin, uint32_t end, V3 o, V3 d, double tm,
                                                         double tmin, double tmax, StreamDraw ds) {
    Traversal<RayRegs> tr;
    SavedRays stk;
    TravScene sub = sc;
    sub.end = end;
    tr.start(sub, begin, o, d, tm, tmin, tmax);
    while (!tr.done(sub)) tr.template step<LEVEL, NEST>(sub, stk, ds);
    BoundaryHit r = tr.best;
    r.face = ds.pos;
    return r;
}
/* the stream of the ray a lane of k_extend / k_trace_extend owns */
template <typename IO>
struct LaneDraw {
    const IO &io; /* never leaves the (inlined) warp loop: the boundary walk gets stream(), a value */
    uint32_t a, b, pos;
    __device__ __forceinline__ double operator()() { return io.draw(a, b, pos); }
    __device__ __forceinline__ StreamDraw stream() const { return io.stream(a, b, pos); }
};

#ifndef RT_EXTEND_REFILL
#define RT_EXTEND_REFILL 16 /* refill when at least this many lanes of the warp are idle (a refill stalls the whole warp on the
                               queue load, three reciprocals and the first op: 16 is 2.7 % faster than 8; 20+ loses lanes) */
#endif
#ifndef RT_MEDIUM_QUORUM
#define RT_MEDIUM_QUORUM 8 /* lanes that must stand on a ConstantMedium before it runs (1 = no hold) */
#endif
#ifndef RT_PREFETCH_AHEAD
#define RT_PREFETCH_AHEAD 16384 /* rays ahead of the fetch counter pulled into L2 at every refill (0 = off; +1 %) */
#endif
#ifndef RT_EXTEND_REFILL_SORTED
#define RT_EXTEND_REFILL_SORTED 24 /* the same when the rays come in sorted order: every refill mixes rays of another region into the
                                      warp, so it pays to drain further (24: +1.7 % over 16, 8: -7 %, 32: -3 %) */
#endif
#ifndef RT_NODE_UNROLL
#define RT_NODE_UNROLL 3 /* boxes a lane may test under one vote (A/B on the B200, 192-spp steps: 3 is +1.3 % over 2; 4 is -1 %) */
#endif
#ifndef RT_NODE_QUORUM_PCT
#define RT_NODE_QUORUM_PCT 35 /* keep running NODE rounds while at least this share of the live lanes sits on a NODE */
#endif

/* The extend loop of one warp (persistent): every lane owns one ray at a time.
 *  - dynamic fetch: a lane whose ray is finished takes the next ray index from a global counter
 *    (one atomic per refill of the warp), so trip-count differences do not idle lanes;
 *  - NODE-first scheduling: while enough lanes sit on a BvhNode box only those run (uniform code, one
 *    ballot); lanes that reached a leaf wait and then run their leaf op in plain SIMT.
 * Each ray still executes its own ops strictly in program order with its own running best, so
 * results are identical to the sequential walk.
 * IO: load(idx, o, d, tm, tmin, tmax, a, b, pos, cur) -> the ray, the identity (a, b) and position of its random
 *     stream and where its result goes; draw(a, b, pos) -> next uniform of that stream; store(cur, Closest, pos).
 * (Two CTA-level designs were built on this walk and measured on the B200 -- leaf requests served warp-wide through
 *  shared-memory queues, and a full pool of rays in shared memory with one queue per phase: both bit-exact, both
 *  slower (1 095 and 1 606 against 2 235 Mrays/s): the queue upkeep costs more instructions and L1/shared-memory
 *  data-pipe cycles than the fuller warps save.  They are in the history of this file, not in the tree.) */
/* MESH = false for programs without Triangle and Ring ops (every book scene): leaving their code out of the kernel makes it 13 %
 * smaller and halves its spills -- k_extend on C3 64.1 -> 61.2 ms. */
template <int NEST, bool MESH, typename IO>
__device__ __forceinline__ void extend_warp_loop(const TravScene ts, uint32_t n_units, uint32_t n, uint32_t *fetch, IO &io) {
    const uint32_t lane = threadIdx.x & 31u;
    constexpr bool kWalkInLoop = NEST >= 1; /* scenes with walked media: boundaries are walked by this loop, t_min varies */
    Traversal<RayRegsT<IO::kConstTmin && !kWalkInLoop>, kWalkInLoop> tr;
    SavedRays stk;
    MediumFrames mf;
    bool active = false, exhausted = false;
    uint32_t rng_a = 0, rng_b = 0, rng_pos = 0, cur = 0;
    const int refill_at = io.refill_at();
    for (;;) {
        uint32_t idle = __ballot_sync(0xffffffffu, !active);
        if (!exhausted && (idle == 0xffffffffu || __popc(idle) >= refill_at)) {
            int leader = __ffs(idle) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(fetch, (uint32_t)__popc(idle));
            base = __shfl_sync(0xffffffffu, base, leader);
            exhausted = base + (uint32_t)__popc(idle) >= n;
            if (!active) {
                uint32_t idx = base + (uint32_t)__popc(idle & ((1u << lane) - 1u));
#if RT_PREFETCH_AHEAD > 0
                if (idx + RT_PREFETCH_AHEAD < n) io.prefetch(idx + RT_PREFETCH_AHEAD);
#endif
                if (idx < n) {
                    V3 o, d;
                    double tm, tmin, tmax;
                    io.load(idx, o, d, tm, tmin, tmax, rng_a, rng_b, rng_pos, cur);
                    tr.start(ts, 0, o, d, tm, tmin, tmax);
                    active = !tr.done(ts);
                    if (!active) io.store(cur, tr.best, rng_pos);
                }
            }
            idle = __ballot_sync(0xffffffffu, !active);
        }
        if (idle == 0xffffffffu) break; /* no lane has a ray and none is left to fetch */
        /* NODE rounds: as long as enough lanes sit on a BvhNode box, run only those (uniform code,
         * one ballot of overhead); lanes that reached a leaf wait.  Then every waiting lane runs its
         * leaf op once (plain SIMT over the few leaf kinds) and the cycle repeats. */
        const uint32_t n_active = 32u - (uint32_t)__popc(idle);
        const uint32_t node_quorum = max(1u, (n_active * (uint32_t)((RT_NODE_QUORUM_PCT * 32 + 50) / 100)) >> 5); /* = RT_NODE_QUORUM_PCT % of the live lanes */
        bool node_ran = false;
#pragma unroll 1
        for (;;) {
            bool in_node = active && tr.at_node();
            uint32_t m = __ballot_sync(0xffffffffu, in_node);
            if ((uint32_t)__popc(m) < node_quorum) break;
            node_ran = true;
            if (in_node) {
                tr.exec_node(ts);
#if RT_NODE_UNROLL >= 2
                if (!tr.done(ts) && tr.at_node()) tr.exec_node(ts); /* second box under the same vote */
#endif
#if RT_NODE_UNROLL >= 3
                if (!tr.done(ts) && tr.at_node()) tr.exec_node(ts);
#endif
#if RT_NODE_UNROLL >= 4
                if (!tr.done(ts) && tr.at_node()) tr.exec_node(ts);
#endif
                if (tr.done(ts)) {
                    io.store(cur, tr.best, rng_pos);
                    active = false;
                }
            }
        }
        /* expensive leaves are held back until enough lanes want the same code: a ConstantMedium costs ~300
         * instructions (two boundary hits, Philox, log) and every ray of a foggy scene meets it, so the
         * lanes standing on one pile up quickly; they run together instead of 5 at a time. */
        bool hold = false;
#if RT_MEDIUM_QUORUM > 1
        {
            bool at_medium = active && (tr.h.x & 0xffu) == RTGPU_OP_MEDIUM;
            uint32_t mm = __ballot_sync(0xffffffffu, at_medium);
            if (mm) {
                /* hold only while something else certainly advances this iteration (a NODE round ran, or another
                 * lane runs a different leaf below); lanes on a NODE under the node quorum do NOT count */
                uint32_t other_leaf = __ballot_sync(0xffffffffu, active && !at_medium && tr.phase() != PH_NODE);
                hold = at_medium && (uint32_t)__popc(mm) < (uint32_t)RT_MEDIUM_QUORUM && (node_ran || other_leaf != 0u);
            }
        }
#endif
        if (active && !hold) {
            switch (tr.phase()) {
            case PH_NODE: break; /* below the quorum: waits for the next NODE rounds */
            case PH_EXACT: tr.exec_node_exact(ts); break;
            case PH_SPHERE: tr.exec_sphere(ts); break;
            case PH_BOX: tr.exec_box(ts); break;
            case PH_TRI:
                if constexpr (MESH) tr.exec_tri(ts);
                break;
            default: {
                LaneDraw<IO> draw{io, rng_a, rng_b, rng_pos};
                tr.template exec_misc<0, NEST, LaneDraw<IO>, MESH>(ts, stk, draw, &mf);
                rng_pos = draw.pos;
                break;
            }
            }
            if (tr.done(ts)) {
                io.store(cur, tr.best, rng_pos);
                active = false;
            }
        }
    }
}

/* ---- HitRecord of the winner (hittable/mod.rs:18-57), rebuilt once per ray ------------------- */
struct HitRec {
    V3 p, n;
    double t, u, v;
    bool front;
    uint32_t mat;
    int32_t gid, face, inst;
};
__device__ __forceinline__ void set_face_normal(HitRec &rec, V3 rdir, V3 outward) { /* hittable/mod.rs:49-56 */
    rec.front = dot(rdir, outward) < 0.;
    rec.n = rec.front ? outward : -outward;
}
__device__ __forceinline__ void sphere_uv(V3 p, double &u, double &v) { /* sphere.rs:30-34 */
    double theta = acos(-p.y);
    double phi = atan2(-p.z, p.x) + RT_PI;
    u = phi / (2. * RT_PI);
    v = theta / RT_PI;
}
template <int AX>
__device__ __forceinline__ void rect_record(HitRec &rec, V3 o, V3 d, double a0, double a1, double b0, double b1, double t) {
    double oa, ob, da, db;
    if (AX == 2) { oa = o.x; ob = o.y; da = d.x; db = d.y; }
    else if (AX == 1) { oa = o.x; ob = o.z; da = d.x; db = d.z; }
    else { oa = o.y; ob = o.z; da = d.y; db = d.z; }
    double a = oa + t * da;
    double b = ob + t * db;
    rec.u = (a - a0) / (a1 - a0);
    rec.v = (b - b0) / (b1 - b0);
    rec.p = o + d * t;
    set_face_normal(rec, d, v3(AX == 0 ? 1. : 0., AX == 1 ? 1. : 0., AX == 2 ? 1. : 0.));
}

/* WANT_UV = false skips the sphere's acos/atan2 (only an ImageTexture reads u, v of a sphere) */
#ifndef RT_RECORD_CALL
#define RT_RECORD_CALL __forceinline__ /* out of line, `rec`, the ray and the kernel's whole parameter block go through local memory: the
                                          shade stage of C3 was 34.5 ms with the call, 30.9 inlined */
#endif
template <bool WANT_UV>
__device__ RT_RECORD_CALL void make_record(const DScene &sc, const Ray &wr, double t, uint32_t opi, uint32_t face, HitRec &rec) {
    const uint4 *op = unit_ptr(sc.prog, opi);
    uint4 h = load_op_header(op);
    uint32_t kind = h.x & 0xffu;
    double p[6];
    load_op_payload(op, p);
    rtgpu_ctx ctx = sc.ctxs[h.x >> 8];
    /* outermost wrapper first: the ray each wrapper hands to its child */
    V3 o = wr.o, d = wr.d;
    V3 dirs[RT_MAX_CHAIN];
    for (uint32_t j = 0; j < ctx.len; ++j) {
        rtgpu_xform xf = load_xform(trav_of(sc), ctx.start + j);
        xform_ray(xf, o, d);
        dirs[j] = d;
    }
    rec.t = t;
    rec.mat = h.z;
    rec.gid = (int32_t)h.w;
    rec.face = 0;
    rec.inst = -1;
    switch (kind) {
    case RTGPU_OP_SPHERE:
    case RTGPU_OP_MSPHERE: { /* sphere.rs:58-64,157-163 */
        V3 c = v3(p[0], p[1], p[2]);
        if (kind == RTGPU_OP_MSPHERE) {
            double q[8];
            load_slot8(op + 4, q); /* second slot = units opi+2, opi+3 */
            c = c + (v3(q[0], q[1], q[2]) - c) * ((wr.tm - p[4]) / (p[5] - p[4]));
        }
        V3 at = o + d * t;
        V3 outward = (at - c) / p[3];
        if constexpr (WANT_UV) {
            sphere_uv(outward, rec.u, rec.v);
        } else {
            rec.u = 0.;
            rec.v = 0.;
        }
        rec.p = at;
        set_face_normal(rec, d, outward);
        break;
    }
    case RTGPU_OP_RECT_XY: rect_record<2>(rec, o, d, p[0], p[1], p[2], p[3], t); break;
    case RTGPU_OP_RECT_XZ: rect_record<1>(rec, o, d, p[0], p[1], p[2], p[3], t); break;
    case RTGPU_OP_RECT_YZ: rect_record<0>(rec, o, d, p[0], p[1], p[2], p[3], t); break;
    case RTGPU_OP_RING: /* ring.rs:48-51 */
        rec.p = o + d * t;
        rec.u = 0.;
        rec.v = 0.;
        set_face_normal(rec, d, v3(0., 1., 0.));
        break;
    case RTGPU_OP_BOX:
        rec.face = (int32_t)face;
        switch (face) { /* boxes.rs:24-69 */
        case 0: case 1: rect_record<2>(rec, o, d, p[0], p[3], p[1], p[4], t); break;
        case 2: case 3: rect_record<1>(rec, o, d, p[0], p[3], p[2], p[5], t); break;
        default: rect_record<0>(rec, o, d, p[1], p[4], p[2], p[5], t); break;
        }
        break;
    case RTGPU_OP_TRI: { /* triangle.rs:62-76 */
        double q[8];
        load_slot8(op + 4, q); /* second slot = units opi+2, opi+3 */
        V3 a = v3(p[0], p[1], p[2]), b = v3(p[3], p[4], p[5]), c = v3(q[0], q[1], q[2]), n = v3(q[3], q[4], q[5]);
        V3 pt = o + d * t;
        double a1 = a.x - b.x, b1 = a.x - c.x, c1 = a.x - pt.x;
        double a2 = a.y - b.y, b2 = a.y - c.y, c2 = a.y - pt.y;
        rec.u = (c1 * b2 - b1 * c2) / (a1 * b2 - b1 * a2);
        rec.v = (a1 * c2 - a2 * c1) / (a1 * b2 - b1 * a2);
        rec.p = pt;
        set_face_normal(rec, d, n);
        break;
    }
    default: /* MEDIUM, constantmedium.rs:70-78 */
        rec.p = o + d * t;
        rec.n = v3(1., 0., 0.);
        rec.u = 0.;
        rec.v = 0.;
        rec.front = true;
        break;
    }
    /* innermost wrapper first: record back to world space (hittable/mod.rs:167-171,250-261,283-285,323-327) */
    for (int j = (int)ctx.len - 1; j >= 0; --j) {
        rtgpu_xform xf = load_xform(trav_of(sc), ctx.start + (uint32_t)j);
        if (rec.inst < 0) rec.inst = (int32_t)xf.gid;
        switch (xf.kind) {
        case RTGPU_XF_TRANSLATE:
            rec.p = rec.p + v3(xf.p[0], xf.p[1], xf.p[2]);
            set_face_normal(rec, dirs[j], rec.n);
            break;
        case RTGPU_XF_ROTATE_Y: {
            double s = xf.p[0], c = xf.p[1];
            V3 pp = rec.p, nn = rec.n;
            pp.x = c * rec.p.x + s * rec.p.z;
            pp.z = -s * rec.p.x + c * rec.p.z;
            nn.x = c * rec.n.x + s * rec.n.z;
            nn.z = -s * rec.n.x + c * rec.n.z;
            rec.p = pp;
            set_face_normal(rec, dirs[j], nn); /* object-space ray vs world-space normal (Q4) */
            break;
        }
        case RTGPU_XF_ZOOM:
            rec.p = rec.p * xf.p[0];
            set_face_normal(rec, dirs[j], rec.n);
            break;
        default:
            rec.front = !rec.front;
            break;
        }
    }
}

/* f64 sin/cos expand to ~200 instructions each (plus the Payne-Hanek slow path); one out-of-line copy keeps
 * k_shade small enough for the instruction cache (ncu: `no_instruction` was its top stall). */
__device__ __noinline__ double rt_sin(double x) { return sin(x); }
/* (sin, cos) of phi = 2 pi r1, r1 in [0, 1): rt_trig.h (0.84 ulp; the general sincos() was a fifth of the Lambertian kernel).
 * An argument outside [0, 2 pi] cannot come from the two callers; it would give NaN, visibly, not a wrong direction. */
__device__ __forceinline__ double2 rt_sincos(double phi) {
    double2 r = make_double2(CUDART_NAN, CUDART_NAN);
    rt_sincos_2pi(phi, r.x, r.y);
    return r;
}

/* ---- textures (texture/mod.rs, texture/perlin.rs) ------------------------------------------ */
__device__ __forceinline__ int32_t floor_i32(double f) { /* `as i32`: saturating, NaN -> 0 */
    double fl = floor(f);
    if (fl != fl) return 0;
    if (fl >= 2147483647.0) return 2147483647;
    if (fl <= -2147483648.0) return (int32_t)0x80000000;
    return (int32_t)fl;
}
__device__ __noinline__ double perlin_noise(const rtgpu_perlin &t, V3 p) { /* perlin.rs:52-99 (double smoothing, Q9) */
    double u = p.x - floor(p.x), v = p.y - floor(p.y), w = p.z - floor(p.z);
    u = u * u * (3. - 2. * u);
    v = v * v * (3. - 2. * v);
    w = w * w * (3. - 2. * w);
    int32_t i = floor_i32(p.x), j = floor_i32(p.y), k = floor_i32(p.z);
    double uu = u * u * (3. - 2. * u), vv = v * v * (3. - 2. * v), ww = w * w * (3. - 2. * w);
    double accum = 0.;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                int idx = __ldg(&t.perm_x[(uint32_t)(i + di) & 255]) ^ __ldg(&t.perm_y[(uint32_t)(j + dj) & 255]) ^
                          __ldg(&t.perm_z[(uint32_t)(k + dk) & 255]);
                V3 c = v3(__ldg(&t.randvec[idx][0]), __ldg(&t.randvec[idx][1]), __ldg(&t.randvec[idx][2]));
                V3 weight_v = v3(u - (double)di, v - (double)dj, w - (double)dk);
                accum += dot(c, weight_v) * ((double)di * uu + (double)(1 - di) * (1. - uu)) *
                         ((double)dj * vv + (double)(1 - dj) * (1. - vv)) * ((double)dk * ww + (double)(1 - dk) * (1. - ww));
            }
    return accum;
}
__device__ double perlin_turb(const rtgpu_perlin &t, V3 p, int depth) { /* perlin.rs:100-112 */
    double accum = 0., weight = 1.;
    V3 tmp = p;
    for (int i = 0; i < depth; ++i) {
        accum += weight * perlin_noise(t, tmp);
        weight *= 0.5;
        tmp = tmp * 2.;
    }
    return fabs(accum);
}
/* TEXMASK: which texture kinds this instantiation can meet (bit k = rt_texture_kind k); the others are
 * compiled out so that the per-material-class shade kernels stay small */
#define RT_TEXMASK_SIMPLE 0x3 /* SolidColor, CheckerTexture over SolidColors */
#define RT_TEXMASK_ALL 0xf
/* in line: with these two out of line the class kernels spill around the calls (A/B on the B200: shade stage of C3 27.9 -> 27.3 ms,
 * of the Cornell box 5.63 -> 5.28 ms) */
#ifndef RT_TEX_CALL
#define RT_TEX_CALL __forceinline__
#endif
#ifndef RT_LPDF_CALL
#define RT_LPDF_CALL __forceinline__
#endif
template <int TEXMASK>
__device__ RT_TEX_CALL V3 texture_value(const TexTables sc, uint32_t id, double u, double v, V3 p) {
    for (int depth = 0; depth < 16; ++depth) {
        const rtgpu_texture &t = sc.texs[id];
        uint32_t kind = t.kind;
        if (kind == 0) /* SolidColor, texture/mod.rs:25-29 */
            return v3(t.p[0], t.p[1], t.p[2]);
        if (kind == 1) { /* CheckerTexture, texture/mod.rs:52-59 */
            double sines = rt_sin(p.x * 10.) * rt_sin(p.y * 10.) * rt_sin(p.z * 10.);
            id = (uint32_t)(sines < 0. ? t.a : t.b);
            continue;
        }
        if constexpr ((TEXMASK & 0x4) != 0) {
            if (kind == 2) { /* NoiseTexture, texture/mod.rs:75-79 */
                double s = 0.5 * (1. + rt_sin(t.p[0] * p.z + 10. * perlin_turb(sc.perlins[t.a], p, 7)));
                return v3(1. * s, 1. * s, 1. * s); /* Color(1,1,1) * 0.5 * (...) */
            }
        }
        if constexpr ((TEXMASK & 0x8) != 0) {
            if (kind == 3) { /* ImageTexture, texture/mod.rs:110-138 */
                if (t.a < 0) return v3(0., 1., 1.);
                const rtgpu_image &im = sc.images[t.a];
                if (u < 0.) u = 0.;
                if (u > 1.) u = 1.; /* f64::clamp keeps NaN */
                if (v < 0.) v = 0.;
                if (v > 1.) v = 1.;
                double fi = u * (double)im.width, fj = v * (double)im.height;
                long long i = (fi > 0.) ? (fi >= 9.2e18 ? 0x7fffffffffffffffLL : (long long)fi) : 0; /* `as usize` */
                long long j = (fj > 0.) ? (fj >= 9.2e18 ? 0x7fffffffffffffffLL : (long long)fj) : 0;
                if (i >= im.width) i = im.width - 1;
                if (j >= im.height) j = im.height - 1;
                const double color_scale = 1. / 255.999;
                const uint8_t *px = sc.texels + im.offset + 3 * (j * (long long)im.width + i);
                return v3((double)__ldg(px) * color_scale, (double)__ldg(px + 1) * color_scale, (double)__ldg(px + 2) * color_scale);
            }
        }
        break;
    }
    return v3(0., 0., 0.);
}

/* ---- random helpers (basic/vec.rs, basic/pdf.rs) --------------------------------------------- */
/* `p.length() < 1.` of the rejection loops (vec.rs:72, :92), without the square root: for a correctly rounded sqrt,
 * fl(sqrt(s)) < 1 <=> s < 1.  (s >= 1: sqrt(s) >= 1 and rounding is monotone.  s < 1: s <= 1 - 2^-53, and
 * (1 - 2^-54)^2 = 1 - 2^-53 + 2^-108 > s, so sqrt(s) lies below the midpoint of 1 - 2^-53 and 1 and rounds to at most
 * 1 - 2^-53.  NaN fails both comparisons.)  s itself is formed as length() forms it. */
__device__ __forceinline__ bool length_below_one(V3 p) { return length_sqr(p) < 1.; }
__device__ __forceinline__ V3 random_in_unit_sphere(rt_rng &rng) { /* vec.rs:69-76 */
    for (;;) {
        uint64_t ux, uy, uz;
        rng.next3_u64(ux, uy, uz); /* = three gen_range(-1, 1) draws, in this order */
        double x = -1. + (1. - -1.) * rt_rng::to_unit_f64(ux);
        double y = -1. + (1. - -1.) * rt_rng::to_unit_f64(uy);
        double z = -1. + (1. - -1.) * rt_rng::to_unit_f64(uz);
        V3 p = v3(x, y, z);
        if (length_below_one(p)) return p;
    }
}
__device__ __forceinline__ V3 random_in_unit_disk(rt_rng &rng) { /* vec.rs:88-96 */
    for (;;) {
        double x = rng.gen_range(-1.0, 1.0);
        double y = rng.gen_range(-1.0, 1.0);
        V3 p = v3(x, y, 0.);
        if (length_below_one(p)) return p;
    }
}
__device__ __forceinline__ V3 random_cosine_direction(rt_rng &rng) { /* pdf.rs:12-21 */
    double r1 = rng.gen();
    double r2 = rng.gen();
    double z = sqrt(1. - r2);
    double phi = 2. * RT_PI * r1;
    const double2 sc2 = rt_sincos(phi);
    const double sn = sc2.x, cs = sc2.y;
    double x = cs * sqrt(r2);
    double y = sn * sqrt(r2);
    return v3(x, y, z);
}
__device__ __forceinline__ V3 random_to_sphere(rt_rng &rng, double radius, double dis_sqr) { /* vec.rs:108-117 */
    double r1 = rng.gen();
    double r2 = rng.gen();
    double z = 1. + r2 * (sqrt(1. - radius * radius / dis_sqr) - 1.);
    double phi = 2. * RT_PI * r1;
    const double2 sc2 = rt_sincos(phi);
    const double sn = sc2.x, cs = sc2.y;
    double x = cs * sqrt(1. - z * z);
    double y = sn * sqrt(1. - z * z);
    return v3(x, y, z);
}
struct Onb {
    V3 u, v, w;
};
__device__ __forceinline__ Onb onb_from_w(V3 n) { /* onb.rs:26-35 */
    Onb o;
    o.w = to_unit(n);
    V3 a = fabs(o.w.x) > 0.9 ? v3(0., 1., 0.) : v3(1., 0., 0.);
    o.v = to_unit(cross(o.w, a));
    o.u = cross(o.w, o.v);
    return o;
}
__device__ __forceinline__ V3 onb_local(const Onb &o, V3 a) { return o.u * a.x + o.v * a.y + o.w * a.z; } /* onb.rs:22-24 */
__device__ __forceinline__ V3 reflect(V3 v, V3 n) { return v - n * dot(v, n) * 2.; } /* vec.rs:119-121 */
__device__ __forceinline__ V3 refract(V3 uv, V3 n, double etai_over_etat) { /* vec.rs:123-128 */
    double cos_theta = fmin(dot(-uv, n), 1.);
    V3 r_out_perp = (uv + n * cos_theta) * etai_over_etat;
    V3 r_out_parallel = -n * sqrt(fabs(1. - length_sqr(r_out_perp)));
    return r_out_perp + r_out_parallel;
}
__device__ __forceinline__ double reflectance(double cosv, double ref_idx) { /* material/mod.rs:112-116 */
    double r0 = (1. - ref_idx) / (1. + ref_idx);
    r0 = r0 * r0;
    double x = 1. - cosv;
    double x2 = x * x;
    double x4 = x2 * x2;
    return r0 + (1. - r0) * (x * x4); /* powi(5) as compiler-rt's __powidf2 evaluates it */
}

/* ---- lights: HittableList::{random, pdf_value} over the lowered `lights` (hittable/mod.rs:121-132) */
template <int AX>
__device__ __forceinline__ double rect_pdf_value(const rtgpu_light &l, V3 o, V3 v) { /* aarect.rs:74-83,157-166,240-249 */
    double best = CUDART_INF;
    if (!hit_rect<AX>(o, v, l.p[0], l.p[1], l.p[2], l.p[3], l.p[4], 0.001, best)) return 0.;
    double area = (l.p[1] - l.p[0]) * (l.p[3] - l.p[2]);
    double dis_sqr = best * best * length_sqr(v);
    V3 outward = v3(AX == 0 ? 1. : 0., AX == 1 ? 1. : 0., AX == 2 ? 1. : 0.);
    V3 n = dot(v, outward) < 0. ? outward : -outward;
    double cosv = fabs(dot(v, n) / length(v));
    return dis_sqr / (cosv * area);
}
/* pdf_value of one member that is not a list */
__device__ __forceinline__ double light_pdf_value(const rtgpu_light &l, V3 o, V3 v) {
    switch (l.kind) {
    case RTGPU_LIGHT_RECT_XY: return rect_pdf_value<2>(l, o, v);
    case RTGPU_LIGHT_RECT_XZ: return rect_pdf_value<1>(l, o, v);
    case RTGPU_LIGHT_RECT_YZ: return rect_pdf_value<0>(l, o, v);
    case RTGPU_LIGHT_SPHERE: { /* sphere.rs:75-83 */
        V3 c = v3(l.p[0], l.p[1], l.p[2]);
        double best = CUDART_INF;
        if (hit_sphere(o, v, length_sqr(v), c, l.p[3], 0.001, best)) {
            double cos_max = sqrt(1. - l.p[3] * l.p[3] / length_sqr(c - o));
            double solid_angle = 2. * RT_PI * (1. - cos_max);
            return 1. / solid_angle;
        }
        return 0.;
    }
    default: return 0.; /* trait default, hittable/mod.rs:62-64 */
    }
}
/* HittableList::pdf_value (hittable/mod.rs:121-128) over `lights`: the mean of the members' values, a member that is a
 * HittableList contributing the mean over ITS members.  Sums run in member order, as upstream; the recursion is unrolled
 * over an explicit stack (RTGPU_MAX_LIGHT_NESTING levels, checked when the scene is created). */
__device__ RT_LPDF_CALL double lights_pdf_value(const LightTable sc, V3 o, V3 v) {
    if (sc.n_lights == sc.n_light_roots) { /* no HittableList inside `lights` (every reference scene): the same sums without the stack */
        double sum = 0.;
        for (uint32_t k = 0; k < sc.n_light_roots; ++k) sum += light_pdf_value(sc.lights[k], o, v);
        return rt_div_zero_fast(sum, (double)sc.n_light_roots); /* 0 whenever the direction misses every light */
    }
    uint32_t at[RTGPU_MAX_LIGHT_NESTING + 1], end[RTGPU_MAX_LIGHT_NESTING + 1];
    double sum[RTGPU_MAX_LIGHT_NESTING + 1], len[RTGPU_MAX_LIGHT_NESTING + 1];
    int sp = 0;
    at[0] = 0;
    end[0] = sc.n_light_roots;
    sum[0] = 0.;
    len[0] = (double)sc.n_light_roots;
    for (;;) {
        if (at[sp] < end[sp]) {
            const rtgpu_light &l = sc.lights[at[sp]++];
            if (l.kind == RTGPU_LIGHT_LIST && sp < RTGPU_MAX_LIGHT_NESTING) {
                ++sp;
                at[sp] = (uint32_t)l.p[0];
                end[sp] = at[sp] + (uint32_t)l.p[1];
                sum[sp] = 0.;
                len[sp] = l.p[1];
            } else {
                sum[sp] += light_pdf_value(l, o, v);
            }
        } else {
            double val = rt_div_zero_fast(sum[sp], len[sp]);
            if (sp == 0) return val;
            --sp;
            sum[sp] += val;
        }
    }
}
/* HittableList::random (hittable/mod.rs:129-132): uniform pick of a member (one draw), then that member's random();
 * a member that is a HittableList picks again */
#ifndef RT_LIGHTS_RANDOM_CALL
#define RT_LIGHTS_RANDOM_CALL __forceinline__ /* out of line, `rng` is passed by address and lives in local memory for the whole kernel */
#endif
__device__ RT_LIGHTS_RANDOM_CALL V3 lights_random(const LightTable sc, V3 o, rt_rng &rng) {
    uint32_t target = rng.gen_index(sc.n_light_roots);
    const rtgpu_light *lp = sc.lights + target;
    for (int depth = 0; lp->kind == RTGPU_LIGHT_LIST && depth < RTGPU_MAX_LIGHT_NESTING; ++depth)
        lp = sc.lights + (uint32_t)lp->p[0] + rng.gen_index((uint32_t)lp->p[1]);
    const rtgpu_light &l = *lp;
    switch (l.kind) {
    case RTGPU_LIGHT_RECT_XY: { /* aarect.rs:85-93 */
        double a = rng.gen_range(l.p[0], l.p[1]);
        double b = rng.gen_range(l.p[2], l.p[3]);
        return v3(a, b, l.p[4]) - o;
    }
    case RTGPU_LIGHT_RECT_XZ: { /* aarect.rs:168-176 */
        double a = rng.gen_range(l.p[0], l.p[1]);
        double b = rng.gen_range(l.p[2], l.p[3]);
        return v3(a, l.p[4], b) - o;
    }
    case RTGPU_LIGHT_RECT_YZ: { /* aarect.rs:251-259 */
        double a = rng.gen_range(l.p[0], l.p[1]);
        double b = rng.gen_range(l.p[2], l.p[3]);
        return v3(l.p[4], a, b) - o;
    }
    case RTGPU_LIGHT_SPHERE: { /* sphere.rs:85-90 */
        V3 direction = v3(l.p[0], l.p[1], l.p[2]) - o;
        double dis_sqr = length_sqr(direction);
        Onb uvw = onb_from_w(direction);
        return onb_local(uvw, random_to_sphere(rng, l.p[3], dis_sqr));
    }
    default:
        return v3(1., 0., 0.); /* trait default, hittable/mod.rs:65-67 */
    }
}
