/*
 * oracle.cpp -- CPU restatement of the reference's `ray_color` hot path.
 *
 * TEST INFRASTRUCTURE.  Nothing under raytracer-2022_b200/ may include, link or call this
 * file; it exists so tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs can
 * check (and time) the CUDA path against the reference's algorithm.
 *
 * What it follows: /root/reference/raytracer/src (Jerx2y/Raytracer-2022), function by
 * function, keeping its structure -- pointer-linked BvhNode tree, virtual dispatch,
 * recursive ray_color, f64 everywhere, per-node 1/dir divisions, per-call triangle
 * normal -- and its quirks.  Each function cites the file:line it
 * restates.  Build flags must forbid FMA contraction and fast-math (see Makefile).
 *
 * PARITY UNPINNED: the reference has no tests, fixtures or golden vectors, and no Rust
 * toolchain exists in this image, so the reference cannot be run here.  The pins are the
 * hand-derived known-answer vectors K1-K12 of SURVEY.md 4 (tests/test_oracle_kat.py).
 *
 * Third-party arithmetic that is not under /root/reference:
 *   rand 0.8.5 / rand_chacha 0.3.1 (Cargo.lock:372,383): `thread_rng()` is an unseedable
 *   ChaCha12 stream, so only distributional parity is possible.  It is replaced by the
 *   Philox4x32-10 slot streams specified in include/rt_rng.h (re-implemented here on
 *   purpose, to keep the oracle independent of product headers): gen::<f64>() is the
 *   53-bit recipe, gen_range for floats is lo+(hi-lo)*u, integer gen_range is rand's
 *   widening-multiply zone rejection.
 */
#include "oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

const double PI = 3.14159265358979323846264338327950288; // std::f64::consts::PI
const double INF = std::numeric_limits<double>::infinity();
const double F64_MAX = std::numeric_limits<double>::max();

thread_local std::string g_err;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 slot streams (spec: include/rt_rng.h; independent implementation)
// ---------------------------------------------------------------------------------------------
void philox(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
    uint32_t c[4] = {c_in[0], c_in[1], c_in[2], c_in[3]};
    uint32_t k[2] = {k_in[0], k_in[1]};
    for (int round = 0; round < 10; ++round) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n[4] = {(uint32_t)(p1 >> 32) ^ c[1] ^ k[0], (uint32_t)p1,
                         (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], (uint32_t)p0};
        std::memcpy(c, n, sizeof n);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    std::memcpy(out, c, 4 * sizeof(uint32_t));
}

const uint32_t DOMAIN_PATH = 0x50415448u, DOMAIN_BATCH = 0x42544348u, DOMAIN_SCENE = 0x53434e45u;

struct Rng {
    uint32_t key[2], stream[3], pos;
    Rng(uint64_t seed, uint32_t s0, uint32_t s1, uint32_t domain, uint32_t start = 0) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        stream[0] = s0;
        stream[1] = s1;
        stream[2] = domain;
        pos = start;
    }
    uint64_t slot() {
        uint32_t ctr[4] = {pos >> 1, stream[0], stream[1], stream[2]}, w[4];
        philox(ctr, key, w);
        uint64_t v = (pos & 1u) ? (((uint64_t)w[2] << 32) | w[3]) : (((uint64_t)w[0] << 32) | w[1]);
        ++pos;
        return v;
    }
    double gen() { return (double)(slot() >> 11) * (1.0 / 9007199254740992.0); }
    double gen_range(double lo, double hi) { return lo + (hi - lo) * gen(); }
    uint32_t gen_index(uint32_t n) {
        uint64_t range = n;
        uint64_t zone = (range << __builtin_clzll(range)) - 1;
        for (;;) {
            unsigned __int128 m = (unsigned __int128)slot() * range;
            if ((uint64_t)m <= zone) return (uint32_t)(m >> 64);
        }
    }
};

// `rand::thread_rng()`: every call site in the reference reaches the generator through a
// thread-local; the oracle points it at the stream of the path (or ray) being evaluated.
thread_local Rng *tl_rng = nullptr;
thread_local uint64_t tl_rays = 0;
Rng &thread_rng() { return *tl_rng; }

// ---------------------------------------------------------------------------------------------
// basic/vec.rs
// ---------------------------------------------------------------------------------------------
struct Vec3 {
    double x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double &at(int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double length_sqr() const { return x * x + y * y + z * z; } // vec.rs:37-39
    double length() const { return std::sqrt(length_sqr()); }   // vec.rs:41-43
};
typedef Vec3 Point3;
typedef Vec3 Color;
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 operator*(Vec3 a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator/(Vec3 a, double s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } // vec.rs:24-26
inline Vec3 cross(Vec3 a, Vec3 b) {                                              // vec.rs:28-34
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline Vec3 to_unit(Vec3 a) { return a / a.length(); } // vec.rs:45-47

Vec3 vec3_random(double l, double r) { // vec.rs:49-56
    Rng &rng = thread_rng();
    double x = rng.gen_range(l, r);
    double y = rng.gen_range(l, r);
    double z = rng.gen_range(l, r);
    return Vec3(x, y, z);
}
Vec3 random_in_unit_sphere() { // vec.rs:69-76
    for (;;) {
        Vec3 p = vec3_random(-1., 1.);
        if (p.length() < 1.) return p;
    }
}
Vec3 random_in_unit_disk() { // vec.rs:88-96
    Rng &rng = thread_rng();
    for (;;) {
        double x = rng.gen_range(-1.0, 1.0);
        double y = rng.gen_range(-1.0, 1.0);
        Vec3 p(x, y, 0.);
        if (p.length() < 1.) return p;
    }
}
Vec3 random_to_sphere(double radius, double dis_sqr) { // vec.rs:108-117
    Rng &rng = thread_rng();
    double r1 = rng.gen();
    double r2 = rng.gen();
    double z = 1. + r2 * (std::sqrt(1. - radius * radius / dis_sqr) - 1.);
    double phi = 2. * PI * r1;
    double x = std::cos(phi) * std::sqrt(1. - z * z);
    double y = std::sin(phi) * std::sqrt(1. - z * z);
    return Vec3(x, y, z);
}
Vec3 reflect(Vec3 v, Vec3 n) { return v - n * dot(v, n) * 2.; } // vec.rs:119-121
Vec3 refract(Vec3 uv, Vec3 n, double etai_over_etat) {         // vec.rs:123-128
    double cos_theta = std::fmin(dot(-uv, n), 1.);
    Vec3 r_out_perp = (uv + n * cos_theta) * etai_over_etat;
    Vec3 r_out_parallel = -n * std::sqrt(std::fabs(1. - r_out_perp.length_sqr()));
    return r_out_perp + r_out_parallel;
}

// basic/ray.rs
struct Ray {
    Point3 orig;
    Vec3 dir;
    double tm;
    Ray() : tm(0) {}
    Ray(Point3 o, Vec3 d, double t) : orig(o), dir(d), tm(t) {}
    Point3 at(double t) const { return orig + dir * t; } // ray.rs:18-20
};

// basic/onb.rs
struct Onb {
    Vec3 axis[3];
    Vec3 u() const { return axis[0]; }
    Vec3 v() const { return axis[1]; }
    Vec3 w() const { return axis[2]; }
    Vec3 local_vec(Vec3 a) const { return u() * a.x + v() * a.y + w() * a.z; } // onb.rs:22-24
    static Onb build_from_w(Vec3 n) {                                         // onb.rs:26-35
        Onb o;
        Vec3 w = to_unit(n);
        Vec3 a = std::fabs(w.x) > 0.9 ? Vec3(0., 1., 0.) : Vec3(1., 0., 0.);
        Vec3 v = to_unit(cross(w, a));
        Vec3 u = cross(w, v);
        o.axis[0] = u;
        o.axis[1] = v;
        o.axis[2] = w;
        return o;
    }
};

// ---------------------------------------------------------------------------------------------
// hittable/bvh/aabb.rs
// ---------------------------------------------------------------------------------------------
struct AABB {
    Point3 min, max;
    AABB() {}
    AABB(Point3 a, Point3 b) : min(a), max(b) {}
    bool hit(const Ray &r, double tmin, double tmax) const { // aabb.rs:15-32
        double t_min = tmin, t_max = tmax;
        for (int i = 0; i < 3; ++i) {
            double inv_d = 1. / r.dir[i];
            double t0 = (min[i] - r.orig[i]) * inv_d;
            double t1 = (max[i] - r.orig[i]) * inv_d;
            if (inv_d < 0.) std::swap(t0, t1);
            t_min = t0 > t_min ? t0 : t_min;
            t_max = t1 < t_max ? t1 : t_max;
            if (t_max <= t_min) return false;
        }
        return true;
    }
    static AABB surrounding_box(AABB b0, AABB b1) { // aabb.rs:34-46 (f64::min/max ignore NaN)
        Point3 small(std::fmin(b0.min.x, b1.min.x), std::fmin(b0.min.y, b1.min.y),
                     std::fmin(b0.min.z, b1.min.z));
        Point3 large(std::fmax(b0.max.x, b1.max.x), std::fmax(b0.max.y, b1.max.y),
                     std::fmax(b0.max.z, b1.max.z));
        return AABB(small, large);
    }
};

// ---------------------------------------------------------------------------------------------
// texture/
// ---------------------------------------------------------------------------------------------
struct Texture {
    virtual ~Texture() {}
    virtual Color value(double u, double v, Point3 p) const = 0;
};
struct SolidColor : Texture {
    Color c;
    explicit SolidColor(Color col) : c(col) {}
    Color value(double, double, Point3) const override { return c; } // texture/mod.rs:25-29
};
struct CheckerTexture : Texture {
    std::shared_ptr<Texture> odd, even;
    Color value(double u, double v, Point3 p) const override { // texture/mod.rs:52-59
        double sines = std::sin(p.x * 10.) * std::sin(p.y * 10.) * std::sin(p.z * 10.);
        return sines < 0. ? odd->value(u, v, p) : even->value(u, v, p);
    }
};
struct Perlin {
    const rt_perlin *t;
    static int32_t floor_i32(double f) { // `as i32`: saturating, NaN -> 0
        double fl = std::floor(f);
        if (fl != fl) return 0;
        if (fl >= 2147483647.0) return 2147483647;
        if (fl <= -2147483648.0) return (int32_t)0x80000000;
        return (int32_t)fl;
    }
    static double trilinear_interp(Vec3 c[2][2][2], double u, double v, double w) { // perlin.rs:81-99
        double uu = u * u * (3. - 2. * u);
        double vv = v * v * (3. - 2. * v);
        double ww = w * w * (3. - 2. * w);
        double accum = 0.;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 2; ++k) {
                    Vec3 weight_v(u - (double)i, v - (double)j, w - (double)k);
                    accum += dot(c[i][j][k], weight_v) *
                             ((double)i * uu + (double)(1 - i) * (1. - uu)) *
                             ((double)j * vv + (double)(1 - j) * (1. - vv)) *
                             ((double)k * ww + (double)(1 - k) * (1. - ww));
                }
        return accum;
    }
    double noise(Point3 p) const { // perlin.rs:52-78 (Hermite smoothing applied twice, Q9)
        double u = p.x - std::floor(p.x);
        double v = p.y - std::floor(p.y);
        double w = p.z - std::floor(p.z);
        u = u * u * (3. - 2. * u);
        v = v * v * (3. - 2. * v);
        w = w * w * (3. - 2. * w);
        int32_t i = floor_i32(p.x), j = floor_i32(p.y), k = floor_i32(p.z);
        Vec3 c[2][2][2];
        for (int di = 0; di < 2; ++di)
            for (int dj = 0; dj < 2; ++dj)
                for (int dk = 0; dk < 2; ++dk) {
                    int idx = t->perm_x[(uint32_t)(i + di) & 255] ^ t->perm_y[(uint32_t)(j + dj) & 255] ^
                              t->perm_z[(uint32_t)(k + dk) & 255];
                    c[di][dj][dk] = Vec3(t->randvec[idx][0], t->randvec[idx][1], t->randvec[idx][2]);
                }
        return trilinear_interp(c, u, v, w);
    }
    double turb(Point3 p, int depth) const { // perlin.rs:100-112
        double accum = 0.;
        Point3 tmp_p = p;
        double weight = 1.;
        for (int i = 0; i < depth; ++i) {
            accum += weight * noise(tmp_p);
            weight *= 0.5;
            tmp_p = tmp_p * 2.;
        }
        return std::fabs(accum);
    }
};
struct NoiseTexture : Texture {
    Perlin noise;
    double scale;
    Color value(double, double, Point3 p) const override { // texture/mod.rs:75-79
        return Color(1., 1., 1.) * 0.5 * (1. + std::sin(scale * p.z + 10. * noise.turb(p, 7)));
    }
};
struct ImageTexture : Texture {
    uint32_t width = 0, height = 0;
    const uint8_t *pixel_color = nullptr; // rows bottom-up (texture/mod.rs:94-100)
    static double clamp01(double x) {     // f64::clamp keeps NaN
        if (x < 0.) return 0.;
        if (x > 1.) return 1.;
        return x;
    }
    static size_t as_usize(double x) { // `as usize`: NaN -> 0, saturating
        if (!(x > 0.)) return 0;
        if (x >= 18446744073709551615.0) return (size_t)-1;
        return (size_t)x;
    }
    Color value(double u, double v, Point3) const override { // texture/mod.rs:110-138
        if (pixel_color == nullptr || width == 0 || height == 0) return Color(0., 1., 1.);
        u = clamp01(u);
        v = clamp01(v);
        size_t i = as_usize(u * (double)width);
        size_t j = as_usize(v * (double)height);
        if (i >= (size_t)width) i = (size_t)width - 1;
        if (j >= (size_t)height) j = (size_t)height - 1;
        double color_scale = 1. / 255.999;
        const uint8_t *pixel = pixel_color + 3 * (j * (size_t)width + i);
        return Color((double)pixel[0] * color_scale, (double)pixel[1] * color_scale,
                     (double)pixel[2] * color_scale);
    }
};

// ---------------------------------------------------------------------------------------------
// hittable/mod.rs: HitRecord, Hittable
// ---------------------------------------------------------------------------------------------
struct Material;
struct HitRecord {
    Point3 p;
    Vec3 normal;
    double t = 0, u = 0, v = 0;
    bool front_face = false;
    const Material *mat_ptr = nullptr;
    int32_t gid = -1, face = 0, inst = -1; // bookkeeping for the parity tests only
    void set_face_normal(const Ray &r, Vec3 outward_normal) { // hittable/mod.rs:49-56
        front_face = dot(r.dir, outward_normal) < 0.;
        normal = front_face ? outward_normal : -outward_normal;
    }
};

struct Hittable {
    int32_t gid = -1;
    virtual ~Hittable() {}
    virtual bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const = 0;
    virtual bool bounding_box(double time0, double time1, AABB &out) const = 0;
    virtual double pdf_value(Point3, Vec3) const { return 0.; }       // hittable/mod.rs:62-64
    virtual Vec3 random(Point3) const { return Vec3(1., 0., 0.); }    // hittable/mod.rs:65-67
};
typedef std::shared_ptr<Hittable> HittablePtr;

// basic/pdf.rs
Vec3 random_cosine_direction() { // pdf.rs:12-21
    Rng &rng = thread_rng();
    double r1 = rng.gen();
    double r2 = rng.gen();
    double z = std::sqrt(1. - r2);
    double phi = 2. * PI * r1;
    double x = std::cos(phi) * std::sqrt(r2);
    double y = std::sin(phi) * std::sqrt(r2);
    return Vec3(x, y, z);
}
struct CosPdf {
    Onb uvw;
    explicit CosPdf(Vec3 w) : uvw(Onb::build_from_w(w)) {}                           // pdf.rs:35-39
    Vec3 generate() const { return uvw.local_vec(random_cosine_direction()); }       // pdf.rs:43-45
    double value(Vec3 direction) const {                                             // pdf.rs:46-53
        double c = dot(to_unit(direction), uvw.w());
        return c <= 0. ? 0. : c / PI;
    }
};

// ---------------------------------------------------------------------------------------------
// material/mod.rs
// ---------------------------------------------------------------------------------------------
struct ScatterRecord {
    bool has_specular = false;
    Ray specular_ray;
    Color attenuation;
    bool has_pdf = false;
    Vec3 pdf_w; // CosPdf::new(rec.normal), built lazily by the caller
};
struct Material {
    int32_t mid = -1;
    virtual ~Material() {}
    virtual bool scatter(const Ray &, const HitRecord &, ScatterRecord &) const { return false; }
    virtual double scattering_pdf(const Ray &, const HitRecord &, const Ray &) const { return 0.; }
    virtual Color emitted(const Ray &, const HitRecord &, double, double, Point3) const {
        return Color(0., 0., 0.);
    }
};
struct Lambertian : Material {
    std::shared_ptr<Texture> albedo;
    bool scatter(const Ray &, const HitRecord &rec, ScatterRecord &s) const override { // :51-57
        s.has_specular = false;
        s.attenuation = albedo->value(rec.u, rec.v, rec.p);
        s.has_pdf = true;
        s.pdf_w = rec.normal;
        return true;
    }
    double scattering_pdf(const Ray &, const HitRecord &rec, const Ray &scattered) const override {
        double cosine = dot(rec.normal, to_unit(scattered.dir)); // :58-65
        return cosine < 0. ? 0. : cosine / PI;
    }
};
struct Metal : Material {
    Color albedo;
    double fuzz; // min(f,1) applied by the constructor (:76-81)
    bool scatter(const Ray &r_in, const HitRecord &rec, ScatterRecord &s) const override { // :85-96
        Vec3 reflected = reflect(to_unit(r_in.dir), rec.normal);
        s.has_specular = true;
        s.specular_ray = Ray(rec.p, reflected + random_in_unit_sphere() * fuzz, 0.);
        s.attenuation = albedo;
        s.has_pdf = false;
        return true;
    }
};
double powi5(double x) { // compiler-rt __powidf2(x, 5): r = x; x^2; x^4; r *= x^4
    double x2 = x * x;
    double x4 = x2 * x2;
    return x * x4;
}
double reflectance(double cos, double ref_idx) { // :112-116
    double r0 = (1. - ref_idx) / (1. + ref_idx);
    r0 = r0 * r0;
    return r0 + (1. - r0) * powi5(1. - cos);
}
struct Dielectric : Material {
    double ir;
    bool scatter(const Ray &r_in, const HitRecord &rec, ScatterRecord &s) const override { // :120-147
        double refraction_ratio = rec.front_face ? 1. / ir : ir;
        Vec3 unit_direction = to_unit(r_in.dir);
        double cos_theta = std::fmin(dot(-unit_direction, rec.normal), 1.);
        double sin_theta = std::sqrt(1. - cos_theta * cos_theta);
        bool cannot_refract = refraction_ratio * sin_theta > 1.;
        double random_double = thread_rng().gen_range(0.0, 1.0);
        Vec3 direction = (cannot_refract || reflectance(cos_theta, refraction_ratio) > random_double)
                             ? reflect(unit_direction, rec.normal)
                             : refract(unit_direction, rec.normal, refraction_ratio);
        s.has_specular = true;
        s.specular_ray = Ray(rec.p, direction, r_in.tm);
        s.attenuation = Color(1., 1., 1.);
        s.has_pdf = false;
        return true;
    }
};
struct DiffuseLight : Material {
    std::shared_ptr<Texture> emit;
    Color emitted(const Ray &, const HitRecord &rec, double u, double v, Point3 p) const override {
        return rec.front_face ? emit->value(u, v, p) : Color(0., 0., 0.); // :173-181
    }
};
struct Isotropic : Material {
    std::shared_ptr<Texture> albedo;
    bool scatter(const Ray &r_in, const HitRecord &rec, ScatterRecord &s) const override { // :206-214
        s.has_specular = true;
        s.specular_ray = Ray(rec.p, random_in_unit_sphere(), r_in.tm);
        s.attenuation = albedo->value(rec.u, rec.v, rec.p);
        s.has_pdf = false;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// hittable/mod.rs: HittableList and the wrappers
// ---------------------------------------------------------------------------------------------
struct HittableList : Hittable {
    std::vector<HittablePtr> objects;
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :90-100
        bool any = false;
        double closest_so_far = t_max;
        for (const HittablePtr &object : objects) {
            HitRecord tmp;
            if (object->hit(r, t_min, closest_so_far, tmp)) {
                closest_so_far = tmp.t;
                rec = tmp;
                any = true;
            }
        }
        return any;
    }
    bool bounding_box(double time0, double time1, AABB &out) const override { // :102-120
        if (objects.empty()) return false;
        AABB output_box(Point3(INF, INF, INF), Point3(-INF, -INF, -INF));
        for (const HittablePtr &object : objects) {
            AABB temp_box;
            if (!object->bounding_box(time0, time1, temp_box)) return false;
            output_box = AABB::surrounding_box(output_box, temp_box);
        }
        out = output_box;
        return true;
    }
    double pdf_value(Point3 o, Vec3 v) const override { // :121-128
        size_t len = objects.size();
        double sum = 0.;
        for (size_t i = 0; i < len; ++i) sum += objects[i]->pdf_value(o, v);
        return sum / (double)len;
    }
    Vec3 random(Point3 o) const override { // :129-132
        uint32_t target = thread_rng().gen_index((uint32_t)objects.size());
        return objects[target]->random(o);
    }
};

struct Translate : Hittable {
    HittablePtr ptr;
    Vec3 offset;
    bool bounding_box(double time0, double time1, AABB &out) const override { // :154-163
        AABB b;
        if (!ptr->bounding_box(time0, time1, b)) return false;
        out = AABB(b.min + offset, b.max + offset);
        return true;
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :165-174
        Ray moved_r(r.orig - offset, r.dir, r.tm);
        if (!ptr->hit(moved_r, t_min, t_max, rec)) return false;
        rec.p = rec.p + offset;
        rec.set_face_normal(moved_r, rec.normal);
        if (rec.inst < 0) rec.inst = gid;
        return true;
    }
};
struct RotateY : Hittable {
    HittablePtr ptr;
    double sin_theta, cos_theta;
    bool has_box;
    AABB aabbox;
    RotateY(HittablePtr p, double angle) : ptr(p) { // :188-232
        double radians = angle * (PI / 180.0);      // f64::to_radians
        sin_theta = std::sin(radians);
        cos_theta = std::cos(radians);
        AABB output_box;
        has_box = p->bounding_box(0., 1., output_box);
        if (has_box) {
            Point3 mn(INF, INF, INF), mx(-INF, -INF, -INF);
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    for (int k = 0; k < 2; ++k) {
                        double x = (double)i * output_box.max.x + (double)(1 - i) * output_box.min.x;
                        double y = (double)j * output_box.max.y + (double)(1 - j) * output_box.min.y;
                        double z = (double)k * output_box.max.z + (double)(1 - k) * output_box.min.z;
                        double newx = cos_theta * x + sin_theta * z;
                        double newz = -sin_theta * x + cos_theta * z;
                        Vec3 tester(newx, y, newz);
                        for (int c = 0; c < 3; ++c) {
                            mn.at(c) = std::fmin(mn[c], tester[c]);
                            mx.at(c) = std::fmax(mx[c], tester[c]);
                        }
                    }
            aabbox = AABB(mn, mx);
        }
    }
    bool bounding_box(double, double, AABB &out) const override { // :236-238
        if (!has_box) return false;
        out = aabbox;
        return true;
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :239-264
        Point3 origin = r.orig;
        Vec3 direction = r.dir;
        origin.x = cos_theta * r.orig.x - sin_theta * r.orig.z;
        origin.z = sin_theta * r.orig.x + cos_theta * r.orig.z;
        direction.x = cos_theta * r.dir.x - sin_theta * r.dir.z;
        direction.z = sin_theta * r.dir.x + cos_theta * r.dir.z;
        Ray rotated_r(origin, direction, r.tm);
        if (!ptr->hit(rotated_r, t_min, t_max, rec)) return false;
        Point3 p = rec.p;
        Vec3 normal = rec.normal;
        p.x = cos_theta * rec.p.x + sin_theta * rec.p.z;
        p.z = -sin_theta * rec.p.x + cos_theta * rec.p.z;
        normal.x = cos_theta * rec.normal.x + sin_theta * rec.normal.z;
        normal.z = -sin_theta * rec.normal.x + cos_theta * rec.normal.z;
        rec.p = p;
        rec.set_face_normal(rotated_r, normal); // object-space ray vs world-space normal (Q4)
        if (rec.inst < 0) rec.inst = gid;
        return true;
    }
};
struct FlipFace : Hittable {
    HittablePtr ptr;
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :281-288
        if (!ptr->hit(r, t_min, t_max, rec)) return false;
        rec.front_face = !rec.front_face;
        if (rec.inst < 0) rec.inst = gid;
        return true;
    }
    bool bounding_box(double time0, double time1, AABB &out) const override { // :289-291
        return ptr->bounding_box(time0, time1, out);
    }
};
struct Zoom : Hittable {
    HittablePtr ptr;
    double rate;
    bool bounding_box(double time0, double time1, AABB &out) const override { // :310-319
        AABB b;
        if (!ptr->bounding_box(time0, time1, b)) return false;
        out = AABB(b.min * rate, b.max * rate);
        return true;
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :321-330
        Ray moved_r(r.orig / rate, r.dir, r.tm); // direction left unscaled (Q5)
        if (!ptr->hit(moved_r, t_min, t_max, rec)) return false;
        rec.p = rec.p * rate;
        rec.set_face_normal(moved_r, rec.normal);
        if (rec.inst < 0) rec.inst = gid;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// hittable/sphere.rs
// ---------------------------------------------------------------------------------------------
void get_sphere_uv(Point3 p, double &u, double &v) { // sphere.rs:30-34
    double theta = std::acos(-p.y);
    double phi = std::atan2(-p.z, p.x) + PI;
    u = phi / (2. * PI);
    v = theta / PI;
}
struct Sphere : Hittable {
    Point3 center;
    double radius;
    const Material *mat_ptr;
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :39-66
        Vec3 oc = r.orig - center;
        double a = r.dir.length_sqr();
        double half_b = dot(oc, r.dir);
        double c = oc.length_sqr() - radius * radius;
        double discriminant = half_b * half_b - a * c;
        if (discriminant < 0.) return false;
        double sqrtd = std::sqrt(discriminant);
        double root = (-half_b - sqrtd) / a;
        if (root < t_min || t_max < root) {
            root = (-half_b + sqrtd) / a;
            if (root < t_min || t_max < root) return false;
        }
        Vec3 outward_normal = (r.at(root) - center) / radius;
        rec = HitRecord();
        get_sphere_uv(outward_normal, rec.u, rec.v);
        rec.p = r.at(root);
        rec.normal = outward_normal;
        rec.t = root;
        rec.mat_ptr = mat_ptr;
        rec.gid = gid;
        rec.set_face_normal(r, outward_normal);
        return true;
    }
    bool bounding_box(double, double, AABB &out) const override { // :68-73
        out = AABB(center - Vec3(radius, radius, radius), center + Vec3(radius, radius, radius));
        return true;
    }
    double pdf_value(Point3 o, Vec3 v) const override { // :75-83
        HitRecord rec;
        if (!hit(Ray(o, v, 0.), 0.001, INF, rec)) return 0.;
        double cos_max = std::sqrt(1. - radius * radius / (center - o).length_sqr());
        double solid_angle = 2. * PI * (1. - cos_max);
        return 1. / solid_angle;
    }
    Vec3 random(Point3 o) const override { // :85-90
        Vec3 direction = center - o;
        double dis_sqr = direction.length_sqr();
        Onb uvw = Onb::build_from_w(direction);
        return uvw.local_vec(random_to_sphere(radius, dis_sqr));
    }
};
struct MovingSphere : Hittable {
    Point3 center0, center1;
    double time0, time1, radius;
    const Material *mat_ptr;
    Point3 center(double time) const { // :124-127
        return center0 + (center1 - center0) * ((time - time0) / (time1 - time0));
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :138-165
        Vec3 oc = r.orig - center(r.tm);
        double a = r.dir.length_sqr();
        double half_b = dot(oc, r.dir);
        double c = oc.length_sqr() - radius * radius;
        double discriminant = half_b * half_b - a * c;
        if (discriminant < 0.) return false;
        double sqrtd = std::sqrt(discriminant);
        double root = (-half_b - sqrtd) / a;
        if (root < t_min || t_max < root) {
            root = (-half_b + sqrtd) / a;
            if (root < t_min || t_max < root) return false;
        }
        Vec3 outward_normal = (r.at(root) - center(r.tm)) / radius;
        rec = HitRecord();
        get_sphere_uv(outward_normal, rec.u, rec.v);
        rec.p = r.at(root);
        rec.normal = outward_normal;
        rec.t = root;
        rec.mat_ptr = mat_ptr;
        rec.gid = gid;
        rec.set_face_normal(r, outward_normal);
        return true;
    }
    bool bounding_box(double t0, double t1, AABB &out) const override { // :167-177
        Vec3 rv(radius, radius, radius);
        AABB box0(center(t0) - rv, center(t0) + rv);
        AABB box1(center(t1) - rv, center(t1) + rv);
        out = AABB::surrounding_box(box0, box1);
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// hittable/aarect.rs -- one class, `axis` = the constant axis (2: XY, 1: XZ, 0: YZ)
// ---------------------------------------------------------------------------------------------
struct AARect : Hittable {
    int axis; // thin axis
    int ia, ib; // in-plane axes in constructor order: XY -> (x,y), XZ -> (x,z), YZ -> (y,z)
    double a0, a1, b0, b1, k;
    const Material *mp;
    int32_t face = 0;
    AARect(int thin, double a0_, double a1_, double b0_, double b1_, double k_, const Material *m)
        : axis(thin), a0(a0_), a1(a1_), b0(b0_), b1(b1_), k(k_), mp(m) {
        if (thin == 2) { ia = 0; ib = 1; }
        else if (thin == 1) { ia = 0; ib = 2; }
        else { ia = 1; ib = 2; }
    }
    bool bounding_box(double, double, AABB &out) const override { // :40-45,123-128,206-211
        Point3 mn, mx;
        mn.at(axis) = k - 0.0001;
        mx.at(axis) = k + 0.0001;
        mn.at(ia) = a0;
        mx.at(ia) = a1;
        mn.at(ib) = b0;
        mx.at(ib) = b1;
        out = AABB(mn, mx);
        return true;
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :46-72 etc.
        double t = (k - r.orig[axis]) / r.dir[axis];
        if (t < t_min || t > t_max) return false; // NaN passes (Q2)
        double a = r.orig[ia] + t * r.dir[ia];
        double b = r.orig[ib] + t * r.dir[ib];
        if (a < a0 || a > a1 || b < b0 || b > b1) return false;
        Vec3 outward_normal(axis == 0 ? 1. : 0., axis == 1 ? 1. : 0., axis == 2 ? 1. : 0.);
        rec = HitRecord();
        rec.p = r.at(t);
        rec.normal = outward_normal;
        rec.t = t;
        rec.u = (a - a0) / (a1 - a0);
        rec.v = (b - b0) / (b1 - b0);
        rec.front_face = true;
        rec.mat_ptr = mp;
        rec.gid = gid;
        rec.face = face;
        rec.set_face_normal(r, outward_normal);
        return true;
    }
    double pdf_value(Point3 o, Vec3 v) const override { // :74-83,157-166,240-249
        HitRecord rec;
        if (!hit(Ray(o, v, 0.), 0.001, INF, rec)) return 0.;
        double area = (a1 - a0) * (b1 - b0);
        double dis_sqr = rec.t * rec.t * v.length_sqr();
        double cos = std::fabs(dot(v, rec.normal) / v.length());
        return dis_sqr / (cos * area);
    }
    Vec3 random(Point3 origin) const override { // :85-93,168-176,251-259
        Rng &rng = thread_rng();
        double ra = rng.gen_range(a0, a1);
        double rb = rng.gen_range(b0, b1);
        Point3 random_point;
        random_point.at(axis) = k;
        random_point.at(ia) = ra;
        random_point.at(ib) = rb;
        return random_point - origin;
    }
};

// hittable/boxes.rs
struct Boxes : Hittable {
    Point3 min, max;
    HittableList sides;
    Boxes(Point3 p0, Point3 p1, const Material *m, int32_t id) : min(p0), max(p1) { // :19-76
        gid = id;
        auto add = [&](int thin, double a0, double a1, double b0, double b1, double k) {
            auto r = std::make_shared<AARect>(thin, a0, a1, b0, b1, k, m);
            r->gid = id;
            r->face = (int32_t)sides.objects.size();
            sides.objects.push_back(r);
        };
        add(2, p0.x, p1.x, p0.y, p1.y, p1.z);
        add(2, p0.x, p1.x, p0.y, p1.y, p0.z);
        add(1, p0.x, p1.x, p0.z, p1.z, p1.y);
        add(1, p0.x, p1.x, p0.z, p1.z, p0.y);
        add(0, p0.y, p1.y, p0.z, p1.z, p1.x);
        add(0, p0.y, p1.y, p0.z, p1.z, p0.x);
    }
    bool bounding_box(double, double, AABB &out) const override { // :79-81
        out = AABB(min, max);
        return true;
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :82-84
        return sides.hit(r, t_min, t_max, rec);
    }
};

// hittable/triangle.rs
struct Triangle : Hittable {
    Point3 a, b, c;
    const Material *mp;
    Vec3 get_normal() const { return to_unit(cross(b - a, c - a)); } // :30-32
    bool inside(Point3 p) const {                                    // :33-46
        return dot(cross(c - a, p - a), cross(c - a, b - a)) >= 0. &&
               dot(cross(a - b, p - b), cross(a - b, c - b)) >= 0. &&
               dot(cross(b - c, p - c), cross(b - c, a - c)) >= 0.;
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :51-77
        Point3 origin = r.orig;
        Vec3 direction = r.dir;
        Vec3 n = get_normal();
        double t = dot(direction, n);
        t = dot(a - origin, n) / t;
        if (t != t || t < t_min || t > t_max) return false;
        Point3 p = origin + direction * t;
        if (!inside(p)) return false;
        double a1 = a.x - b.x, b1 = a.x - c.x, c1 = a.x - p.x;
        double a2 = a.y - b.y, b2 = a.y - c.y, c2 = a.y - p.y;
        double beta = (c1 * b2 - b1 * c2) / (a1 * b2 - b1 * a2);
        double gama = (a1 * c2 - a2 * c1) / (a1 * b2 - b1 * a2);
        rec = HitRecord();
        rec.p = p;
        rec.normal = n;
        rec.t = t;
        rec.u = beta;
        rec.v = gama;
        rec.front_face = true;
        rec.mat_ptr = mp;
        rec.gid = gid;
        rec.set_face_normal(r, n);
        return true;
    }
    bool bounding_box(double, double, AABB &out) const override { // :79-92
        out = AABB(Point3(std::fmin(a.x, std::fmin(b.x, c.x)), std::fmin(a.y, std::fmin(b.y, c.y)),
                          std::fmin(a.z, std::fmin(b.z, c.z))),
                   Point3(std::fmax(a.x, std::fmax(b.x, c.x)), std::fmax(a.y, std::fmax(b.y, c.y)),
                          std::fmax(a.z, std::fmax(b.z, c.z))));
        return true;
    }
};

// hittable/ring.rs
struct Ring : Hittable {
    double r, t, dis_min, dis_max;
    const Material *mat;
    Ring(double r_, double t_, const Material *m) : r(r_), t(t_), mat(m) { // :22-30
        dis_min = (r - t) * (r - t);
        dis_max = (r + t) * (r + t);
    }
    bool hit(const Ray &ray, double t_min, double t_max, HitRecord &rec) const override { // :34-52
        double tt = -ray.orig.y / ray.dir.y;
        if (tt != tt || tt < t_min || tt > t_max) return false;
        Point3 p = ray.at(tt);
        double d = p.x * p.x + p.z * p.z;
        if (d < dis_min || d > dis_max) return false;
        rec = HitRecord();
        rec.p = p;
        rec.normal = Vec3(0., 1., 0.);
        rec.t = tt;
        rec.u = 0.;
        rec.v = 0.;
        rec.front_face = false;
        rec.mat_ptr = mat;
        rec.gid = gid;
        rec.set_face_normal(ray, Vec3(0., 1., 0.));
        return true;
    }
    bool bounding_box(double, double, AABB &out) const override { // :54-62
        double thickness = 0.0001;
        double rr = r + t;
        out = AABB(Point3(-rr, -thickness, -rr), Point3(rr, thickness, rr));
        return true;
    }
};

// hittable/constantmedium.rs
struct ConstantMedium : Hittable {
    HittablePtr boundary;
    const Material *phase_function; // Isotropic
    double neg_inv_density;
    bool bounding_box(double t0, double t1, AABB &out) const override { // :46-48
        return boundary->bounding_box(t0, t1, out);
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :49-83
        HitRecord rec1, rec2;
        if (!boundary->hit(r, -INF, INF, rec1)) return false;
        if (!boundary->hit(r, rec1.t + 0.0001, INF, rec2)) return false;
        rec1.t = std::fmax(rec1.t, t_min);
        rec2.t = std::fmin(rec2.t, t_max);
        if (rec1.t >= rec2.t) return false;
        rec1.t = std::fmax(rec1.t, 0.);
        double ray_length = r.dir.length();
        double distance_inside_boundary = (rec2.t - rec1.t) * ray_length;
        double rnd = thread_rng().gen();
        double hit_distance = neg_inv_density * (std::log(rnd) / std::log(2.718281828459045)); // rnd.log(E)
        if (hit_distance > distance_inside_boundary) return false;
        rec = HitRecord();
        rec.t = rec1.t + hit_distance / ray_length;
        rec.p = r.at(rec1.t + hit_distance / ray_length);
        rec.normal = Vec3(1., 0., 0.);
        rec.u = 0.;
        rec.v = 0.;
        rec.front_face = true;
        rec.mat_ptr = phase_function;
        rec.gid = gid;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// hittable/bvh/mod.rs
// ---------------------------------------------------------------------------------------------
struct BvhNode : Hittable {
    AABB aabbox;
    HittablePtr left, right;
    static int box_compare(const HittablePtr &a, const HittablePtr &b, int axis) { // :19-29
        AABB box_a, box_b;
        a->bounding_box(0., 0., box_a);
        b->bounding_box(0., 0., box_b);
        if (box_a.min[axis] < box_b.min[axis]) return -1;
        if (box_a.min[axis] > box_b.min[axis]) return 1;
        return 0;
    }
    // :34-81.  Returns nullptr (the reference panics) on an empty list or a missing box.
    static std::shared_ptr<BvhNode> new_vec(std::vector<HittablePtr> objects, double time0, double time1) {
        int axis = (int)thread_rng().gen_index(3);
        size_t objects_span = objects.size();
        HittablePtr left, right;
        if (objects_span == 0) return nullptr;
        if (objects_span == 1) {
            HittablePtr obj0 = objects.back();
            left = obj0;
            right = obj0;
        } else if (objects_span == 2) {
            HittablePtr obj0 = objects.back();
            objects.pop_back();
            HittablePtr obj1 = objects.back();
            objects.pop_back();
            if (box_compare(obj0, obj1, axis) < 0) {
                left = obj0;
                right = obj1;
            } else {
                left = obj1;
                right = obj0;
            }
        } else {
            std::stable_sort(objects.begin(), objects.end(),
                             [axis](const HittablePtr &a, const HittablePtr &b) {
                                 return box_compare(a, b, axis) < 0;
                             }); // slice::sort_by is stable
            std::vector<HittablePtr> right_vec(objects.begin() + objects_span / 2, objects.end());
            objects.resize(objects_span / 2);
            left = new_vec(objects, time0, time1);
            right = new_vec(right_vec, time0, time1);
            if (!left || !right) return nullptr;
        }
        AABB left_box, right_box;
        if (!left->bounding_box(time0, time1, left_box)) return nullptr;
        if (!right->bounding_box(time0, time1, right_box)) return nullptr;
        auto node = std::make_shared<BvhNode>();
        node->aabbox = AABB::surrounding_box(left_box, right_box);
        node->left = left;
        node->right = right;
        return node;
    }
    bool hit(const Ray &r, double t_min, double t_max, HitRecord &rec) const override { // :86-101
        if (!aabbox.hit(r, t_min, t_max)) return false;
        HitRecord recl;
        if (left->hit(r, t_min, t_max, recl)) {
            HitRecord recr;
            if (right->hit(r, t_min, recl.t, recr))
                rec = recr;
            else
                rec = recl;
            return true;
        }
        return right->hit(r, t_min, t_max, rec);
    }
    bool bounding_box(double, double, AABB &out) const override { // :102-104
        out = aabbox;
        return true;
    }
};

// basic/camera.rs
struct Camera {
    Point3 origin, lower_left_corner;
    Vec3 horizontal, vertical, u, v, w;
    double lens_radius, time0, time1;
    Camera() {}
    Camera(Point3 lookfrom, Point3 lookat, Vec3 vup, double vfov, double aspect_ratio, double aperture,
           double focus_dist, double t0, double t1) { // camera.rs:24-62
        double theta = vfov * (PI / 180.0);
        double h = std::tan(theta / 2.);
        double viewport_height = 2. * h;
        double viewport_width = aspect_ratio * viewport_height;
        w = to_unit(lookfrom - lookat);
        u = to_unit(cross(vup, w));
        v = cross(w, u);
        origin = lookfrom;
        horizontal = u * viewport_width * focus_dist;
        vertical = v * viewport_height * focus_dist;
        lower_left_corner = origin - horizontal / 2. - vertical / 2. - w * focus_dist;
        lens_radius = aperture / 2.;
        time0 = t0;
        time1 = t1;
    }
    Ray get_ray(double s, double t) const { // camera.rs:64-73
        Vec3 rd = random_in_unit_disk() * lens_radius;
        Vec3 offset = u * rd.x + v * rd.y;
        Rng &rng = thread_rng();
        return Ray(origin + offset, lower_left_corner + horizontal * s + vertical * t - origin - offset,
                   rng.gen_range(time0, time1));
    }
};

} // namespace

// ---------------------------------------------------------------------------------------------
// scene: the pointer tree built from an rt_graph
// ---------------------------------------------------------------------------------------------
struct orc_scene {
    rt_graph g; // shallow copy; arrays below own the data
    std::vector<rt_texture> textures;
    std::vector<rt_material> materials;
    std::vector<rt_object> objects;
    std::vector<int32_t> list_items, light_ids;
    std::vector<rt_perlin> perlins;
    std::vector<rt_image> images;
    std::vector<uint8_t> texels;

    std::vector<std::shared_ptr<Texture>> tex;
    std::vector<std::shared_ptr<Material>> mat;
    std::vector<HittablePtr> obj;
    HittablePtr world;
    HittableList lights;
    Camera cam;
    Color background;
};

namespace {

std::shared_ptr<Texture> build_texture(orc_scene &s, int id, int depth = 0) {
    if (id < 0 || id >= (int)s.textures.size() || depth > 16) throw std::string("bad texture id");
    if (s.tex[id]) return s.tex[id];
    const rt_texture &t = s.textures[id];
    std::shared_ptr<Texture> out;
    switch (t.kind) {
    case RT_TEX_SOLID:
        out = std::make_shared<SolidColor>(Color(t.p[0], t.p[1], t.p[2]));
        break;
    case RT_TEX_CHECKER: {
        auto c = std::make_shared<CheckerTexture>();
        c->odd = build_texture(s, t.a, depth + 1);
        c->even = build_texture(s, t.b, depth + 1);
        out = c;
        break;
    }
    case RT_TEX_NOISE: {
        if (t.a < 0 || t.a >= (int)s.perlins.size()) throw std::string("bad perlin id");
        auto n = std::make_shared<NoiseTexture>();
        n->noise.t = &s.perlins[t.a];
        n->scale = t.p[0];
        out = n;
        break;
    }
    case RT_TEX_IMAGE: {
        auto im = std::make_shared<ImageTexture>();
        if (t.a >= 0) {
            if (t.a >= (int)s.images.size()) throw std::string("bad image id");
            const rt_image &ri = s.images[t.a];
            im->width = (uint32_t)ri.width;
            im->height = (uint32_t)ri.height;
            im->pixel_color = s.texels.data() + ri.offset;
        }
        out = im;
        break;
    }
    default:
        throw std::string("bad texture kind");
    }
    s.tex[id] = out;
    return out;
}

const Material *build_material(orc_scene &s, int id) {
    if (id < 0 || id >= (int)s.materials.size()) throw std::string("bad material id");
    if (s.mat[id]) return s.mat[id].get();
    const rt_material &m = s.materials[id];
    std::shared_ptr<Material> out;
    switch (m.kind) {
    case RT_MAT_LAMBERTIAN: {
        auto l = std::make_shared<Lambertian>();
        l->albedo = build_texture(s, m.tex);
        out = l;
        break;
    }
    case RT_MAT_METAL: {
        auto l = std::make_shared<Metal>();
        l->albedo = Color(m.p[0], m.p[1], m.p[2]);
        l->fuzz = m.p[3] < 1. ? m.p[3] : 1.; // Metal::new, material/mod.rs:76-81
        out = l;
        break;
    }
    case RT_MAT_DIELECTRIC: {
        auto l = std::make_shared<Dielectric>();
        l->ir = m.p[0];
        out = l;
        break;
    }
    case RT_MAT_DIFFUSE_LIGHT: {
        auto l = std::make_shared<DiffuseLight>();
        l->emit = build_texture(s, m.tex);
        out = l;
        break;
    }
    case RT_MAT_ISOTROPIC: {
        auto l = std::make_shared<Isotropic>();
        l->albedo = build_texture(s, m.tex);
        out = l;
        break;
    }
    default:
        throw std::string("bad material kind");
    }
    out->mid = id;
    s.mat[id] = out;
    return out.get();
}

HittablePtr build_object(orc_scene &s, int id, int depth = 0) {
    if (id < 0 || id >= (int)s.objects.size()) throw std::string("bad object id");
    if (depth > 4096) throw std::string("object graph too deep (cycle?)");
    if (s.obj[id]) return s.obj[id];
    const rt_object &o = s.objects[id];
    const double *p = o.p;
    HittablePtr out;
    switch (o.kind) {
    case RT_OBJ_SPHERE: {
        auto h = std::make_shared<Sphere>();
        h->center = Point3(p[0], p[1], p[2]);
        h->radius = p[3];
        h->mat_ptr = build_material(s, o.mat);
        out = h;
        break;
    }
    case RT_OBJ_MOVING_SPHERE: {
        auto h = std::make_shared<MovingSphere>();
        h->center0 = Point3(p[0], p[1], p[2]);
        h->center1 = Point3(p[3], p[4], p[5]);
        h->time0 = p[6];
        h->time1 = p[7];
        h->radius = p[8];
        h->mat_ptr = build_material(s, o.mat);
        out = h;
        break;
    }
    case RT_OBJ_XY_RECT:
    case RT_OBJ_XZ_RECT:
    case RT_OBJ_YZ_RECT: {
        int thin = o.kind == RT_OBJ_XY_RECT ? 2 : (o.kind == RT_OBJ_XZ_RECT ? 1 : 0);
        out = std::make_shared<AARect>(thin, p[0], p[1], p[2], p[3], p[4], build_material(s, o.mat));
        break;
    }
    case RT_OBJ_BOX:
        out = std::make_shared<Boxes>(Point3(p[0], p[1], p[2]), Point3(p[3], p[4], p[5]),
                                      build_material(s, o.mat), id);
        break;
    case RT_OBJ_TRIANGLE: {
        auto h = std::make_shared<Triangle>();
        h->a = Point3(p[0], p[1], p[2]);
        h->b = Point3(p[3], p[4], p[5]);
        h->c = Point3(p[6], p[7], p[8]);
        h->mp = build_material(s, o.mat);
        out = h;
        break;
    }
    case RT_OBJ_RING:
        out = std::make_shared<Ring>(p[0], p[1], build_material(s, o.mat));
        break;
    case RT_OBJ_MEDIUM: {
        auto h = std::make_shared<ConstantMedium>();
        h->boundary = build_object(s, o.c0, depth + 1);
        h->phase_function = build_material(s, o.mat);
        h->neg_inv_density = -1. / p[0]; // constantmedium.rs:29,39
        out = h;
        break;
    }
    case RT_OBJ_TRANSLATE: {
        auto h = std::make_shared<Translate>();
        h->ptr = build_object(s, o.c0, depth + 1);
        h->offset = Vec3(p[0], p[1], p[2]);
        out = h;
        break;
    }
    case RT_OBJ_ROTATE_Y:
        out = std::make_shared<RotateY>(build_object(s, o.c0, depth + 1), p[0]);
        break;
    case RT_OBJ_FLIP_FACE: {
        auto h = std::make_shared<FlipFace>();
        h->ptr = build_object(s, o.c0, depth + 1);
        out = h;
        break;
    }
    case RT_OBJ_ZOOM: {
        auto h = std::make_shared<Zoom>();
        h->ptr = build_object(s, o.c0, depth + 1);
        h->rate = p[0];
        out = h;
        break;
    }
    case RT_OBJ_BVH_NODE: {
        auto h = std::make_shared<BvhNode>();
        h->left = build_object(s, o.c0, depth + 1);
        h->right = build_object(s, o.c1, depth + 1);
        h->aabbox = AABB(Point3(p[0], p[1], p[2]), Point3(p[3], p[4], p[5]));
        out = h;
        break;
    }
    case RT_OBJ_LIST: {
        auto h = std::make_shared<HittableList>();
        if (o.c0 < 0 || o.c1 < 0 || (size_t)o.c0 + (size_t)o.c1 > s.list_items.size())
            throw std::string("bad list range");
        for (int i = 0; i < o.c1; ++i) h->objects.push_back(build_object(s, s.list_items[o.c0 + i], depth + 1));
        out = h;
        break;
    }
    default:
        throw std::string("bad object kind");
    }
    out->gid = id;
    s.obj[id] = out;
    return out;
}

// main.rs:233-278
Color ray_color(const Ray &r, Color background, const Hittable &world, const HittableList &lights, int depth) {
    if (depth <= 0) return Color(0., 0., 0.);
    HitRecord rec;
    ++tl_rays;
    if (!world.hit(r, 0.001, F64_MAX, rec)) return background;
    Color emitted = rec.mat_ptr->emitted(r, rec, rec.u, rec.v, rec.p);
    ScatterRecord srec;
    if (!rec.mat_ptr->scatter(r, rec, srec)) return emitted;
    if (srec.has_specular)
        return srec.attenuation * ray_color(srec.specular_ray, background, world, lights, depth - 1);
    CosPdf cos_pdf(srec.pdf_w);
    Vec3 direction;
    double pdf_val;
    if (lights.objects.empty()) {
        // Q12: the shipped code panics on an empty `lights`; books 1-2 sample the
        // cosine pdf alone.  Decision shared with the CUDA path: empty lights -> CosPdf only.
        direction = cos_pdf.generate();
        pdf_val = cos_pdf.value(direction);
    } else {
        // MixturePdf::generate / value, pdf.rs:95-104, over HittablePdf(lights, rec.p) and CosPdf
        if (thread_rng().gen_range(0.0, 1.0) < 0.5)
            direction = lights.random(rec.p);
        else
            direction = cos_pdf.generate();
        pdf_val = 0.5 * lights.pdf_value(rec.p, direction) + 0.5 * cos_pdf.value(direction);
    }
    Ray scattered(rec.p, direction, r.tm);
    return emitted + srec.attenuation * rec.mat_ptr->scattering_pdf(r, rec, scattered) *
                         ray_color(scattered, background, world, lights, depth - 1) / pdf_val;
}

Ray camera_sample(const orc_scene &s, int w, int h, int x, int y) { // main.rs:145-149
    Rng &rng = thread_rng();
    double rand_u = rng.gen();
    double rand_v = rng.gen();
    double u = ((double)x + rand_u) / (double)(w - 1);
    double v = ((double)y + rand_v) / (double)(h - 1);
    return s.cam.get_ray(u, v);
}

void fill_hit(const HitRecord &rec, bool hit, uint32_t pos, orc_hit &o) {
    std::memset(&o, 0, sizeof o);
    o.hit = hit ? 1 : 0;
    o.gid = -1;
    o.inst = -1;
    o.mat = -1;
    o.rng_pos = pos;
    if (!hit) return;
    o.t = rec.t;
    o.p[0] = rec.p.x; o.p[1] = rec.p.y; o.p[2] = rec.p.z;
    o.n[0] = rec.normal.x; o.n[1] = rec.normal.y; o.n[2] = rec.normal.z;
    o.u = rec.u;
    o.v = rec.v;
    o.front_face = rec.front_face ? 1 : 0;
    o.gid = rec.gid;
    o.face = rec.face;
    o.inst = rec.inst;
    o.mat = rec.mat_ptr ? rec.mat_ptr->mid : -1;
}

void listing(const Hittable *h, std::vector<int32_t> &out, std::vector<double> *boxes) {
    if (const BvhNode *n = dynamic_cast<const BvhNode *>(h)) {
        out.push_back(-1);
        if (boxes) {
            const double b[6] = {n->aabbox.min.x, n->aabbox.min.y, n->aabbox.min.z,
                                 n->aabbox.max.x, n->aabbox.max.y, n->aabbox.max.z};
            boxes->insert(boxes->end(), b, b + 6);
        }
        listing(n->left.get(), out, boxes);
        listing(n->right.get(), out, boxes);
    } else {
        out.push_back(h->gid);
    }
}

} // namespace

// ---------------------------------------------------------------------------------------------
// C interface
// ---------------------------------------------------------------------------------------------
extern "C" {

const char *orc_last_error(void) { return g_err.c_str(); }

orc_scene *orc_scene_create(const rt_graph *g) {
    orc_scene *s = new orc_scene();
    try {
        s->g = *g;
        s->textures.assign(g->textures, g->textures + g->n_textures);
        s->materials.assign(g->materials, g->materials + g->n_materials);
        s->objects.assign(g->objects, g->objects + g->n_objects);
        s->list_items.assign(g->list_items, g->list_items + g->n_list_items);
        s->light_ids.assign(g->lights, g->lights + g->n_lights);
        s->perlins.assign(g->perlins, g->perlins + g->n_perlins);
        s->images.assign(g->images, g->images + g->n_images);
        s->texels.assign(g->texels, g->texels + g->n_texel_bytes);
        s->tex.resize(s->textures.size());
        s->mat.resize(s->materials.size());
        s->obj.resize(s->objects.size());
        for (int i = 0; i < (int)s->textures.size(); ++i) build_texture(*s, i);
        for (int i = 0; i < (int)s->materials.size(); ++i) build_material(*s, i);
        // children always precede parents in a graph made by constructor calls; build in id
        // order so that deep chains do not recurse
        for (int i = 0; i < (int)s->objects.size(); ++i) build_object(*s, i);
        if (g->world >= 0) s->world = build_object(*s, g->world);
        for (int32_t id : s->light_ids) s->lights.objects.push_back(build_object(*s, id));
        const rt_camera &c = g->camera;
        s->cam = Camera(Point3(c.lookfrom[0], c.lookfrom[1], c.lookfrom[2]),
                        Point3(c.lookat[0], c.lookat[1], c.lookat[2]), Vec3(c.vup[0], c.vup[1], c.vup[2]),
                        c.vfov, c.aspect_ratio, c.aperture, c.focus_dist, c.time0, c.time1);
        s->background = Color(g->background[0], g->background[1], g->background[2]);
    } catch (const std::string &e) {
        g_err = e;
        delete s;
        return nullptr;
    }
    return s;
}

void orc_scene_destroy(orc_scene *s) { delete s; }

int orc_trace_batch(const orc_scene *s, int64_t n, const orc_ray *rays, const uint64_t *keys, orc_hit *out) {
    if (!s || !s->world) {
        g_err = "no world";
        return -1;
    }
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    if (n < 4096) nt = 1;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([=]() {
            for (int64_t i = (n * t) / nt; i < (n * (t + 1)) / nt; ++i) {
                Rng rng(keys ? keys[i] : (uint64_t)i, 0, 0, DOMAIN_BATCH);
                tl_rng = &rng;
                const orc_ray &q = rays[i];
                Ray r(Point3(q.o[0], q.o[1], q.o[2]), Vec3(q.d[0], q.d[1], q.d[2]), q.tm);
                HitRecord rec;
                bool h = s->world->hit(r, q.t_min, q.t_max, rec);
                fill_hit(rec, h, rng.pos, out[i]);
            }
            tl_rng = nullptr;
        });
    for (auto &th : pool) th.join();
    return 0;
}

int orc_render_region(const orc_scene *s, int w, int h, int x0, int x1, int y0, int y1, int spp_begin,
                      int spp_end, int max_depth, uint64_t seed, int threads, double *accum,
                      orc_stats *stats) {
    if (!s || !s->world || w < 2 || h < 2 || threads < 1) {
        g_err = "bad render arguments";
        return -1;
    }
    int rows = y1 - y0;
    // main.rs:93-99: Fisher-Yates over the scanlines
    std::vector<int> random_line_id(rows);
    {
        Rng rng(seed, 0x4c494e45u /* "LINE" */, 0, DOMAIN_SCENE);
        for (int i = 0; i < rows; ++i) {
            random_line_id[i] = y0 + i;
            uint32_t target = rng.gen_index((uint32_t)i + 1);
            std::swap(random_line_id[i], random_line_id[target]);
        }
    }
    if (threads > rows) threads = std::max(rows, 1);
    int section = rows / threads; // main.rs:41
    std::vector<uint64_t> rays(threads, 0);
    auto t_begin = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (int tid = 0; tid < threads; ++tid) {
        int line_beg = tid * section;
        int line_end = tid == threads - 1 ? rows : line_beg + section; // main.rs:112-116
        pool.emplace_back([=, &rays, &random_line_id]() {
            tl_rays = 0;
            for (int y_id = line_beg; y_id < line_end; ++y_id)
                for (int x = x0; x < x1; ++x) { // main.rs:140-153
                    int y = random_line_id[y_id];
                    Color pixel_color(0., 0., 0.);
                    for (int i = spp_begin; i < spp_end; ++i) {
                        Rng rng(seed, (uint32_t)(y * w + x), (uint32_t)i, DOMAIN_PATH);
                        tl_rng = &rng;
                        Ray r = camera_sample(*s, w, h, x, y);
                        pixel_color = pixel_color + ray_color(r, s->background, *s->world, s->lights, max_depth);
                    }
                    double *px = accum + 3 * ((size_t)y * w + x);
                    px[0] = pixel_color.x;
                    px[1] = pixel_color.y;
                    px[2] = pixel_color.z;
                }
            tl_rng = nullptr;
            rays[tid] = tl_rays;
        });
    }
    for (auto &th : pool) th.join();
    auto t_end = std::chrono::steady_clock::now();
    if (stats) {
        stats->rays = 0;
        for (uint64_t r : rays) stats->rays += r;
        stats->samples = (uint64_t)rows * (uint64_t)(x1 - x0) * (uint64_t)(spp_end - spp_begin);
        stats->seconds = std::chrono::duration<double>(t_end - t_begin).count();
    }
    return 0;
}

int orc_render(const orc_scene *s, int w, int h, int spp_begin, int spp_end, int max_depth, uint64_t seed,
               int threads, double *accum, orc_stats *stats) {
    return orc_render_region(s, w, h, 0, w, 0, h, spp_begin, spp_end, max_depth, seed, threads, accum, stats);
}

int orc_sample(const orc_scene *s, int w, int h, int x, int y, int k, int max_depth, uint64_t seed,
               double rgb[3]) {
    Rng rng(seed, (uint32_t)(y * w + x), (uint32_t)k, DOMAIN_PATH);
    tl_rng = &rng;
    tl_rays = 0;
    Ray r = camera_sample(*s, w, h, x, y);
    Color c = ray_color(r, s->background, *s->world, s->lights, max_depth);
    tl_rng = nullptr;
    rgb[0] = c.x;
    rgb[1] = c.y;
    rgb[2] = c.z;
    return (int)tl_rays;
}

int orc_camera_ray(const orc_scene *s, int w, int h, int x, int y, int k, uint64_t seed, orc_ray *out,
                   uint32_t *rng_pos) {
    Rng rng(seed, (uint32_t)(y * w + x), (uint32_t)k, DOMAIN_PATH);
    tl_rng = &rng;
    Ray r = camera_sample(*s, w, h, x, y);
    tl_rng = nullptr;
    out->o[0] = r.orig.x; out->o[1] = r.orig.y; out->o[2] = r.orig.z;
    out->d[0] = r.dir.x; out->d[1] = r.dir.y; out->d[2] = r.dir.z;
    out->tm = r.tm;
    out->t_min = 0.001;
    out->t_max = F64_MAX;
    if (rng_pos) *rng_pos = rng.pos;
    return 0;
}

} // extern "C"

namespace {
// ray_color with the rays of one bounce level recorded (same control flow as ray_color above)
struct RayDump {
    int bounce;
    int64_t cap, n;
    orc_ray *out;
};
Color ray_color_dump(const Ray &r, const orc_scene &s, int depth, int level, RayDump &d) {
    if (depth <= 0) return Color(0., 0., 0.);
    if (level == d.bounce && d.n < d.cap) {
        orc_ray &q = d.out[d.n++];
        q.o[0] = r.orig.x; q.o[1] = r.orig.y; q.o[2] = r.orig.z;
        q.d[0] = r.dir.x; q.d[1] = r.dir.y; q.d[2] = r.dir.z;
        q.tm = r.tm;
        q.t_min = 0.001;
        q.t_max = F64_MAX;
    }
    if (level >= d.bounce) return Color(0., 0., 0.);
    HitRecord rec;
    if (!s.world->hit(r, 0.001, F64_MAX, rec)) return s.background;
    ScatterRecord srec;
    if (!rec.mat_ptr->scatter(r, rec, srec)) return Color(0., 0., 0.);
    if (srec.has_specular) return ray_color_dump(srec.specular_ray, s, depth - 1, level + 1, d);
    CosPdf cos_pdf(srec.pdf_w);
    Vec3 direction;
    if (s.lights.objects.empty())
        direction = cos_pdf.generate();
    else if (thread_rng().gen_range(0.0, 1.0) < 0.5)
        direction = s.lights.random(rec.p);
    else
        direction = cos_pdf.generate();
    return ray_color_dump(Ray(rec.p, direction, r.tm), s, depth - 1, level + 1, d);
}
} // namespace

extern "C" {

int64_t orc_dump_rays(const orc_scene *s, int w, int h, int spp, int max_depth, uint64_t seed, int bounce,
                      int64_t cap, orc_ray *out) {
    /* The first `cap` rays of the given bounce in (y, x, sample) order.  Rows are traced by all host threads, a band of rows
     * per round, each row into its own buffer; the buffers are appended in row order, so the result does not depend on
     * the number of threads (tests dump 2^20 rays per scene and bounce). */
    unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
    int64_t n = 0;
    for (int y0 = 0; y0 < h && n < cap; y0 += (int)nt) {
        int rows = std::min<int>((int)nt, h - y0);
        std::vector<std::vector<orc_ray>> buf((size_t)rows);
        std::vector<std::thread> pool;
        for (int k = 0; k < rows; ++k)
            pool.emplace_back([&, k]() {
                int y = y0 + k;
                int64_t room = cap - n; /* no row needs more than what is still missing */
                buf[k].resize((size_t)std::min<int64_t>(room, (int64_t)w * spp));
                RayDump d{bounce, (int64_t)buf[k].size(), 0, buf[k].data()};
                for (int x = 0; x < w && d.n < d.cap; ++x)
                    for (int i = 0; i < spp && d.n < d.cap; ++i) {
                        Rng rng(seed, (uint32_t)(y * w + x), (uint32_t)i, DOMAIN_PATH);
                        tl_rng = &rng;
                        Ray r = camera_sample(*s, w, h, x, y);
                        ray_color_dump(r, *s, max_depth, 0, d);
                    }
                tl_rng = nullptr;
                buf[k].resize((size_t)d.n);
            });
        for (auto &th : pool) th.join();
        for (int k = 0; k < rows && n < cap; ++k) {
            int64_t take = std::min<int64_t>((int64_t)buf[k].size(), cap - n);
            std::memcpy(out + n, buf[k].data(), (size_t)take * sizeof(orc_ray));
            n += take;
        }
    }
    return n;
}

void orc_write_color(const double rgb_sum[3], int spp, uint8_t out[3]) { // main.rs:280-299
    for (int c = 0; c < 3; ++c) {
        double r = rgb_sum[c];
        if (r != r) r = 0.;
        double x = std::sqrt(r / (double)spp);
        // f64::clamp(0.0, 0.999): NaN stays NaN; `as u8` saturates and maps NaN to 0
        if (x < 0.0) x = 0.0;
        if (x > 0.999) x = 0.999;
        double f = std::floor(x * 255.999);
        out[c] = (f != f) ? 0 : (f <= 0. ? 0 : (f >= 255. ? 255 : (uint8_t)f));
    }
}

double orc_lights_pdf_value(const orc_scene *s, const double o[3], const double v[3]) {
    return s->lights.pdf_value(Point3(o[0], o[1], o[2]), Vec3(v[0], v[1], v[2]));
}
void orc_lights_random(const orc_scene *s, const double o[3], uint64_t key, double out[3]) {
    Rng rng(key, 0, 0, DOMAIN_BATCH);
    tl_rng = &rng;
    Vec3 d = s->lights.random(Point3(o[0], o[1], o[2]));
    tl_rng = nullptr;
    out[0] = d.x;
    out[1] = d.y;
    out[2] = d.z;
}
void orc_texture_value(const orc_scene *s, int tex, double u, double v, const double p[3], double out[3]) {
    Color c = s->tex[tex]->value(u, v, Point3(p[0], p[1], p[2]));
    out[0] = c.x;
    out[1] = c.y;
    out[2] = c.z;
}
double orc_perlin_noise(const orc_scene *s, int perlin, const double p[3]) {
    Perlin pn{&s->perlins[perlin]};
    return pn.noise(Point3(p[0], p[1], p[2]));
}
double orc_perlin_turb(const orc_scene *s, int perlin, const double p[3], int depth) {
    Perlin pn{&s->perlins[perlin]};
    return pn.turb(Point3(p[0], p[1], p[2]), depth);
}
double orc_reflectance(double cos, double ref_idx) { return reflectance(cos, ref_idx); }
int orc_aabb_hit(const double bmin[3], const double bmax[3], const orc_ray *r) {
    AABB b(Point3(bmin[0], bmin[1], bmin[2]), Point3(bmax[0], bmax[1], bmax[2]));
    return b.hit(Ray(Point3(r->o[0], r->o[1], r->o[2]), Vec3(r->d[0], r->d[1], r->d[2]), r->tm), r->t_min,
                 r->t_max)
               ? 1
               : 0;
}
void orc_camera_derive(rt_camera *c) {
    Camera cam(Point3(c->lookfrom[0], c->lookfrom[1], c->lookfrom[2]),
               Point3(c->lookat[0], c->lookat[1], c->lookat[2]), Vec3(c->vup[0], c->vup[1], c->vup[2]),
               c->vfov, c->aspect_ratio, c->aperture, c->focus_dist, c->time0, c->time1);
    auto put = [](double *d, Vec3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
    put(c->origin, cam.origin);
    put(c->lower_left_corner, cam.lower_left_corner);
    put(c->horizontal, cam.horizontal);
    put(c->vertical, cam.vertical);
    put(c->u, cam.u);
    put(c->v, cam.v);
    put(c->w, cam.w);
    c->lens_radius = cam.lens_radius;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox(ctr, key, out); }
double orc_stream_gen(uint64_t seed, uint32_t s0, uint32_t s1, uint32_t domain, uint32_t pos) {
    Rng rng(seed, s0, s1, domain, pos);
    return rng.gen();
}
uint32_t orc_stream_index(uint64_t seed, uint32_t s0, uint32_t s1, uint32_t domain, uint32_t pos, uint32_t n,
                          uint32_t *pos_after) {
    Rng rng(seed, s0, s1, domain, pos);
    uint32_t r = rng.gen_index(n);
    if (pos_after) *pos_after = rng.pos;
    return r;
}

int64_t orc_bvh_build(const orc_scene *s, const int32_t *ids, int32_t n, double time0, double time1,
                      uint64_t seed, uint32_t s0, uint32_t pos, int32_t *out, int64_t cap, double *boxes,
                      uint32_t *pos_after) {
    std::vector<HittablePtr> objs;
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || ids[i] >= (int)s->obj.size() || !s->obj[ids[i]]) return -1;
        objs.push_back(s->obj[ids[i]]);
    }
    Rng rng(seed, s0, 0, DOMAIN_SCENE, pos);
    tl_rng = &rng;
    std::shared_ptr<BvhNode> root = BvhNode::new_vec(objs, time0, time1);
    tl_rng = nullptr;
    if (pos_after) *pos_after = rng.pos;
    if (!root) return -1;
    std::vector<int32_t> l;
    std::vector<double> b;
    listing(root.get(), l, boxes ? &b : nullptr);
    if ((int64_t)l.size() > cap) return -1;
    std::copy(l.begin(), l.end(), out);
    if (boxes) std::copy(b.begin(), b.end(), boxes);
    return (int64_t)l.size();
}

int64_t orc_bvh_listing(const orc_scene *s, int32_t root, int32_t *out, int64_t cap, double *boxes) {
    if (root < 0 || root >= (int)s->obj.size()) return -1;
    std::vector<int32_t> l;
    std::vector<double> b;
    listing(s->obj[root].get(), l, boxes ? &b : nullptr);
    if ((int64_t)l.size() > cap) return -1;
    std::copy(l.begin(), l.end(), out);
    if (boxes) std::copy(b.begin(), b.end(), boxes);
    return (int64_t)l.size();
}

} // extern "C"
