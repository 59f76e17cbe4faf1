/*
 * c_api.cpp -- C entry points over rt::Scene / rt::lower for ctypes (tests, bench.py).
 * Exceptions never cross the boundary: failures return -1 / NULL and set rtg_last_error().
 */
#include "rt_host.h"

#include <cstring>

struct rtg_scene {
    rt::Scene sc;
    explicit rtg_scene(uint64_t seed) : sc(seed) {}
};
struct rtg_lowered {
    rt::Lowered l;
};

static thread_local std::string g_err;

#define RTG_TRY(expr_on_fail)                                                                                      \
    catch (const rt::Error &e) {                                                                                   \
        g_err = e.what;                                                                                            \
        return expr_on_fail;                                                                                       \
    }                                                                                                              \
    catch (const std::exception &e) {                                                                              \
        g_err = e.what();                                                                                          \
        return expr_on_fail;                                                                                       \
    }

extern "C" {

const char *rtg_last_error(void) { return g_err.c_str(); }

rtg_scene *rtg_new(uint64_t seed) {
    try {
        return new rtg_scene(seed);
    }
    RTG_TRY(nullptr)
}
void rtg_free(rtg_scene *s) { delete s; }

int32_t rtg_add_texture(rtg_scene *s, int kind, int a, int b, const double *p, int np) {
    try {
        return s->sc.add_texture(kind, a, b, p, np);
    }
    RTG_TRY(-1)
}
int32_t rtg_add_noise_texture(rtg_scene *s, double scale) {
    try {
        return rt::NoiseTexture::new_(s->sc, scale).id;
    }
    RTG_TRY(-1)
}
int32_t rtg_add_image_texture_file(rtg_scene *s, const char *path) {
    try {
        return rt::ImageTexture::new_(s->sc, path).id;
    }
    RTG_TRY(-1)
}
int32_t rtg_add_image_texture_rgb8(rtg_scene *s, int w, int h, const uint8_t *top_down_rgb) {
    try {
        int32_t im = s->sc.add_image_rgb8(w, h, top_down_rgb);
        return s->sc.add_texture(RT_TEX_IMAGE, im, -1, nullptr, 0);
    }
    RTG_TRY(-1)
}
int32_t rtg_add_material(rtg_scene *s, int kind, int tex, const double *p, int np) {
    try {
        return s->sc.add_material(kind, tex, p, np);
    }
    RTG_TRY(-1)
}
int32_t rtg_add_object(rtg_scene *s, int kind, int mat, int c0, int c1, const double *p, int np) {
    try {
        if (kind == RT_OBJ_LIST || kind == RT_OBJ_BVH_NODE) throw rt::Error{"use rtg_add_list / rtg_bvh"};
        return s->sc.add_object(kind, mat, c0, c1, p, np);
    }
    RTG_TRY(-1)
}
int32_t rtg_add_list(rtg_scene *s, const int32_t *ids, int n) {
    try {
        return s->sc.add_list(ids, n);
    }
    RTG_TRY(-1)
}
int32_t rtg_bvh(rtg_scene *s, const int32_t *ids, int n, double time0, double time1) {
    try {
        for (int i = 0; i < n; ++i) s->sc.check_obj(ids[i]);
        return s->sc.bvh_new_vec(std::vector<int32_t>(ids, ids + n), time0, time1);
    }
    RTG_TRY(-1)
}
int32_t rtg_load_obj_triangles(rtg_scene *s, const char *path, int mat, int32_t *first_id, int32_t *count) {
    try {
        std::vector<rt::ObjMesh> models = rt::load_obj(path);
        int32_t first = (int32_t)s->sc.objects.size(), n = 0;
        for (const rt::ObjMesh &m : models)
            for (size_t v = 0; v < m.indices.size() / 3; ++v) {
                double p[9];
                for (int c = 0; c < 3; ++c)
                    for (int k = 0; k < 3; ++k) p[3 * c + k] = (double)m.positions[3 * m.indices[3 * v + c] + k];
                s->sc.add_object(RT_OBJ_TRIANGLE, mat, -1, -1, p, 9);
                ++n;
            }
        *first_id = first;
        *count = n;
        return 0;
    }
    RTG_TRY(-1)
}
int rtg_bounding_box(rtg_scene *s, int32_t obj, double time0, double time1, double out6[6]) {
    try {
        rt::Box3 b;
        if (!s->sc.bounding_box(obj, time0, time1, b)) return 0;
        for (int c = 0; c < 3; ++c) {
            out6[c] = b.mn[c];
            out6[3 + c] = b.mx[c];
        }
        return 1;
    }
    RTG_TRY(-1)
}
int rtg_set_world(rtg_scene *s, int32_t obj) {
    try {
        s->sc.check_obj(obj);
        s->sc.world = obj;
        return 0;
    }
    RTG_TRY(-1)
}
int rtg_set_lights(rtg_scene *s, const int32_t *ids, int n) {
    try {
        for (int i = 0; i < n; ++i) s->sc.check_obj(ids[i]);
        s->sc.lights.assign(ids, ids + n);
        return 0;
    }
    RTG_TRY(-1)
}
int rtg_set_camera(rtg_scene *s, const double lookfrom[3], const double lookat[3], const double vup[3],
                   double vfov, double aspect_ratio, double aperture, double focus_dist, double time0,
                   double time1) {
    try {
        s->sc.set_camera(rt::Vec3(lookfrom[0], lookfrom[1], lookfrom[2]), rt::Vec3(lookat[0], lookat[1], lookat[2]),
                         rt::Vec3(vup[0], vup[1], vup[2]), vfov, aspect_ratio, aperture, focus_dist, time0, time1);
        return 0;
    }
    RTG_TRY(-1)
}
int rtg_set_background(rtg_scene *s, const double rgb[3]) {
    for (int c = 0; c < 3; ++c) s->sc.background[c] = rgb[c];
    return 0;
}
int rtg_build_config(rtg_scene *s, const char *name, double aspect_ratio, const char *asset_dir) {
    try {
        rt::SceneOptions o;
        o.asset_dir = asset_dir ? asset_dir : "";
        if (!o.asset_dir.empty() && o.asset_dir.back() != '/') o.asset_dir += '/';
        rt::build_config(s->sc, name, aspect_ratio, o);
        return 0;
    }
    RTG_TRY(-1)
}
uint32_t rtg_rng_pos(const rtg_scene *s) { return s->sc.rng.pos; }
/* include/rt_rng.h as the product compiles it (the same header the CUDA kernels include): one Philox4x32-10 block, and
 * n slots / gen() / gen_range() / gen_index() draws of stream (seed; a, b, domain) from slot `start`.
 * tests/test_rng.py pins them against the Random123 known-answer vectors and against rand 0.8.5's published recipes. */
void rtg_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    rt_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
void rtg_rng_stream(uint64_t seed, uint32_t a, uint32_t b, uint32_t domain, uint32_t start, int n, uint64_t *slots, double *gen,
                    double *range_m1_1, uint32_t *index, uint32_t index_n) {
    for (int pass = 0; pass < 4; ++pass) {
        rt_rng r;
        r.init(seed, a, b, domain, start);
        for (int i = 0; i < n; ++i) {
            if (pass == 0 && slots) slots[i] = r.next_u64();
            if (pass == 1 && gen) gen[i] = r.gen();
            if (pass == 2 && range_m1_1) range_m1_1[i] = r.gen_range(-1.0, 1.0);
            if (pass == 3 && index) index[i] = r.gen_index(index_n);
        }
    }
}
int rtg_graph(const rtg_scene *s, rt_graph *out) {
    *out = s->sc.view();
    return 0;
}

rtg_lowered *rtg_lower(const rt_graph *g, int elide_singletons) {
    try {
        rtg_lowered *l = new rtg_lowered();
        try {
            l->l = rt::lower(*g, elide_singletons != 0);
        } catch (...) {
            delete l;
            throw;
        }
        return l;
    }
    RTG_TRY(nullptr)
}
void rtg_lowered_free(rtg_lowered *l) { delete l; }
int rtg_lowered_desc(const rtg_lowered *l, rtgpu_scene_desc *out) {
    *out = l->l.desc();
    return 0;
}

int rtg_decode_jpeg(const char *path, int *w, int *h, uint8_t *rgb_out, size_t cap) {
    std::vector<uint8_t> rgb;
    std::string err;
    if (!rt::decode_jpeg(path, *w, *h, rgb, err)) {
        g_err = err;
        return -1;
    }
    if (rgb_out && cap >= rgb.size()) std::memcpy(rgb_out, rgb.data(), rgb.size());
    return 0;
}

int rtg_encode_jpeg(const char *path, const uint8_t *rgb, int w, int h, int quality) {
    std::string err;
    if (!rt::encode_jpeg(path, rgb, w, h, quality, err)) {
        g_err = err;
        return -1;
    }
    return 0;
}

} // extern "C"
