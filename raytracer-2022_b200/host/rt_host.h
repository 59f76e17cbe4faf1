/*
 * rt_host.h -- host side of the B200 path tracer: a C++ mirror of the reference's
 * scene-building API (same constructor names, argument order and error behaviour as
 * /root/reference/raytracer/src), the BVH build with the reference's split rule, OBJ and
 * JPEG loading, and the lowering to the flat arrays of include/rtgpu.h.
 *
 * No Rust toolchain exists in this image, so the host that north_star describes in Rust is
 * written in C++ above the same C ABI.  `new` is a C++ keyword: constructors are spelled
 * `new_` / `new_arc` / `new_list` / `new_vec`.
 *
 * Every constructor appends one entry to a rt::Scene (the object graph of
 * include/rt_graph.h) and returns a typed handle; handles replace the reference's
 * by-value generics (`Sphere<M>`) and `Arc<dyn Hittable>`.
 */
#ifndef RT_HOST_H
#define RT_HOST_H

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/rt_graph.h"
#include "../../include/rt_rng.h"
#include "../../include/rtgpu.h"

namespace rt {

struct Vec3 { // basic/vec.rs:12-17
    double x, y, z;
    Vec3() : x(0), y(0), z(0) {}
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    static Vec3 new_(double a, double b, double c) { return Vec3(a, b, c); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double length_sqr() const { return x * x + y * y + z * z; }
    double length() const;
    Vec3 to_unit() const;
};
typedef Vec3 Point3;
typedef Vec3 Color;
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator/(Vec3 a, double s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

struct Tex { int32_t id; };
struct Mat { int32_t id; };
struct Obj { int32_t id; };

struct Box3 { // hittable/bvh/aabb.rs:4-8
    double mn[3], mx[3];
};

/* The reference panics (unwrap/assert!/panic!) on bad input; here every failure throws
 * rt::Error, which the C entry points turn into an error code + message. */
struct Error {
    std::string what;
};

class Scene {
  public:
    explicit Scene(uint64_t seed);

    /* `rand::thread_rng()` of the scene-construction code: one sequential stream */
    rt_rng rng;
    uint64_t seed;
    double gen() { return rng.gen(); }
    double gen_range(double lo, double hi) { return rng.gen_range(lo, hi); }
    Vec3 random_vec(double lo, double hi); // Vec3::random, basic/vec.rs:49-56

    std::vector<rt_texture> textures;
    std::vector<rt_material> materials;
    std::vector<rt_object> objects;
    std::vector<int32_t> list_items;
    std::vector<rt_perlin> perlins;
    std::vector<rt_image> images;
    std::vector<uint8_t> texels;
    std::vector<int32_t> lights;
    int32_t world = -1;
    rt_camera camera;
    double background[3] = {0, 0, 0};

    int32_t add_texture(int kind, int a, int b, const double *p, int np);
    int32_t add_material(int kind, int tex, const double *p, int np);
    /* generic object constructor: fills the derived fields of ROTATE_Y and MEDIUM */
    int32_t add_object(int kind, int mat, int c0, int c1, const double *p, int np);
    int32_t add_list(const int32_t *ids, int n);
    int32_t add_perlin();                                              // Perlin::new, perlin.rs:16-34
    int32_t add_image_rgb8(int w, int h, const uint8_t *top_down_rgb); // rows get flipped (mod.rs:94-100)
    int32_t add_image_file(const std::string &path);                   // image::open(...).unwrap()

    /* Hittable::bounding_box for a graph object; false = None */
    bool bounding_box(int32_t obj, double time0, double time1, Box3 &out) const;
    /* BvhNode::new_vec, hittable/bvh/mod.rs:34-81 */
    int32_t bvh_new_vec(std::vector<int32_t> ids, double time0, double time1);

    void set_camera(Vec3 lookfrom, Vec3 lookat, Vec3 vup, double vfov, double aspect_ratio, double aperture,
                    double focus_dist, double time0, double time1); // Camera::new, camera.rs:24-62

    rt_graph view() const; // pointers stay valid until the next mutation
    void check_obj(int32_t id) const;
    void check_mat(int32_t id) const;
    void check_tex(int32_t id) const;
};

/* ---- texture/mod.rs ---------------------------------------------------------------------- */
struct SolidColor { static Tex new_(Scene &s, Color c); };                          // :20
struct CheckerTexture { static Tex new_(Scene &s, Color c1, Color c2); };           // :43
struct NoiseTexture { static Tex new_(Scene &s, double scale); };                   // :69
struct ImageTexture { static Tex new_(Scene &s, const std::string &filename); };    // :89

/* ---- material/mod.rs --------------------------------------------------------------------- */
struct Lambertian {
    static Mat new_(Scene &s, Color a);      // :43
    static Mat new_arc(Scene &s, Tex albedo); // :37
};
struct Metal { static Mat new_(Scene &s, Color a, double f); };         // :76
struct Dielectric { static Mat new_(Scene &s, double index_of_refraction); }; // :106
struct DiffuseLight {
    static Mat new_(Scene &s, Color c);     // :166
    static Mat new_arc(Scene &s, Tex emit); // :160
};
struct Isotropic {
    static Mat new_(Scene &s, Color c);   // :199
    static Mat new_arc(Scene &s, Tex a);  // :193
};

/* ---- hittable/ ---------------------------------------------------------------------------- */
struct Sphere { static Obj new_(Scene &s, Point3 center, double radius, Mat m); }; // sphere.rs:22
struct MovingSphere {                                                              // sphere.rs:106
    static Obj new_(Scene &s, Point3 center0, Point3 center1, double time0, double time1, double radius, Mat m);
};
struct XYRect { static Obj new_(Scene &s, double x0, double x1, double y0, double y1, double k, Mat m); }; // aarect.rs:27
struct XZRect { static Obj new_(Scene &s, double x0, double x1, double z0, double z1, double k, Mat m); }; // aarect.rs:110
struct YZRect { static Obj new_(Scene &s, double y0, double y1, double z0, double z1, double k, Mat m); }; // aarect.rs:193
struct Boxes { static Obj new_(Scene &s, Point3 p0, Point3 p1, Mat m); };                                // boxes.rs:19
struct Triangle { static Obj new_(Scene &s, Point3 a, Point3 b, Point3 c, Mat m); };                     // triangle.rs:22
struct Ring { static Obj new_(Scene &s, double r, double t, Mat m); };                                    // ring.rs:22
struct ConstantMedium {
    static Obj new_(Scene &s, Obj boundary, double d, Color c);    // constantmedium.rs:36
    static Obj new_arc(Scene &s, Obj boundary, double d, Tex a);   // constantmedium.rs:26
};
struct Translate { static Obj new_(Scene &s, Obj p, Vec3 displacement); }; // hittable/mod.rs:144
struct RotateY { static Obj new_(Scene &s, Obj p, double angle); };        // hittable/mod.rs:188
struct FlipFace { static Obj new_(Scene &s, Obj p); };                     // hittable/mod.rs:275
struct Zoom { static Obj new_(Scene &s, Obj p, double rate); };            // hittable/mod.rs:303

struct HittableList { // hittable/mod.rs:70-87
    std::vector<Obj> objects;
    void add(Obj o) { objects.push_back(o); }
    Obj as_object(Scene &s) const; // a list used as a Hittable (e.g. a medium boundary)
};
struct BvhNode {
    static Obj new_list(Scene &s, const HittableList &list, double time0, double time1); // bvh/mod.rs:30
    static Obj new_vec(Scene &s, const std::vector<Obj> &objects, double time0, double time1); // bvh/mod.rs:34
};

/* tobj::load_obj(path, {single_index, triangulate}) as used at scene.rs:368-375: positions are
 * parsed as f32 and widened (scene.rs:386-388); one model per `o`/`g` group that has faces. */
struct ObjMesh {
    std::string name;
    std::vector<float> positions;  // xyz per vertex
    std::vector<uint32_t> indices; // 3 per triangle
};
std::vector<ObjMesh> load_obj(const std::string &path);

/* Baseline JPEG -> RGB8, top row first (stands in for image 0.23.14 / jpeg-decoder 0.1.22) */
bool decode_jpeg(const std::string &path, int &w, int &h, std::vector<uint8_t> &rgb, std::string &err);
bool decode_jpeg_mem(const uint8_t *data, size_t n, int &w, int &h, std::vector<uint8_t> &rgb, std::string &err);
/* RGB8 (top row first) -> baseline JPEG, 4:4:4, IJG quality scale; stands in for image 0.23.14's write_to(Jpeg(q)), main.rs:213-221 */
bool encode_jpeg_mem(const uint8_t *rgb, int w, int h, int quality, std::vector<uint8_t> &out, std::string &err);
bool encode_jpeg(const std::string &path, const uint8_t *rgb, int w, int h, int quality, std::string &err);

/* ---- scene.rs ------------------------------------------------------------------------------ */
struct SceneOptions {
    std::string asset_dir; // where earthmap.jpg / obj/Shuttle.obj live ("source/" upstream)
};
/* world list + lights list as the scene functions return them */
struct SceneLists {
    HittableList world, lights;
};
SceneLists random_scene(Scene &s);                            // scene.rs:22-84
SceneLists cornell_box(Scene &s);                             // scene.rs:165-196
SceneLists cornell_smoke(Scene &s);                           // scene.rs:199-257
SceneLists final_scene(Scene &s, const SceneOptions &o);      // scene.rs:260-362
SceneLists wwscene(Scene &s, const SceneOptions &o);          // scene.rs:468-571 (without get_ship: Ship.obj is missing upstream)
void get_shuttle(Scene &s, HittableList &world, const SceneOptions &o, double zoom, double angle, Vec3 offset); // scene.rs:364-414
/* the five BASELINE.json configs: scene function + camera + background + BvhNode::new_list
 * over the world list (main.rs:68-90).  name: "random_scene" | "cornell_box" | "final_scene" |
 * "obj_cornell" | "cornell_smoke" | "cornell_book3" | "wwscene". */
void build_config(Scene &s, const std::string &name, double aspect_ratio, const SceneOptions &o);

/* ---- lowering: graph -> flat arrays of include/rtgpu.h -------------------------------------- */
struct Lowered {
    std::vector<rtgpu_op> ops;
    std::vector<rtgpu_xform> xforms;
    std::vector<uint32_t> chain;
    std::vector<rtgpu_ctx> ctxs;
    std::vector<rtgpu_material> materials;
    std::vector<rtgpu_texture> textures;
    std::vector<rtgpu_perlin> perlins;
    std::vector<rtgpu_image> images;
    std::vector<uint8_t> texels;
    std::vector<rtgpu_light> lights;
    uint32_t n_light_roots = 0; /* members of `lights` itself; members of nested lists follow */
    uint32_t max_push_depth = 0;
    rtgpu_camera camera;
    double background[3];
    rtgpu_scene_desc desc() const;
};
/* elide_singletons: skip the second visit of a `left == right` BvhNode when the subtree has no
 * ConstantMedium */
Lowered lower(const rt_graph &g, bool elide_singletons = true);

} // namespace rt

/* ---- C entry points over the above (Python tests/bench bind these with ctypes) -------------- */
extern "C" {
typedef struct rtg_scene rtg_scene;
typedef struct rtg_lowered rtg_lowered;

const char *rtg_last_error(void);
rtg_scene *rtg_new(uint64_t seed);
void rtg_free(rtg_scene *s);
int32_t rtg_add_texture(rtg_scene *s, int kind, int a, int b, const double *p, int np);
int32_t rtg_add_noise_texture(rtg_scene *s, double scale);
int32_t rtg_add_image_texture_file(rtg_scene *s, const char *path);
int32_t rtg_add_image_texture_rgb8(rtg_scene *s, int w, int h, const uint8_t *top_down_rgb);
int32_t rtg_add_material(rtg_scene *s, int kind, int tex, const double *p, int np);
int32_t rtg_add_object(rtg_scene *s, int kind, int mat, int c0, int c1, const double *p, int np);
int32_t rtg_add_list(rtg_scene *s, const int32_t *ids, int n);
int32_t rtg_bvh(rtg_scene *s, const int32_t *ids, int n, double time0, double time1);
int32_t rtg_load_obj_triangles(rtg_scene *s, const char *path, int mat, int32_t *first_id, int32_t *count);
int rtg_bounding_box(rtg_scene *s, int32_t obj, double time0, double time1, double out6[6]);
int rtg_set_world(rtg_scene *s, int32_t obj);
int rtg_set_lights(rtg_scene *s, const int32_t *ids, int n);
int rtg_set_camera(rtg_scene *s, const double lookfrom[3], const double lookat[3], const double vup[3],
                   double vfov, double aspect_ratio, double aperture, double focus_dist, double time0,
                   double time1);
int rtg_set_background(rtg_scene *s, const double rgb[3]);
int rtg_build_config(rtg_scene *s, const char *name, double aspect_ratio, const char *asset_dir);
uint32_t rtg_rng_pos(const rtg_scene *s);
int rtg_graph(const rtg_scene *s, rt_graph *out);

rtg_lowered *rtg_lower(const rt_graph *g, int elide_singletons);
void rtg_lowered_free(rtg_lowered *l);
int rtg_lowered_desc(const rtg_lowered *l, rtgpu_scene_desc *out);

int rtg_decode_jpeg(const char *path, int *w, int *h, uint8_t *rgb_out, size_t cap);
int rtg_encode_jpeg(const char *path, const uint8_t *rgb, int w, int h, int quality);
}

#endif /* RT_HOST_H */
