/*
 * graph.cpp -- rt::Scene: the constructor mirror of the reference's scene API, bounding
 * boxes, BvhNode::new_vec with the reference's split rule, Camera::new and Perlin::new.
 * Compiled with -ffp-contract=off: derived values (boxes, camera frame, sin/cos) must be the
 * exact f64 results the reference's operation order gives.
 */
#include "rt_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace rt {

static const double PI = 3.14159265358979323846264338327950288;
static const double INF = std::numeric_limits<double>::infinity();

double Vec3::length() const { return std::sqrt(length_sqr()); }
Vec3 Vec3::to_unit() const { return *this / length(); }

Scene::Scene(uint64_t seed_) : seed(seed_) {
    rng.init(seed_, 0, 0, RT_DOMAIN_SCENE);
    std::memset(&camera, 0, sizeof camera);
    // a default camera so that a graph without set_camera is still well-formed
    set_camera(Vec3(0, 0, -1), Vec3(0, 0, 0), Vec3(0, 1, 0), 40., 1., 0., 1., 0., 1.);
}

Vec3 Scene::random_vec(double lo, double hi) { // basic/vec.rs:49-56
    double x = gen_range(lo, hi);
    double y = gen_range(lo, hi);
    double z = gen_range(lo, hi);
    return Vec3(x, y, z);
}

void Scene::check_obj(int32_t id) const {
    if (id < 0 || id >= (int32_t)objects.size()) throw Error{"object id out of range"};
}
void Scene::check_mat(int32_t id) const {
    if (id < 0 || id >= (int32_t)materials.size()) throw Error{"material id out of range"};
}
void Scene::check_tex(int32_t id) const {
    if (id < 0 || id >= (int32_t)textures.size()) throw Error{"texture id out of range"};
}

int32_t Scene::add_texture(int kind, int a, int b, const double *p, int np) {
    rt_texture t;
    std::memset(&t, 0, sizeof t);
    t.kind = kind;
    t.a = a;
    t.b = b;
    if (np > 4) throw Error{"texture: too many parameters"};
    for (int i = 0; i < np; ++i) t.p[i] = p[i];
    switch (kind) {
    case RT_TEX_SOLID:
        break;
    case RT_TEX_CHECKER:
        check_tex(a);
        check_tex(b);
        break;
    case RT_TEX_NOISE:
        if (a < 0 || a >= (int)perlins.size()) throw Error{"noise texture: perlin id out of range"};
        break;
    case RT_TEX_IMAGE:
        if (a >= (int)images.size()) throw Error{"image texture: image id out of range"};
        break;
    default:
        throw Error{"unknown texture kind"};
    }
    textures.push_back(t);
    return (int32_t)textures.size() - 1;
}

int32_t Scene::add_material(int kind, int tex, const double *p, int np) {
    rt_material m;
    std::memset(&m, 0, sizeof m);
    m.kind = kind;
    m.tex = tex;
    if (np > 4) throw Error{"material: too many parameters"};
    for (int i = 0; i < np; ++i) m.p[i] = p[i];
    switch (kind) {
    case RT_MAT_LAMBERTIAN:
    case RT_MAT_DIFFUSE_LIGHT:
    case RT_MAT_ISOTROPIC:
        check_tex(tex);
        break;
    case RT_MAT_METAL:
        m.tex = -1;
        m.p[3] = m.p[3] < 1. ? m.p[3] : 1.; // Metal::new, material/mod.rs:76-81
        break;
    case RT_MAT_DIELECTRIC:
        m.tex = -1;
        break;
    default:
        throw Error{"unknown material kind"};
    }
    materials.push_back(m);
    return (int32_t)materials.size() - 1;
}

int32_t Scene::add_object(int kind, int mat, int c0, int c1, const double *p, int np) {
    rt_object o;
    std::memset(&o, 0, sizeof o);
    o.kind = kind;
    o.mat = mat;
    o.c0 = c0;
    o.c1 = c1;
    if (np > 12) throw Error{"object: too many parameters"};
    for (int i = 0; i < np; ++i) o.p[i] = p[i];
    auto need = [&](int n) {
        if (np < n) throw Error{"object: too few parameters"};
    };
    switch (kind) {
    case RT_OBJ_SPHERE:
        need(4);
        check_mat(mat);
        o.c0 = o.c1 = -1;
        break;
    case RT_OBJ_MOVING_SPHERE:
        need(9);
        check_mat(mat);
        o.c0 = o.c1 = -1;
        break;
    case RT_OBJ_XY_RECT:
    case RT_OBJ_XZ_RECT:
    case RT_OBJ_YZ_RECT:
        need(5);
        check_mat(mat);
        o.c0 = o.c1 = -1;
        break;
    case RT_OBJ_BOX:
        need(6);
        check_mat(mat);
        o.c0 = o.c1 = -1;
        break;
    case RT_OBJ_TRIANGLE:
        need(9);
        check_mat(mat);
        o.c0 = o.c1 = -1;
        break;
    case RT_OBJ_RING:
        need(2);
        check_mat(mat);
        o.c0 = o.c1 = -1;
        break;
    case RT_OBJ_MEDIUM:
        need(1);
        check_obj(c0);
        check_mat(mat);
        if (materials[mat].kind != RT_MAT_ISOTROPIC) throw Error{"ConstantMedium: phase function must be Isotropic"};
        o.c1 = -1;
        o.p[1] = -1. / o.p[0]; // constantmedium.rs:29,39
        break;
    case RT_OBJ_TRANSLATE:
        need(3);
        check_obj(c0);
        o.mat = -1;
        o.c1 = -1;
        break;
    case RT_OBJ_ROTATE_Y: { // RotateY::new, hittable/mod.rs:188-232
        need(1);
        check_obj(c0);
        o.mat = -1;
        double radians = o.p[0] * (PI / 180.0); // f64::to_radians
        double sin_theta = std::sin(radians);
        double cos_theta = std::cos(radians);
        o.p[1] = sin_theta;
        o.p[2] = cos_theta;
        Box3 b;
        if (bounding_box(c0, 0., 1., b)) {
            double mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    for (int k = 0; k < 2; ++k) {
                        double x = (double)i * b.mx[0] + (double)(1 - i) * b.mn[0];
                        double y = (double)j * b.mx[1] + (double)(1 - j) * b.mn[1];
                        double z = (double)k * b.mx[2] + (double)(1 - k) * b.mn[2];
                        double newx = cos_theta * x + sin_theta * z;
                        double newz = -sin_theta * x + cos_theta * z;
                        double tester[3] = {newx, y, newz};
                        for (int c = 0; c < 3; ++c) {
                            mn[c] = std::fmin(mn[c], tester[c]);
                            mx[c] = std::fmax(mx[c], tester[c]);
                        }
                    }
            o.c1 = 1;
            for (int c = 0; c < 3; ++c) {
                o.p[3 + c] = mn[c];
                o.p[6 + c] = mx[c];
            }
        } else {
            o.c1 = 0;
        }
        break;
    }
    case RT_OBJ_FLIP_FACE:
        check_obj(c0);
        o.mat = -1;
        o.c1 = -1;
        break;
    case RT_OBJ_ZOOM:
        need(1);
        check_obj(c0);
        o.mat = -1;
        o.c1 = -1;
        break;
    case RT_OBJ_BVH_NODE:
        need(6);
        check_obj(c0);
        check_obj(c1);
        o.mat = -1;
        break;
    case RT_OBJ_LIST:
        if (c0 < 0 || c1 < 0 || (size_t)c0 + (size_t)c1 > list_items.size()) throw Error{"list: bad item range"};
        o.mat = -1;
        break;
    default:
        throw Error{"unknown object kind"};
    }
    objects.push_back(o);
    return (int32_t)objects.size() - 1;
}

int32_t Scene::add_list(const int32_t *ids, int n) {
    int32_t first = (int32_t)list_items.size();
    for (int i = 0; i < n; ++i) {
        check_obj(ids[i]);
        list_items.push_back(ids[i]);
    }
    return add_object(RT_OBJ_LIST, -1, first, n, nullptr, 0);
}

int32_t Scene::add_perlin() { // Perlin::new, texture/perlin.rs:16-49
    rt_perlin t;
    for (int i = 0; i < 256; ++i) {
        Vec3 v = random_vec(-1., 1.).to_unit(); // not rejection-sampled upstream either (perlin.rs:20)
        t.randvec[i][0] = v.x;
        t.randvec[i][1] = v.y;
        t.randvec[i][2] = v.z;
    }
    int32_t *perms[3] = {t.perm_x, t.perm_y, t.perm_z};
    for (int a = 0; a < 3; ++a) {
        int32_t *p = perms[a];
        for (int i = 0; i < 256; ++i) p[i] = i;
        for (int i = 255; i >= 0; --i) { // perlin.rs:41-48
            uint32_t target = rng.gen_index((uint32_t)i + 1);
            std::swap(p[i], p[target]);
        }
    }
    perlins.push_back(t);
    return (int32_t)perlins.size() - 1;
}

int32_t Scene::add_image_rgb8(int w, int h, const uint8_t *top_down_rgb) {
    if (w <= 0 || h <= 0) throw Error{"image: empty"};
    rt_image im;
    im.width = w;
    im.height = h;
    im.offset = (int64_t)texels.size();
    // texture/mod.rs:94-100: stored row y = source row (height - y - 1)
    for (int y = 0; y < h; ++y) {
        const uint8_t *src = top_down_rgb + (size_t)(h - y - 1) * (size_t)w * 3;
        texels.insert(texels.end(), src, src + (size_t)w * 3);
    }
    while (texels.size() % 16) texels.push_back(0);
    images.push_back(im);
    return (int32_t)images.size() - 1;
}

int32_t Scene::add_image_file(const std::string &path) {
    int w = 0, h = 0;
    std::vector<uint8_t> rgb;
    std::string err;
    if (!decode_jpeg(path, w, h, rgb, err)) throw Error{"ImageTexture::new(\"" + path + "\"): " + err}; // unwrap()
    return add_image_rgb8(w, h, rgb.data());
}

bool Scene::bounding_box(int32_t id, double time0, double time1, Box3 &out) const {
    check_obj(id);
    const rt_object &o = objects[id];
    const double *p = o.p;
    auto set = [&](double x0, double y0, double z0, double x1, double y1, double z1) {
        out.mn[0] = x0; out.mn[1] = y0; out.mn[2] = z0;
        out.mx[0] = x1; out.mx[1] = y1; out.mx[2] = z1;
    };
    auto surround = [](const Box3 &a, const Box3 &b) { // aabb.rs:34-46
        Box3 r;
        for (int c = 0; c < 3; ++c) {
            r.mn[c] = std::fmin(a.mn[c], b.mn[c]);
            r.mx[c] = std::fmax(a.mx[c], b.mx[c]);
        }
        return r;
    };
    switch (o.kind) {
    case RT_OBJ_SPHERE: // sphere.rs:68-73
        set(p[0] - p[3], p[1] - p[3], p[2] - p[3], p[0] + p[3], p[1] + p[3], p[2] + p[3]);
        return true;
    case RT_OBJ_MOVING_SPHERE: { // sphere.rs:124-127,167-177
        double r = p[8];
        Box3 b[2];
        double tt[2] = {time0, time1};
        for (int k = 0; k < 2; ++k) {
            double f = (tt[k] - p[6]) / (p[7] - p[6]);
            for (int c = 0; c < 3; ++c) {
                double ctr = p[c] + (p[3 + c] - p[c]) * f;
                b[k].mn[c] = ctr - r;
                b[k].mx[c] = ctr + r;
            }
        }
        out = surround(b[0], b[1]);
        return true;
    }
    case RT_OBJ_XY_RECT: // aarect.rs:40-45
        set(p[0], p[2], p[4] - 0.0001, p[1], p[3], p[4] + 0.0001);
        return true;
    case RT_OBJ_XZ_RECT: // aarect.rs:123-128
        set(p[0], p[4] - 0.0001, p[2], p[1], p[4] + 0.0001, p[3]);
        return true;
    case RT_OBJ_YZ_RECT: // aarect.rs:206-211
        set(p[4] - 0.0001, p[0], p[2], p[4] + 0.0001, p[1], p[3]);
        return true;
    case RT_OBJ_BOX: // boxes.rs:79-81
        set(p[0], p[1], p[2], p[3], p[4], p[5]);
        return true;
    case RT_OBJ_TRIANGLE: // triangle.rs:79-92
        for (int c = 0; c < 3; ++c) {
            out.mn[c] = std::fmin(p[c], std::fmin(p[3 + c], p[6 + c]));
            out.mx[c] = std::fmax(p[c], std::fmax(p[3 + c], p[6 + c]));
        }
        return true;
    case RT_OBJ_RING: { // ring.rs:54-62
        double rr = p[0] + p[1];
        set(-rr, -0.0001, -rr, rr, 0.0001, rr);
        return true;
    }
    case RT_OBJ_MEDIUM: // constantmedium.rs:46-48
    case RT_OBJ_FLIP_FACE: // hittable/mod.rs:289-291
        return bounding_box(o.c0, time0, time1, out);
    case RT_OBJ_TRANSLATE: { // hittable/mod.rs:154-163
        Box3 b;
        if (!bounding_box(o.c0, time0, time1, b)) return false;
        for (int c = 0; c < 3; ++c) {
            out.mn[c] = b.mn[c] + p[c];
            out.mx[c] = b.mx[c] + p[c];
        }
        return true;
    }
    case RT_OBJ_ROTATE_Y: // hittable/mod.rs:236-238: the box cached by the constructor
        if (!o.c1) return false;
        set(p[3], p[4], p[5], p[6], p[7], p[8]);
        return true;
    case RT_OBJ_ZOOM: { // hittable/mod.rs:310-319
        Box3 b;
        if (!bounding_box(o.c0, time0, time1, b)) return false;
        for (int c = 0; c < 3; ++c) {
            out.mn[c] = b.mn[c] * p[0];
            out.mx[c] = b.mx[c] * p[0];
        }
        return true;
    }
    case RT_OBJ_BVH_NODE: // bvh/mod.rs:102-104
        set(p[0], p[1], p[2], p[3], p[4], p[5]);
        return true;
    case RT_OBJ_LIST: { // hittable/mod.rs:102-120
        if (o.c1 == 0) return false;
        Box3 acc;
        set(INF, INF, INF, -INF, -INF, -INF);
        acc = out;
        for (int i = 0; i < o.c1; ++i) {
            Box3 b;
            if (!bounding_box(list_items[o.c0 + i], time0, time1, b)) return false;
            acc = surround(acc, b);
        }
        out = acc;
        return true;
    }
    }
    throw Error{"bounding_box: unknown object kind"};
}

int32_t Scene::bvh_new_vec(std::vector<int32_t> ids, double time0, double time1) { // bvh/mod.rs:34-81
    int axis = (int)rng.gen_index(3); // drawn first, for every node, also for 1- and 2-object nodes (:35)
    size_t span = ids.size();
    int32_t left, right;
    if (span == 0) throw Error{"BvhNode::new_vec: Get empty vec"}; // :41-43
    auto min_at_0 = [&](int32_t id) {                              // box_compare, :19-29
        Box3 b;
        if (!bounding_box(id, 0., 0., b)) throw Error{"BvhNode::box_compare: no bounding box"}; // unwrap()
        return b.mn[axis];
    };
    if (span == 1) { // :44-47
        left = right = ids[0];
    } else if (span == 2) { // :48-57: obj0 = last, obj1 = first
        int32_t obj0 = ids[1], obj1 = ids[0];
        if (min_at_0(obj0) < min_at_0(obj1)) {
            left = obj0;
            right = obj1;
        } else {
            left = obj1;
            right = obj0;
        }
    } else { // :58-66
        std::vector<std::pair<double, int32_t>> keyed;
        keyed.reserve(span);
        for (int32_t id : ids) keyed.emplace_back(min_at_0(id), id);
        std::stable_sort(keyed.begin(), keyed.end(),
                         [](const std::pair<double, int32_t> &a, const std::pair<double, int32_t> &b) {
                             return a.first < b.first;
                         });
        std::vector<int32_t> l, r;
        for (size_t i = 0; i < span; ++i) (i < span / 2 ? l : r).push_back(keyed[i].second);
        left = bvh_new_vec(l, time0, time1);
        right = bvh_new_vec(r, time0, time1);
    }
    Box3 lb, rb;
    if (!bounding_box(left, time0, time1, lb) || !bounding_box(right, time0, time1, rb))
        throw Error{"BvhNode::new_vec: No bounding box in bvh_node constructor."}; // :76,79
    double box[6];
    for (int c = 0; c < 3; ++c) {
        box[c] = std::fmin(lb.mn[c], rb.mn[c]);
        box[3 + c] = std::fmax(lb.mx[c], rb.mx[c]);
    }
    return add_object(RT_OBJ_BVH_NODE, -1, left, right, box, 6);
}

void Scene::set_camera(Vec3 lookfrom, Vec3 lookat, Vec3 vup, double vfov, double aspect_ratio, double aperture,
                       double focus_dist, double time0, double time1) { // camera.rs:24-62
    double theta = vfov * (PI / 180.0);
    double h = std::tan(theta / 2.);
    double viewport_height = 2. * h;
    double viewport_width = aspect_ratio * viewport_height;
    Vec3 w = (lookfrom - lookat).to_unit();
    Vec3 u = cross(vup, w).to_unit();
    Vec3 v = cross(w, u);
    Vec3 origin = lookfrom;
    Vec3 horizontal = u * viewport_width * focus_dist;
    Vec3 vertical = v * viewport_height * focus_dist;
    Vec3 llc = origin - horizontal / 2. - vertical / 2. - w * focus_dist;
    auto put = [](double *d, Vec3 a) { d[0] = a.x; d[1] = a.y; d[2] = a.z; };
    put(camera.lookfrom, lookfrom);
    put(camera.lookat, lookat);
    put(camera.vup, vup);
    camera.vfov = vfov;
    camera.aspect_ratio = aspect_ratio;
    camera.aperture = aperture;
    camera.focus_dist = focus_dist;
    put(camera.origin, origin);
    put(camera.lower_left_corner, llc);
    put(camera.horizontal, horizontal);
    put(camera.vertical, vertical);
    put(camera.u, u);
    put(camera.v, v);
    put(camera.w, w);
    camera.lens_radius = aperture / 2.;
    camera.time0 = time0;
    camera.time1 = time1;
}

rt_graph Scene::view() const {
    rt_graph g;
    std::memset(&g, 0, sizeof g);
    g.textures = textures.data();
    g.materials = materials.data();
    g.objects = objects.data();
    g.list_items = list_items.data();
    g.perlins = perlins.data();
    g.images = images.data();
    g.texels = texels.data();
    g.lights = lights.data();
    g.n_textures = (int32_t)textures.size();
    g.n_materials = (int32_t)materials.size();
    g.n_objects = (int32_t)objects.size();
    g.n_list_items = (int32_t)list_items.size();
    g.n_perlins = (int32_t)perlins.size();
    g.n_images = (int32_t)images.size();
    g.n_lights = (int32_t)lights.size();
    g.world = world;
    g.n_texel_bytes = (int64_t)texels.size();
    g.camera = camera;
    for (int c = 0; c < 3; ++c) g.background[c] = background[c];
    return g;
}

/* ---- constructor mirror --------------------------------------------------------------------- */
Tex SolidColor::new_(Scene &s, Color c) {
    double p[3] = {c.x, c.y, c.z};
    return Tex{s.add_texture(RT_TEX_SOLID, -1, -1, p, 3)};
}
Tex CheckerTexture::new_(Scene &s, Color c1, Color c2) { // odd = c1, even = c2 (texture/mod.rs:43-48)
    Tex odd = SolidColor::new_(s, c1), even = SolidColor::new_(s, c2);
    return Tex{s.add_texture(RT_TEX_CHECKER, odd.id, even.id, nullptr, 0)};
}
Tex NoiseTexture::new_(Scene &s, double scale) {
    int32_t pn = s.add_perlin();
    double p[1] = {scale};
    return Tex{s.add_texture(RT_TEX_NOISE, pn, -1, p, 1)};
}
Tex ImageTexture::new_(Scene &s, const std::string &filename) {
    int32_t im = s.add_image_file(filename);
    return Tex{s.add_texture(RT_TEX_IMAGE, im, -1, nullptr, 0)};
}

Mat Lambertian::new_(Scene &s, Color a) { return new_arc(s, SolidColor::new_(s, a)); }
Mat Lambertian::new_arc(Scene &s, Tex albedo) { return Mat{s.add_material(RT_MAT_LAMBERTIAN, albedo.id, nullptr, 0)}; }
Mat Metal::new_(Scene &s, Color a, double f) {
    double p[4] = {a.x, a.y, a.z, f};
    return Mat{s.add_material(RT_MAT_METAL, -1, p, 4)};
}
Mat Dielectric::new_(Scene &s, double ir) {
    double p[1] = {ir};
    return Mat{s.add_material(RT_MAT_DIELECTRIC, -1, p, 1)};
}
Mat DiffuseLight::new_(Scene &s, Color c) { return new_arc(s, SolidColor::new_(s, c)); }
Mat DiffuseLight::new_arc(Scene &s, Tex emit) { return Mat{s.add_material(RT_MAT_DIFFUSE_LIGHT, emit.id, nullptr, 0)}; }
Mat Isotropic::new_(Scene &s, Color c) { return new_arc(s, SolidColor::new_(s, c)); }
Mat Isotropic::new_arc(Scene &s, Tex a) { return Mat{s.add_material(RT_MAT_ISOTROPIC, a.id, nullptr, 0)}; }

Obj Sphere::new_(Scene &s, Point3 c, double r, Mat m) {
    double p[4] = {c.x, c.y, c.z, r};
    return Obj{s.add_object(RT_OBJ_SPHERE, m.id, -1, -1, p, 4)};
}
Obj MovingSphere::new_(Scene &s, Point3 c0, Point3 c1, double t0, double t1, double r, Mat m) {
    double p[9] = {c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, t0, t1, r};
    return Obj{s.add_object(RT_OBJ_MOVING_SPHERE, m.id, -1, -1, p, 9)};
}
Obj XYRect::new_(Scene &s, double x0, double x1, double y0, double y1, double k, Mat m) {
    double p[5] = {x0, x1, y0, y1, k};
    return Obj{s.add_object(RT_OBJ_XY_RECT, m.id, -1, -1, p, 5)};
}
Obj XZRect::new_(Scene &s, double x0, double x1, double z0, double z1, double k, Mat m) {
    double p[5] = {x0, x1, z0, z1, k};
    return Obj{s.add_object(RT_OBJ_XZ_RECT, m.id, -1, -1, p, 5)};
}
Obj YZRect::new_(Scene &s, double y0, double y1, double z0, double z1, double k, Mat m) {
    double p[5] = {y0, y1, z0, z1, k};
    return Obj{s.add_object(RT_OBJ_YZ_RECT, m.id, -1, -1, p, 5)};
}
Obj Boxes::new_(Scene &s, Point3 p0, Point3 p1, Mat m) {
    double p[6] = {p0.x, p0.y, p0.z, p1.x, p1.y, p1.z};
    return Obj{s.add_object(RT_OBJ_BOX, m.id, -1, -1, p, 6)};
}
Obj Triangle::new_(Scene &s, Point3 a, Point3 b, Point3 c, Mat m) {
    double p[9] = {a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z};
    return Obj{s.add_object(RT_OBJ_TRIANGLE, m.id, -1, -1, p, 9)};
}
Obj Ring::new_(Scene &s, double r, double t, Mat m) {
    double p[2] = {r, t};
    return Obj{s.add_object(RT_OBJ_RING, m.id, -1, -1, p, 2)};
}
Obj ConstantMedium::new_(Scene &s, Obj boundary, double d, Color c) { return new_arc(s, boundary, d, SolidColor::new_(s, c)); }
Obj ConstantMedium::new_arc(Scene &s, Obj boundary, double d, Tex a) {
    Mat iso = Isotropic::new_arc(s, a);
    double p[1] = {d};
    return Obj{s.add_object(RT_OBJ_MEDIUM, iso.id, boundary.id, -1, p, 1)};
}
Obj Translate::new_(Scene &s, Obj c, Vec3 d) {
    double p[3] = {d.x, d.y, d.z};
    return Obj{s.add_object(RT_OBJ_TRANSLATE, -1, c.id, -1, p, 3)};
}
Obj RotateY::new_(Scene &s, Obj c, double angle) {
    double p[1] = {angle};
    return Obj{s.add_object(RT_OBJ_ROTATE_Y, -1, c.id, -1, p, 1)};
}
Obj FlipFace::new_(Scene &s, Obj c) { return Obj{s.add_object(RT_OBJ_FLIP_FACE, -1, c.id, -1, nullptr, 0)}; }
Obj Zoom::new_(Scene &s, Obj c, double rate) {
    double p[1] = {rate};
    return Obj{s.add_object(RT_OBJ_ZOOM, -1, c.id, -1, p, 1)};
}
Obj HittableList::as_object(Scene &s) const {
    std::vector<int32_t> ids;
    for (Obj o : objects) ids.push_back(o.id);
    return Obj{s.add_list(ids.data(), (int)ids.size())};
}
Obj BvhNode::new_list(Scene &s, const HittableList &list, double t0, double t1) { return new_vec(s, list.objects, t0, t1); }
Obj BvhNode::new_vec(Scene &s, const std::vector<Obj> &objects, double t0, double t1) {
    std::vector<int32_t> ids;
    for (Obj o : objects) ids.push_back(o.id);
    return Obj{s.bvh_new_vec(ids, t0, t1)};
}

} // namespace rt
