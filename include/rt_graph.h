/*
 * rt_graph.h -- POD description of a scene as the reference builds it (an object
 * graph), before it is lowered to flat GPU arrays.
 *
 * One entry per constructor call of the reference's scene-building API
 * (/root/reference/raytracer/src/scene.rs and the `new` functions cited per kind
 * below).  It is the neutral hand-over format between
 *   - the host scene API  (raytracer-2022_b200/host, C++ mirror of the Rust constructors),
 *   - the lowering        (graph -> rtgpu_scene_desc, include/rtgpu.h), and
 *   - the CPU oracle      (oracle/, test infrastructure, builds a pointer tree from it).
 * Everything is plain C: fixed-size structs, indices instead of pointers.
 */
#ifndef RT_GRAPH_H
#define RT_GRAPH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- textures: texture/mod.rs ------------------------------------------------ */
enum rt_texture_kind {
    RT_TEX_SOLID = 0,   /* SolidColor::new(color)          texture/mod.rs:20   p[0..2] = rgb           */
    RT_TEX_CHECKER = 1, /* CheckerTexture::new(c1,c2)      texture/mod.rs:43   a = odd tex, b = even   */
    RT_TEX_NOISE = 2,   /* NoiseTexture::new(scale)        texture/mod.rs:69   a = perlin id, p[0]=scale */
    RT_TEX_IMAGE = 3    /* ImageTexture::new(path)         texture/mod.rs:89   a = image id (-1: empty) */
};
typedef struct rt_texture {
    int32_t kind;
    int32_t a, b;
    int32_t pad_;
    double p[4];
} rt_texture;

/* ---- materials: material/mod.rs ---------------------------------------------- */
enum rt_material_kind {
    RT_MAT_LAMBERTIAN = 0,    /* material/mod.rs:37,43   tex = albedo                       */
    RT_MAT_METAL = 1,         /* material/mod.rs:76      p[0..2] = albedo, p[3] = min(fuzz,1) */
    RT_MAT_DIELECTRIC = 2,    /* material/mod.rs:106     p[0] = ir                          */
    RT_MAT_DIFFUSE_LIGHT = 3, /* material/mod.rs:160,166 tex = emit                         */
    RT_MAT_ISOTROPIC = 4      /* material/mod.rs:193,199 tex = albedo                       */
};
typedef struct rt_material {
    int32_t kind;
    int32_t tex;
    double p[4];
} rt_material;

/* ---- hittables: hittable/ ----------------------------------------------------- */
enum rt_object_kind {
    RT_OBJ_SPHERE = 0,        /* sphere.rs:22        p = cx cy cz r                               */
    RT_OBJ_MOVING_SPHERE = 1, /* sphere.rs:106       p = c0(3) c1(3) time0 time1 r                */
    RT_OBJ_XY_RECT = 2,       /* aarect.rs:27        p = x0 x1 y0 y1 k                            */
    RT_OBJ_XZ_RECT = 3,       /* aarect.rs:110       p = x0 x1 z0 z1 k                            */
    RT_OBJ_YZ_RECT = 4,       /* aarect.rs:193       p = y0 y1 z0 z1 k                            */
    RT_OBJ_BOX = 5,           /* boxes.rs:19         p = p0(3) p1(3); six rects, fixed order      */
    RT_OBJ_TRIANGLE = 6,      /* triangle.rs:22      p = a(3) b(3) c(3)                           */
    RT_OBJ_MEDIUM = 7,        /* constantmedium.rs:26,36  c0 = boundary, mat = Isotropic,
                                                      p[0] = density, p[1] = -1/density          */
    RT_OBJ_TRANSLATE = 8,     /* hittable/mod.rs:144 c0 = child, p[0..2] = offset                 */
    RT_OBJ_ROTATE_Y = 9,      /* hittable/mod.rs:188 c0 = child, p[0] = degrees, p[1] = sin,
                                                      p[2] = cos, p[3..8] = box, c1 = has box     */
    RT_OBJ_FLIP_FACE = 10,    /* hittable/mod.rs:275 c0 = child                                   */
    RT_OBJ_ZOOM = 11,         /* hittable/mod.rs:303 c0 = child, p[0] = rate                      */
    RT_OBJ_BVH_NODE = 12,     /* bvh/mod.rs:34       c0 = left, c1 = right, p[0..5] = box         */
    RT_OBJ_LIST = 13,         /* hittable/mod.rs:75  c0 = first index into list_items, c1 = count */
    RT_OBJ_RING = 14          /* ring.rs:22          p = r t; annulus in the y = 0 plane, radii r-t .. r+t         */
};
typedef struct rt_object {
    int32_t kind;
    int32_t mat; /* material id, -1 for containers */
    int32_t c0, c1;
    double p[12];
} rt_object;

/* Perlin::new()  texture/perlin.rs:16-34 */
typedef struct rt_perlin {
    double randvec[256][3];
    int32_t perm_x[256];
    int32_t perm_y[256];
    int32_t perm_z[256];
} rt_perlin;

/* ImageTexture: RGB8, rows stored bottom-up as texture/mod.rs:94-100 leaves them */
typedef struct rt_image {
    int32_t width, height;
    int64_t offset; /* byte offset into rt_graph.texels */
} rt_image;

/* Camera::new  basic/camera.rs:24-62: constructor arguments, then the derived fields
 * (computed on the host with the reference's operation order). */
typedef struct rt_camera {
    double lookfrom[3], lookat[3], vup[3];
    double vfov, aspect_ratio, aperture, focus_dist;
    double origin[3];
    double lower_left_corner[3];
    double horizontal[3];
    double vertical[3];
    double u[3], v[3], w[3];
    double lens_radius;
    double time0, time1;
} rt_camera;

typedef struct rt_graph {
    const rt_texture *textures;
    const rt_material *materials;
    const rt_object *objects;
    const int32_t *list_items;
    const rt_perlin *perlins;
    const rt_image *images;
    const uint8_t *texels;
    const int32_t *lights; /* object ids; the `lights` HittableList of main.rs:89 */
    int32_t n_textures, n_materials, n_objects, n_list_items;
    int32_t n_perlins, n_images, n_lights;
    int32_t world; /* object id the integrator calls hit() on (main.rs:90, :243) */
    int64_t n_texel_bytes;
    rt_camera camera;
    double background[3]; /* main.rs:51 */
} rt_graph;

#ifdef __cplusplus
}
#endif
#endif /* RT_GRAPH_H */
