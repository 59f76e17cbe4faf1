/*
 * rtgpu.h -- C ABI of the B200 path-tracing core (librtgpu.so).
 *
 * The reference (Jerx2y/Raytracer-2022) has no FFI, plugin or operator interface: its hot
 * path is called in exactly one place, raytracer/src/main.rs:150
 *     pixel_color += ray_color(r, background, &world, &lights, MAX_DEPTH);
 * inside the pixel/sample loops of main.rs:140-153, with the camera built at main.rs:68-78
 * and the scene + BVH at main.rs:89-90.  This header is the seam a Rust `extern "C"` crate
 * would bind in place of those loops (binding stub: INTEGRATION.md):
 *
 *   rtgpu_scene_create   replaces  scene::*() + BvhNode::new_list          main.rs:89-90
 *                        (the host has already lowered the object graph to the flat arrays
 *                         below: raytracer-2022_b200/host/lower.cpp)
 *   rtgpu_render         replaces  the thread pool + pixel/sample loops     main.rs:109-183
 *                                  + ray_color                              main.rs:233-278
 *   rtgpu_trace_batch    replaces  BvhNode::hit for a dumped ray batch      hittable/bvh/mod.rs:86-101
 *   rtgpu_resolve        replaces  write_color + the row flip               main.rs:191-201,280-299
 *
 * Plain pointers and sizes only; no C++ or torch types.  All functions return rtgpu_status
 * and never throw or abort; rtgpu_last_error() gives the message of the last failure on
 * the calling thread.  There is no CPU fallback: without a CUDA device every entry point
 * that needs one returns RTGPU_ERR_CUDA.
 */
#ifndef RTGPU_H
#define RTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rtgpu_status {
    RTGPU_OK = 0,
    RTGPU_ERR_INVALID_SCENE = 1, /* the reference would panic: empty BVH, missing box, bad index */
    RTGPU_ERR_INVALID_ARG = 2,
    RTGPU_ERR_CUDA = 3,
    RTGPU_ERR_OOM = 4,
    RTGPU_ERR_NCCL = 5
} rtgpu_status;

/* ------------------------------------------------------------------------------------------
 * The traversal program: the whole object graph flattened in the reference's fixed
 * left-then-right visiting order (hittable/bvh/mod.rs:86-101, hittable/mod.rs:90-100), with
 * a skip index per BvhNode.  One op = 64 bytes = 16-byte header + 6 doubles, 64-byte aligned,
 * so a warp fetches it with four 128-bit loads.  Two-slot ops carry 8 more doubles in the
 * next slot.
 * ---------------------------------------------------------------------------------------- */
enum rtgpu_op_kind {
    RTGPU_OP_NODE = 0,     /* AABB: d = min(3) max(3); miss -> jump to `skip`              */
    RTGPU_OP_SPHERE = 1,   /* d = center(3) radius                                         */
    RTGPU_OP_MSPHERE = 2,  /* 2 slots: d = center0(3) radius time0 time1 | center1(3)      */
    RTGPU_OP_RECT_XY = 3,  /* d = a0 a1 b0 b1 k  (aarect.rs argument order)                */
    RTGPU_OP_RECT_XZ = 4,
    RTGPU_OP_RECT_YZ = 5,
    RTGPU_OP_BOX = 6,      /* d = p0(3) p1(3); six rects in the order of boxes.rs:24-69    */
    RTGPU_OP_TRI = 7,      /* 2 slots: d = a(3) b(3) | c(3) n(3); n = unit normal,
                              precomputed with triangle.rs:30-32's operation order         */
    RTGPU_OP_MEDIUM = 8,   /* d[0] = -1/density; boundary = ops [i+1, skip)                */
                           /* (a boundary may itself hold ONE more level of MEDIUM ops, not two)  */
    RTGPU_OP_PUSH = 9,     /* enter wrappers xforms[chain[a .. a+b)] (outermost first)     */
    RTGPU_OP_POP = 10,     /* leave them: restore the ray saved by the matching PUSH       */
    RTGPU_OP_RING = 11     /* d = dis_min dis_max: (r-t)^2, (r+t)^2, annulus in the plane y = 0 (ring.rs)  */
};

typedef struct rtgpu_op {
    uint32_t kind_ctx; /* kind | (context id << 8) */
    uint32_t skip;     /* NODE / MEDIUM / PUSH: index of the first op after the subtree */
    uint32_t a;        /* primitives, MEDIUM: material id;  PUSH: first chain index */
    uint32_t b;        /* primitives, MEDIUM: graph object id;  PUSH: chain length  */
    double d[6];
} rtgpu_op;

/* one wrapper of hittable/mod.rs:135-331 */
enum rtgpu_xform_kind { RTGPU_XF_TRANSLATE = 0, RTGPU_XF_ROTATE_Y = 1, RTGPU_XF_ZOOM = 2, RTGPU_XF_FLIP = 3 };
typedef struct rtgpu_xform {
    uint32_t kind;
    uint32_t gid; /* graph object id of the wrapper */
    double p[3];  /* TRANSLATE: offset;  ROTATE_Y: sin, cos;  ZOOM: rate */
} rtgpu_xform;

/* context = every wrapper around an op, outermost first: chain[start .. start+len) */
typedef struct rtgpu_ctx {
    uint32_t start, len;
} rtgpu_ctx;

typedef struct rtgpu_material {
    uint32_t kind; /* rt_material_kind (include/rt_graph.h) */
    uint32_t tex;
    double p[4];
} rtgpu_material;

typedef struct rtgpu_texture {
    uint32_t kind; /* rt_texture_kind */
    int32_t a, b;
    uint32_t pad_;
    double p[4];
} rtgpu_texture;

typedef struct rtgpu_image {
    int32_t width, height;
    int64_t offset; /* bytes into texels; RGB8, rows bottom-up (texture/mod.rs:94-100) */
} rtgpu_image;

/* Perlin tables: texture/perlin.rs:8-13 */
typedef struct rtgpu_perlin {
    double randvec[256][3];
    int32_t perm_x[256], perm_y[256], perm_z[256];
} rtgpu_perlin;

/* members of the `lights` HittableList (main.rs:89).  Sphere and the three rects implement pdf_value/random
 * (sphere.rs:75-90, aarect.rs:74-93,157-176,240-259) and so does a HittableList nested in `lights`
 * (hittable/mod.rs:121-132: mean of its members' pdf_value, uniform pick for random) -- LIST, whose members are the
 * entries [first, first + count) of the same array.  Anything else keeps the trait defaults (pdf 0, direction
 * (1,0,0), hittable/mod.rs:62-67) and is lowered as DEAD.  The members of `lights` itself are the first
 * n_light_roots entries. */
enum rtgpu_light_kind { RTGPU_LIGHT_RECT_XY = 0, RTGPU_LIGHT_RECT_XZ = 1, RTGPU_LIGHT_RECT_YZ = 2,
                        RTGPU_LIGHT_SPHERE = 3, RTGPU_LIGHT_DEAD = 4, RTGPU_LIGHT_LIST = 5 };
#define RTGPU_MAX_LIGHT_NESTING 4 /* lists inside lists inside `lights` */
typedef struct rtgpu_light {
    uint32_t kind;
    uint32_t gid;
    double p[5]; /* rect: a0 a1 b0 b1 k;  sphere: center(3) radius;  LIST: first, count (exact small integers) */
} rtgpu_light;

typedef struct rtgpu_camera { /* basic/camera.rs:9-20 */
    double origin[3], lower_left_corner[3], horizontal[3], vertical[3];
    double u[3], v[3];
    double lens_radius, time0, time1;
} rtgpu_camera;

typedef struct rtgpu_scene_desc {
    const rtgpu_op *ops;
    const rtgpu_xform *xforms;
    const uint32_t *chain;
    const rtgpu_ctx *ctxs;
    const rtgpu_material *materials;
    const rtgpu_texture *textures;
    const rtgpu_perlin *perlins;
    const rtgpu_image *images;
    const uint8_t *texels;
    const rtgpu_light *lights;
    uint32_t n_ops, n_xforms, n_chain, n_ctxs;
    uint32_t n_materials, n_textures, n_perlins, n_images, n_lights;
    uint32_t max_push_depth; /* deepest PUSH nesting in the program (informative: rtgpu_scene_create counts it itself) */
    uint32_t n_light_roots;  /* members of `lights` itself = lights[0 .. n_light_roots); the rest are members of nested lists */
    uint32_t pad_;
    uint64_t n_texel_bytes;
    rtgpu_camera camera;
    double background[3]; /* main.rs:51 */
} rtgpu_scene_desc;

typedef struct rtgpu_scene rtgpu_scene;

/* A ray of a dumped batch: hit(r, t_min, t_max).  ray_color uses t_min = 0.001 and
 * t_max = f64::MAX (main.rs:243). */
typedef struct rtgpu_ray {
    double o[3], d[3];
    double tm, t_min, t_max;
} rtgpu_ray;

/* HitRecord of hittable/mod.rs:18-26 plus the identity of what was hit */
typedef struct rtgpu_hit {
    double t;
    double p[3], n[3];
    double u, v;
    int32_t hit, front_face;
    int32_t gid;      /* graph object id of the leaf (sphere, rect, box, triangle, medium) */
    int32_t face;     /* side of a box, in boxes.rs list order; else 0 */
    int32_t inst;     /* graph object id of the innermost wrapper, -1 if none */
    int32_t mat;      /* material id */
    uint32_t rng_pos; /* slots consumed from the ray's stream by ConstantMedium draws */
    uint32_t pad_;
} rtgpu_hit;

typedef struct rtgpu_render_params {
    int32_t width, height;
    int32_t spp_begin, spp_end; /* global sample indices [begin, end) rendered by this call */
    int32_t max_depth;          /* main.rs:39 */
    int32_t device;             /* CUDA device ordinal; -1 = current */
    uint64_t seed;
    int64_t max_paths_in_flight; /* size of the path pool (248 B of device memory per path); 0 = 15 Mi paths (3.9 GB).  Paths
                                    that end are replaced by new camera rays while samples remain (path regeneration). */
    void *stream;                /* cudaStream_t; NULL = default stream */
    int32_t accumulate;          /* 0: overwrite accum, 1: add to it */
    int32_t accum_on_device;     /* 0: accum is host memory, 1: device memory */
} rtgpu_render_params;

#define RTGPU_MAX_BOUNCE_STATS 64
typedef struct rtgpu_stats {
    uint64_t rays;    /* closest-hit queries issued by the integrator (main.rs:243) */
    uint64_t samples; /* camera paths */
    double ms_total;  /* CUDA-event time of the whole call on its stream (copies included) */
    double ms_extend; /* summed duration of the extend kernel launches */
    double ms_shade;  /* summed duration of the shade kernel launches */
    double ms_other;  /* ray generation + accumulation */
    uint64_t launches;
    uint64_t launches_extend; /* = iterations of the path pool */
    uint64_t rays_per_bounce[RTGPU_MAX_BOUNCE_STATS];
    uint64_t pool_paths; /* paths the pool holds (the most that are in flight at once) */
    uint64_t pool_bytes; /* device memory of the pool and of the pixel sums */
} rtgpu_stats;

const char *rtgpu_last_error(void);
int rtgpu_device_count(void);

/* Copies every array of `desc` to the device; the caller keeps ownership of the host arrays. */
rtgpu_status rtgpu_scene_create(const rtgpu_scene_desc *desc, int device, rtgpu_scene **out);
rtgpu_status rtgpu_scene_destroy(rtgpu_scene *scene);

/* Renders sample indices [spp_begin, spp_end) of every pixel and writes (or adds) the
 * per-pixel RGB radiance SUMS -- not averages; main.rs:150 -- to accum: 3*width*height
 * floats, pixel (x, y) at (y*width + x)*3, y up as in v = (y + xi)/(H-1) (main.rs:147).
 * Precision: upstream sums a pixel's samples in f64; here every sample is added as an exact 64-bit fixed-point
 * integer (2^-32 units), so the sum of this call is exact to 2^-33 per sample in any order -- the image is
 * bit-reproducible and does not depend on max_paths_in_flight -- and is rounded ONCE to the float32 buffer. */
rtgpu_status rtgpu_render(rtgpu_scene *scene, const rtgpu_render_params *params, float *accum,
                          rtgpu_stats *stats);

/* world.hit(rays[i], t_min, t_max) for a batch; ConstantMedium draws of ray i come from stream
 * BATCH(keys[i]) (include/rt_rng.h); keys == NULL means keys[i] = i.  Host pointers. */
rtgpu_status rtgpu_trace_batch(rtgpu_scene *scene, const rtgpu_ray *rays, int64_t n,
                               const uint64_t *keys, rtgpu_hit *out, double *ms_kernel);

/* The camera rays the renderer generates (main.rs:145-149, camera.rs:64-73) for sample k of
 * pixels [first_pixel, first_pixel+n) in y*width+x order.  Host pointers. */
rtgpu_status rtgpu_camera_rays(rtgpu_scene *scene, int32_t width, int32_t height, int64_t first_pixel,
                               int64_t n, int32_t k, uint64_t seed, rtgpu_ray *out, uint32_t *rng_pos);

/* One-sample radiance ray_color(...) of pixels [first_pixel, first_pixel+n), sample k, in f64.
 * Debug/parity entry point: same kernels as rtgpu_render (the n paths are all in flight at once; k_tail is not used).
 * out_rgb: 3*n doubles; out_rays (may be NULL): segments traced per path. */
rtgpu_status rtgpu_sample_radiance(rtgpu_scene *scene, int32_t width, int32_t height, int64_t first_pixel,
                                   int64_t n, int32_t k, int32_t max_depth, uint64_t seed, double *out_rgb,
                                   uint32_t *out_rays);

/* The frame of main.rs:109-183 on several GPUs of one box: device g of G renders ALL pixels for the sample indices
 * [spp_begin + g*n/G, spp_begin + (g+1)*n/G) (one host thread per device, where upstream has one per scanline block),
 * then ONE ncclReduce(sum, float32, 3*width*height) brings the per-device sums to devices[0] and from there to the host
 * buffer accum_root.  The scene is created on every device from `desc` and destroyed again.  devices == NULL means
 * 0 .. n_devices-1.  params->stream / accumulate / accum_on_device must be 0.  stats are summed over the devices
 * (ms_total: the slowest one); *ms_reduce is the wall time of the reduce.  NCCL is loaded at run time (libnccl.so.2). */
rtgpu_status rtgpu_multi_render(const rtgpu_scene_desc *desc, const rtgpu_render_params *params, int32_t n_devices,
                                const int32_t *devices, float *accum_root, rtgpu_stats *stats, double *ms_reduce);

/* write_color (main.rs:280-299) on every pixel + the row flip of main.rs:197: rgb8 is
 * width*height*3 bytes, top row first.  accum may be host or device memory. */
rtgpu_status rtgpu_resolve(const float *accum, int32_t width, int32_t height, int32_t spp,
                           int32_t accum_on_device, uint8_t *rgb8);

#ifdef __cplusplus
}
#endif
#endif /* RTGPU_H */
