/*
 * oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  See oracle.cpp for what it restates.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>

#include "../include/rt_graph.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

typedef struct orc_ray {
    double o[3], d[3];
    double tm, t_min, t_max;
} orc_ray;

typedef struct orc_hit {
    double t;
    double p[3], n[3];
    double u, v;
    int32_t hit, front_face;
    int32_t gid;  /* graph object id of the leaf that produced the record */
    int32_t face; /* side index for Boxes (list order of boxes.rs:24-69), else 0 */
    int32_t inst; /* graph object id of the innermost wrapper around the leaf, -1 if none */
    int32_t mat;  /* graph material id */
    uint32_t rng_pos; /* slots consumed from the ray's stream (ConstantMedium draws) */
    uint32_t pad_;
} orc_hit;

typedef struct orc_stats {
    uint64_t rays;    /* closest-hit queries issued by ray_color (main.rs:243) */
    uint64_t samples; /* camera paths */
    double seconds;   /* render loop only */
} orc_stats;

orc_scene *orc_scene_create(const rt_graph *g);
void orc_scene_destroy(orc_scene *s);
const char *orc_last_error(void);

/* world.hit(ray, t_min, t_max) for a dumped ray batch; ray i draws from stream BATCH(keys[i]) */
int orc_trace_batch(const orc_scene *s, int64_t n, const orc_ray *rays, const uint64_t *keys,
                    orc_hit *out);

/* main.rs:109-163 pixel loop, sample indices [spp_begin, spp_end), `threads` scanline blocks.
 * accum: 3*w*h doubles, pixel (x, y) at (y*w + x)*3 with y up (v = (y+xi)/(h-1)). */
int orc_render(const orc_scene *s, int w, int h, int spp_begin, int spp_end, int max_depth,
               uint64_t seed, int threads, double *accum, orc_stats *stats);
/* same, restricted to rows [y0,y1) and columns [x0,x1) (bounded CPU-baseline samples) */
int orc_render_region(const orc_scene *s, int w, int h, int x0, int x1, int y0, int y1,
                      int spp_begin, int spp_end, int max_depth, uint64_t seed, int threads,
                      double *accum, orc_stats *stats);
/* one sample: ray_color of pixel (x,y), sample index k; returns rays traced */
int orc_sample(const orc_scene *s, int w, int h, int x, int y, int k, int max_depth,
               uint64_t seed, double rgb[3]);
/* camera ray of pixel (x,y), sample k (main.rs:145-149 + camera.rs:64-73); rng_pos = slots used */
int orc_camera_ray(const orc_scene *s, int w, int h, int x, int y, int k, uint64_t seed,
                   orc_ray *out, uint32_t *rng_pos);
/* dump the rays ray_color issues at bounce `bounce` (0 = camera rays) for the first samples of
 * the image in scan order, until `cap` rays are written; returns the count */
int64_t orc_dump_rays(const orc_scene *s, int w, int h, int spp, int max_depth, uint64_t seed,
                      int bounce, int64_t cap, orc_ray *out);

/* main.rs:280-299 */
void orc_write_color(const double rgb_sum[3], int spp, uint8_t out[3]);

/* pieces exposed for known-answer tests */
double orc_lights_pdf_value(const orc_scene *s, const double o[3], const double v[3]);
void orc_lights_random(const orc_scene *s, const double o[3], uint64_t key, double out[3]);
void orc_texture_value(const orc_scene *s, int tex, double u, double v, const double p[3],
                       double out[3]);
double orc_perlin_noise(const orc_scene *s, int perlin, const double p[3]);
double orc_perlin_turb(const orc_scene *s, int perlin, const double p[3], int depth);
double orc_reflectance(double cos, double ref_idx);
int orc_aabb_hit(const double bmin[3], const double bmax[3], const orc_ray *r);
void orc_camera_derive(rt_camera *cam); /* Camera::new: fills the derived fields */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double orc_stream_gen(uint64_t seed, uint32_t s0, uint32_t s1, uint32_t domain, uint32_t pos);
uint32_t orc_stream_index(uint64_t seed, uint32_t s0, uint32_t s1, uint32_t domain, uint32_t pos,
                          uint32_t n, uint32_t *pos_after);

/* BvhNode::new_vec (bvh/mod.rs:34-81) over objects ids[0..n) of the graph the scene was made
 * from, axis draws from stream SCENE(seed, s0, 0) starting at slot `pos`.  Writes the tree in
 * pre-order: entry = -1 for a BvhNode followed by its two subtrees, else the leaf's object id
 * (a singleton node lists its leaf twice).  boxes (6 doubles per BvhNode, pre-order) may be NULL.
 * Returns the number of entries, or -1. */
int64_t orc_bvh_build(const orc_scene *s, const int32_t *ids, int32_t n, double time0,
                      double time1, uint64_t seed, uint32_t s0, uint32_t pos, int32_t *out,
                      int64_t cap, double *boxes, uint32_t *pos_after);
/* the same pre-order listing for an existing RT_OBJ_BVH_NODE of the graph */
int64_t orc_bvh_listing(const orc_scene *s, int32_t root, int32_t *out, int64_t cap, double *boxes);

#ifdef __cplusplus
}
#endif
#endif
