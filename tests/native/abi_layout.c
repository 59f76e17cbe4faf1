/*
 * abi_layout.c -- keeps the unbuilt Rust binding of INTEGRATION.md honest.  No Rust toolchain exists in this image, so the
 * `#[repr(C)]` structs of the `rtgpu-sys` stub cannot be compiled against include/rtgpu.h.  A `#[repr(C)]` struct has the C
 * layout of its fields in order; the numbers below are the sizes and offsets THE STUB'S DECLARATIONS imply (u32/i32/f32 = 4,
 * u64/i64/f64/pointers = 8, natural alignment).  If include/rtgpu.h drifts from the stub, this file stops compiling
 * (tests/test_host_and_abi.py builds it with `gcc -std=c11 -c`), and the stub in INTEGRATION.md section 1 must follow.
 */
#include <stddef.h>

#include "../../include/rtgpu.h"

#define SIZE(T, n) _Static_assert(sizeof(T) == (n), "sizeof(" #T ") != " #n)
#define OFF(T, f, n) _Static_assert(offsetof(T, f) == (n), "offsetof(" #T ", " #f ") != " #n)

/* RtgpuOp { kind_ctx: u32, skip: u32, a: u32, b: u32, d: [f64; 6] } */
SIZE(rtgpu_op, 64); OFF(rtgpu_op, kind_ctx, 0); OFF(rtgpu_op, skip, 4); OFF(rtgpu_op, a, 8); OFF(rtgpu_op, b, 12); OFF(rtgpu_op, d, 16);
/* RtgpuXform { kind: u32, gid: u32, p: [f64; 3] } */
SIZE(rtgpu_xform, 32); OFF(rtgpu_xform, kind, 0); OFF(rtgpu_xform, gid, 4); OFF(rtgpu_xform, p, 8);
/* RtgpuCtx { start: u32, len: u32 } */
SIZE(rtgpu_ctx, 8); OFF(rtgpu_ctx, start, 0); OFF(rtgpu_ctx, len, 4);
/* RtgpuMaterial { kind: u32, tex: u32, p: [f64; 4] } */
SIZE(rtgpu_material, 40); OFF(rtgpu_material, kind, 0); OFF(rtgpu_material, tex, 4); OFF(rtgpu_material, p, 8);
/* RtgpuTexture { kind: u32, a: i32, b: i32, pad: u32, p: [f64; 4] } */
SIZE(rtgpu_texture, 48); OFF(rtgpu_texture, kind, 0); OFF(rtgpu_texture, a, 4); OFF(rtgpu_texture, b, 8); OFF(rtgpu_texture, pad_, 12);
OFF(rtgpu_texture, p, 16);
/* RtgpuImage { width: i32, height: i32, offset: i64 } */
SIZE(rtgpu_image, 16); OFF(rtgpu_image, width, 0); OFF(rtgpu_image, height, 4); OFF(rtgpu_image, offset, 8);
/* RtgpuPerlin { randvec: [[f64; 3]; 256], perm_x: [i32; 256], perm_y: [i32; 256], perm_z: [i32; 256] } */
SIZE(rtgpu_perlin, 256 * 24 + 3 * 256 * 4); OFF(rtgpu_perlin, randvec, 0); OFF(rtgpu_perlin, perm_x, 6144); OFF(rtgpu_perlin, perm_y, 7168);
OFF(rtgpu_perlin, perm_z, 8192);
/* RtgpuLight { kind: u32, gid: u32, p: [f64; 5] } */
SIZE(rtgpu_light, 48); OFF(rtgpu_light, kind, 0); OFF(rtgpu_light, gid, 4); OFF(rtgpu_light, p, 8);
/* RtgpuCamera { origin, lower_left_corner, horizontal, vertical, u, v: [f64; 3], lens_radius, time0, time1: f64 } */
SIZE(rtgpu_camera, 168); OFF(rtgpu_camera, origin, 0); OFF(rtgpu_camera, lower_left_corner, 24); OFF(rtgpu_camera, horizontal, 48);
OFF(rtgpu_camera, vertical, 72); OFF(rtgpu_camera, u, 96); OFF(rtgpu_camera, v, 120); OFF(rtgpu_camera, lens_radius, 144);
OFF(rtgpu_camera, time0, 152); OFF(rtgpu_camera, time1, 160);
/* RtgpuSceneDesc { ops, xforms, chain, ctxs, materials, textures, perlins, images, texels, lights: *const _ (10 pointers),
 *                  n_ops, n_xforms, n_chain, n_ctxs, n_materials, n_textures, n_perlins, n_images, n_lights, max_push_depth,
 *                  n_light_roots, pad: u32 (12 of them), n_texel_bytes: u64, camera: RtgpuCamera, background: [f64; 3] } */
SIZE(rtgpu_scene_desc, 80 + 48 + 8 + 168 + 24); OFF(rtgpu_scene_desc, ops, 0); OFF(rtgpu_scene_desc, xforms, 8); OFF(rtgpu_scene_desc, chain, 16);
OFF(rtgpu_scene_desc, ctxs, 24); OFF(rtgpu_scene_desc, materials, 32); OFF(rtgpu_scene_desc, textures, 40); OFF(rtgpu_scene_desc, perlins, 48);
OFF(rtgpu_scene_desc, images, 56); OFF(rtgpu_scene_desc, texels, 64); OFF(rtgpu_scene_desc, lights, 72); OFF(rtgpu_scene_desc, n_ops, 80);
OFF(rtgpu_scene_desc, n_xforms, 84); OFF(rtgpu_scene_desc, n_chain, 88); OFF(rtgpu_scene_desc, n_ctxs, 92); OFF(rtgpu_scene_desc, n_materials, 96);
OFF(rtgpu_scene_desc, n_textures, 100); OFF(rtgpu_scene_desc, n_perlins, 104); OFF(rtgpu_scene_desc, n_images, 108);
OFF(rtgpu_scene_desc, n_lights, 112); OFF(rtgpu_scene_desc, max_push_depth, 116); OFF(rtgpu_scene_desc, n_light_roots, 120);
OFF(rtgpu_scene_desc, pad_, 124); OFF(rtgpu_scene_desc, n_texel_bytes, 128); OFF(rtgpu_scene_desc, camera, 136);
OFF(rtgpu_scene_desc, background, 304);
/* RtgpuRay { o: [f64; 3], d: [f64; 3], tm: f64, t_min: f64, t_max: f64 } */
SIZE(rtgpu_ray, 72); OFF(rtgpu_ray, o, 0); OFF(rtgpu_ray, d, 24); OFF(rtgpu_ray, tm, 48); OFF(rtgpu_ray, t_min, 56); OFF(rtgpu_ray, t_max, 64);
/* RtgpuHit { t: f64, p: [f64; 3], n: [f64; 3], u: f64, v: f64, hit: i32, front_face: i32, gid: i32, face: i32, inst: i32, mat: i32,
 *            rng_pos: u32, pad: u32 } */
SIZE(rtgpu_hit, 104); OFF(rtgpu_hit, t, 0); OFF(rtgpu_hit, p, 8); OFF(rtgpu_hit, n, 32); OFF(rtgpu_hit, u, 56); OFF(rtgpu_hit, v, 64);
OFF(rtgpu_hit, hit, 72); OFF(rtgpu_hit, front_face, 76); OFF(rtgpu_hit, gid, 80); OFF(rtgpu_hit, face, 84); OFF(rtgpu_hit, inst, 88);
OFF(rtgpu_hit, mat, 92); OFF(rtgpu_hit, rng_pos, 96); OFF(rtgpu_hit, pad_, 100);
/* RtgpuRenderParams { width, height, spp_begin, spp_end, max_depth, device: i32, seed: u64, max_paths_in_flight: i64,
 *                     stream: *mut c_void, accumulate: i32, accum_on_device: i32 } */
SIZE(rtgpu_render_params, 56); OFF(rtgpu_render_params, width, 0); OFF(rtgpu_render_params, height, 4); OFF(rtgpu_render_params, spp_begin, 8);
OFF(rtgpu_render_params, spp_end, 12); OFF(rtgpu_render_params, max_depth, 16); OFF(rtgpu_render_params, device, 20);
OFF(rtgpu_render_params, seed, 24); OFF(rtgpu_render_params, max_paths_in_flight, 32); OFF(rtgpu_render_params, stream, 40);
OFF(rtgpu_render_params, accumulate, 48); OFF(rtgpu_render_params, accum_on_device, 52);
/* RtgpuStats { rays: u64, samples: u64, ms_total, ms_extend, ms_shade, ms_other: f64, launches: u64, launches_extend: u64,
 *              rays_per_bounce: [u64; 64], pool_paths: u64, pool_bytes: u64 } */
SIZE(rtgpu_stats, 64 + 64 * 8 + 16); OFF(rtgpu_stats, rays, 0); OFF(rtgpu_stats, samples, 8); OFF(rtgpu_stats, ms_total, 16);
OFF(rtgpu_stats, ms_extend, 24); OFF(rtgpu_stats, ms_shade, 32); OFF(rtgpu_stats, ms_other, 40); OFF(rtgpu_stats, launches, 48);
OFF(rtgpu_stats, launches_extend, 56); OFF(rtgpu_stats, rays_per_bounce, 64); OFF(rtgpu_stats, pool_paths, 576); OFF(rtgpu_stats, pool_bytes, 584);
/* status codes used by the stub's `check()` */
_Static_assert(RTGPU_OK == 0 && RTGPU_ERR_INVALID_SCENE == 1 && RTGPU_ERR_INVALID_ARG == 2 && RTGPU_ERR_CUDA == 3 && RTGPU_ERR_OOM == 4 &&
                   RTGPU_ERR_NCCL == 5, "status codes");
/* op / xform / light kinds the Rust lowering emits */
_Static_assert(RTGPU_OP_NODE == 0 && RTGPU_OP_SPHERE == 1 && RTGPU_OP_MSPHERE == 2 && RTGPU_OP_RECT_XY == 3 && RTGPU_OP_RECT_XZ == 4 &&
                   RTGPU_OP_RECT_YZ == 5 && RTGPU_OP_BOX == 6 && RTGPU_OP_TRI == 7 && RTGPU_OP_MEDIUM == 8 && RTGPU_OP_PUSH == 9 &&
                   RTGPU_OP_POP == 10 && RTGPU_OP_RING == 11, "op kinds");
_Static_assert(RTGPU_XF_TRANSLATE == 0 && RTGPU_XF_ROTATE_Y == 1 && RTGPU_XF_ZOOM == 2 && RTGPU_XF_FLIP == 3, "xform kinds");
_Static_assert(RTGPU_LIGHT_RECT_XY == 0 && RTGPU_LIGHT_RECT_XZ == 1 && RTGPU_LIGHT_RECT_YZ == 2 && RTGPU_LIGHT_SPHERE == 3 && RTGPU_LIGHT_DEAD == 4 &&
                   RTGPU_LIGHT_LIST == 5, "light kinds");
