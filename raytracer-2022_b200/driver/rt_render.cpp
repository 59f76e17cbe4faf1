/*
 * rt_render.cpp -- native driver: what `cargo run --release` does upstream (raytracer/src/main.rs:28-231),
 * with the hot path on B200s.
 *
 *   [1/5] consts -> camera            main.rs:33-78     (command line instead of editing consts)
 *   [2/5] scene + BvhNode::new_list   main.rs:89-90     rt::build_config (host/scenes.cpp), rt::lower
 *         render                      main.rs:93-163    rtgpu_multi_render: one host thread per GPU, spp-sharded
 *   [3/5] collect                     main.rs:171-183   ... and ONE ncclReduce(sum) of the float accumulation buffers
 *   [4/5] write_color + row flip      main.rs:191-201   rtgpu_resolve
 *   [5/5] image file                  main.rs:213-221   JPEG at quality 100 (host/jpeg_encode.cpp) when --out ends in .jpg, else binary PPM
 *
 * Build: see __graft_entry__.build().  Links librthost.so and librtgpu.so (which loads NCCL at run time).
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../host/rt_host.h"

namespace {

struct Args {
    std::string scene = "final_scene", out = "output.ppm", assets = "assets/";
    int width = 800, height = 800, spp = 100, depth = 50, devices = 1;
    uint64_t seed = 2022;
};

bool parse(int argc, char **argv, Args &a) {
    for (int i = 1; i < argc; ++i) {
        std::string k = argv[i];
        auto val = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (k == "--scene") a.scene = val();
        else if (k == "--out") a.out = val();
        else if (k == "--assets") a.assets = val();
        else if (k == "--width") a.width = atoi(val());
        else if (k == "--height") a.height = atoi(val());
        else if (k == "--spp") a.spp = atoi(val());
        else if (k == "--depth") a.depth = atoi(val());
        else if (k == "--devices") a.devices = atoi(val());
        else if (k == "--seed") a.seed = strtoull(val(), nullptr, 10);
        else {
            fprintf(stderr,
                    "usage: rt_render [--scene random_scene|cornell_box|cornell_smoke|cornell_book3|final_scene|obj_cornell|wwscene]\n"
                    "                 [--width W --height H --spp N --depth D --seed S --devices G --assets DIR --out FILE.ppm|FILE.jpg]\n");
            return false;
        }
    }
    if (!a.assets.empty() && a.assets.back() != '/') a.assets += '/';
    return a.width >= 2 && a.height >= 2 && a.spp >= 1 && a.devices >= 1;
}

} // namespace

int main(int argc, char **argv) {
    Args a;
    if (!parse(argc, argv, a)) return 2;
    auto t_begin = std::chrono::steady_clock::now();
    printf("[1/5] Initializing...\nIMAGE SIZE: %dx%d\nSAMPLE PER PIXEL: %d\nMAX DEPTH: %d\nSCENE: %s (seed %llu)\n", a.width, a.height,
           a.spp, a.depth, a.scene.c_str(), (unsigned long long)a.seed);
    int have = rtgpu_device_count();
    if (have < 1) {
        fprintf(stderr, "no CUDA device: the hot path has no CPU fallback\n");
        return 1;
    }
    if (a.devices > have) a.devices = have;
    if (a.devices > a.spp) a.devices = a.spp;

    // ---- scene: the reference's constructors, BvhNode::new_list, then lowering ----
    rt::Lowered lowered;
    try {
        rt::Scene scene(a.seed);
        rt::SceneOptions opt;
        opt.asset_dir = a.assets;
        rt::build_config(scene, a.scene, (double)a.width / (double)a.height, opt);
        rt_graph g = scene.view();
        lowered = rt::lower(g);
        printf("[2/5] Rendering with %d GPU(s)... (%d objects -> %zu ops, %.1f KB)\n", a.devices, g.n_objects, lowered.ops.size(),
               lowered.ops.size() * 64 / 1024.0);
    } catch (const rt::Error &e) {
        fprintf(stderr, "scene error: %s\n", e.what.c_str());
        return 1;
    }
    rtgpu_scene_desc desc = lowered.desc();
    const size_t n_float = (size_t)3 * a.width * a.height;

    // ---- render + collect: rtgpu_multi_render shards the samples over the GPUs (device g takes sample indices
    //      [g*spp/G, (g+1)*spp/G) of every pixel) and sums the per-GPU radiance sums with one ncclReduce ----
    std::vector<float> accum(n_float);
    rtgpu_render_params p;
    memset(&p, 0, sizeof p);
    p.width = a.width;
    p.height = a.height;
    p.spp_begin = 0;
    p.spp_end = a.spp;
    p.max_depth = a.depth;
    p.device = -1;
    p.seed = a.seed;
    rtgpu_stats stats;
    double ms_reduce = 0.;
    auto t_render = std::chrono::steady_clock::now();
    if (rtgpu_multi_render(&desc, &p, a.devices, nullptr, accum.data(), &stats, &ms_reduce) != RTGPU_OK) {
        fprintf(stderr, "render: %s\n", rtgpu_last_error());
        return 1;
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_render).count();
    printf("[3/5] Collecting results... (one reduce of %.1f MB: %.2f ms)\n", n_float * 4 / 1e6, ms_reduce);
    /* device time of the slowest GPU; the wall time also holds scene upload and, with several GPUs, ncclCommInitAll (seconds) */
    printf("      render %.3f s on the device (%.1f Mrays/s, %.1f Msamples/s, %.2f rays/sample); %.3f s wall with set-up\n", stats.ms_total * 1e-3,
           stats.rays / (stats.ms_total * 1e3), stats.samples / (stats.ms_total * 1e3), (double)stats.rays / (double)(stats.samples ? stats.samples : 1),
           sec);

    // ---- image ----
    printf("[4/5] Generating image...\n");
    std::vector<uint8_t> rgb((size_t)3 * a.width * a.height);
    if (rtgpu_resolve(accum.data(), a.width, a.height, a.spp, 0, rgb.data()) != RTGPU_OK) {
        fprintf(stderr, "resolve: %s\n", rtgpu_last_error());
        return 1;
    }
    printf("[5/5] Output image as \"%s\"\n", a.out.c_str());
    auto ends_with = [&](const char *suf) { size_t n = strlen(suf); return a.out.size() >= n && a.out.compare(a.out.size() - n, n, suf) == 0; };
    if (ends_with(".jpg") || ends_with(".jpeg")) { // main.rs:213-221: JPEG at quality 100
        std::string err;
        if (!rt::encode_jpeg(a.out, rgb.data(), a.width, a.height, 100, err)) {
            fprintf(stderr, "%s\n", err.c_str());
            return 1;
        }
    } else {
        FILE *f = fopen(a.out.c_str(), "wb");
        if (!f) {
            fprintf(stderr, "Outputting image fails.\n");
            return 1;
        }
        fprintf(f, "P6\n%d %d\n255\n", a.width, a.height);
        fwrite(rgb.data(), 1, rgb.size(), f);
        fclose(f);
    }
    printf("\n      All Work Done.\n      Elapsed Time: %.2f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count());
    return 0;
}
