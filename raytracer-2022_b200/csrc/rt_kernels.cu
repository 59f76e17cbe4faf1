/*
 * rt_kernels.cu -- wavefront path tracer for sm_100a and the C ABI of include/rtgpu.h.
 *
 * A render keeps a bounded POOL of paths in flight (path regeneration): one queue entry per live path, two queues
 * that swap roles.  Per iteration, in stream order and without a host round trip (kernels that do not depend on each other --
 * k_regen and the sort, the class kernels among themselves -- go to side streams and are joined with events):
 *
 *   k_iter_begin  resets the per-iteration counters on the device and publishes the length the queue is topped up to
 *   k_regen       tops the queue up with new camera rays while samples remain (main.rs:145-149, camera.rs:64-73),
 *                 so every launch below carries a full queue whatever the depth of its paths
 *   (sort)        where the program is large enough to gain from it, the survivors are walked in the order of a 32-bit
 *                 origin-cell / direction key that the shade kernels wrote next to every ray; only (key, entry) pairs are
 *                 sorted, the records stay put
 *   k_extend      closest hit of every ray against the flat traversal program
 *                 (BvhNode::hit, hittable/bvh/mod.rs:86-101, unrolled -- rt_device.cuh)
 *   k_classify    bins the hits by material class into index lists
 *   k_shade<C>    one kernel per class: HitRecord of the winner, emitted/scatter/MixturePdf (ray_color,
 *                 main.rs:233-278); survivors are appended to the other queue with one warp-aggregated atomic
 *                 (ballot + popc prefix) -- the compaction between bounces; a path that ends adds its radiance to
 *                 its pixel
 *
 * Once every path has been generated and only a few thousand are alive, k_tail runs them to their end in one launch.
 * Queue lengths, the sample cursor and the ray counters live in a control block on the device (WaveCtl); the host
 * reads a 16-byte snapshot of it asynchronously every iteration and only uses it to size the sort and to stop issuing iterations.
 * Pixel sums are exact: every contribution is added as a 64-bit fixed-point integer (2^-32 units), so the sum does
 * not depend on the order in which paths end and the image is bit-reproducible (k_finish converts to float).
 * Randomness: one Philox slot stream per path, keyed (seed, pixel, sample) -- include/rt_rng.h.
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "rt_device.cuh"

namespace {

thread_local std::string g_err;

#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) {                                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                                            \
            return e_ == cudaErrorMemoryAllocation ? RTGPU_ERR_OOM : RTGPU_ERR_CUDA;                               \
        }                                                                                                          \
    } while (0)

#ifndef RT_THREADS
#define RT_THREADS 128
#endif
constexpr int kThreads = RT_THREADS;
#ifndef RT_SHADE_THREADS
#define RT_SHADE_THREADS 64
#endif
#ifndef RT_SHADE_BLOCKS_PER_SM
#define RT_SHADE_BLOCKS_PER_SM 12 /* 24 warps/SM at <= 80 registers (RT_SHADE_MIN_BLOCKS).  A/B on C3, ms in the shade stage:
                                     8 blocks/SM (<= 128 registers) 30.1, 10 (<= 96) 28.8, 12 (<= 80) 27.8 */
#endif
constexpr int kShadeThreads = RT_SHADE_THREADS;

#ifndef RT_BLOCKS_PER_SM
#define RT_BLOCKS_PER_SM 8
#endif
constexpr int kBlocksPerSM = RT_BLOCKS_PER_SM;

/* One queue entry = a 64-byte ray record + a 32-byte state record (AoS): k_extend reads the ray record of
 * consecutive entries with four 128-bit loads (fully coalesced), and the shade kernels, which reach entries
 * through the per-class index lists, touch 3 sectors per path instead of the 12 an SoA layout costs. */
struct __align__(16) RayRec {
    double ox, oy, oz, dx, dy, dz, tm;
    uint32_t pid; /* path id inside the render call: sample-major, pixel = pid % npix */
    uint32_t rng; /* next slot of the path's Philox stream (low 20 bits) | segments the path may still trace << 20 */
};
#define RT_POS_BITS 20
#define RT_POS_MASK ((1u << RT_POS_BITS) - 1u)
struct __align__(16) StateRec {
    double bx, by, bz, pad_; /* throughput */
};
static_assert(sizeof(RayRec) == 64 && sizeof(StateRec) == 32, "queue records");
struct PathQueue {
    RayRec *ray;
    StateRec *st;
};
#ifndef RT_STREAM_HINTS
#define RT_STREAM_HINTS 1 /* queue entries are read once and written once by k_extend: keep them out of L1, which holds the program (+2 %) */
#endif
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
#if RT_STREAM_HINTS
    double2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
#else
    return *p;
#endif
}
/* k_extend's read of a queue entry (the shade kernels gather theirs through the class lists and use plain loads) */
__device__ __forceinline__ RayRec ld_ray_stream(const RayRec *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = ld_stream(q), b = ld_stream(q + 1), c = ld_stream(q + 2), d = ld_stream(q + 3);
    RayRec r;
    r.ox = a.x; r.oy = a.y; r.oz = b.x; r.dx = b.y; r.dy = c.x; r.dz = c.y; r.tm = d.x;
    r.pid = (uint32_t)__double2loint(d.y);
    r.rng = (uint32_t)__double2hiint(d.y);
    return r;
}
__device__ __forceinline__ RayRec ld_ray(const RayRec *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = q[0], b = q[1], c = q[2], d = q[3];
    RayRec r;
    r.ox = a.x; r.oy = a.y; r.oz = b.x; r.dx = b.y; r.dy = c.x; r.dz = c.y; r.tm = d.x;
    r.pid = (uint32_t)__double2loint(d.y);
    r.rng = (uint32_t)__double2hiint(d.y);
    return r;
}
__device__ __forceinline__ void st_ray(RayRec *p, V3 o, V3 d, double tm, uint32_t pid, uint32_t rng) {
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(o.x, o.y);
    q[1] = make_double2(o.z, d.x);
    q[2] = make_double2(d.y, d.z);
    q[3] = make_double2(tm, __hiloint2double((int)rng, (int)pid));
}
__device__ __forceinline__ V3 ld_beta(const StateRec *p) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
    double2 a = q[0], b = q[1];
    return v3(a.x, a.y, b.x);
}
__device__ __forceinline__ void st_beta(StateRec *p, V3 b) {
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(b.x, b.y);
    q[1] = make_double2(b.z, 0.);
}
/* what k_extend leaves for the shade kernels: ONE 16-byte record per queue entry (one sector write per ray) */
struct __align__(16) HitRec16 {
    double t;
    uint32_t op;  /* unit index | box face << 29; RT_NO_HIT = miss */
    uint32_t rng; /* the path's Philox slot after the walk (ConstantMedium draws), low 20 bits | shade class << 29 */
};
struct HitBuf {
    HitRec16 *rec;
};
__device__ __forceinline__ HitRec16 ld_hit(const HitRec16 *p) {
    uint4 v = *reinterpret_cast<const uint4 *>(p);
    HitRec16 h;
    h.t = __hiloint2double((int)v.y, (int)v.x);
    h.op = v.z;
    h.rng = v.w;
    return h;
}
struct WaveParams {
    uint64_t seed;
    uint32_t first_pixel, npix_wave, sample_begin, n_paths;
    int32_t width, height;
    uint32_t max_depth, pool; /* segments per path (main.rs:39); paths kept in flight */
};
/* control block of a render on the device */
struct WaveCtl {
    uint32_t count[2];       /* entries of queue 0 / 1 */
    uint32_t gen_next;       /* next path id to generate */
    uint32_t fetch;          /* k_extend's ray cursor */
    uint32_t class_count[8]; /* entries of the class lists of this iteration */
    unsigned long long rays; /* closest-hit queries so far */
    uint32_t iter;
    uint32_t n_old; /* entries of the current queue that come from the last bounce (the rest are new camera rays) */
    unsigned long long len_hist[RTGPU_MAX_BOUNCE_STATS + 1]; /* paths by number of segments traced */
    uint32_t gen_first; /* id of the first path this iteration generates (k_regen makes count[cur] - n_old of them) */
};

__device__ __forceinline__ void path_rng(rt_rng &rng, const WaveParams &wp, uint32_t pid, uint32_t pos) {
    uint32_t pixel = wp.first_pixel + pid % wp.npix_wave;
    uint32_t sample = wp.sample_begin + pid / wp.npix_wave;
    rng.init(wp.seed, pixel, sample, RT_DOMAIN_PATH, pos);
}

/* main.rs:145-149 + Camera::get_ray, camera.rs:64-73 */
__device__ __forceinline__ Ray camera_ray(const rtgpu_camera &cam, int x, int y, int w, int h, rt_rng &rng) {
    double rand_u = rng.gen();
    double rand_v = rng.gen();
    double s = ((double)x + rand_u) / (double)(w - 1);
    double t = ((double)y + rand_v) / (double)(h - 1);
    V3 cu = v3(cam.u[0], cam.u[1], cam.u[2]), cv = v3(cam.v[0], cam.v[1], cam.v[2]);
    V3 origin = v3(cam.origin[0], cam.origin[1], cam.origin[2]);
    V3 llc = v3(cam.lower_left_corner[0], cam.lower_left_corner[1], cam.lower_left_corner[2]);
    V3 horizontal = v3(cam.horizontal[0], cam.horizontal[1], cam.horizontal[2]);
    V3 vertical = v3(cam.vertical[0], cam.vertical[1], cam.vertical[2]);
    V3 rd = random_in_unit_disk(rng) * cam.lens_radius; /* sampled even when the aperture is 0 */
    V3 offset = cu * rd.x + cv * rd.y;
    Ray r;
    r.o = origin + offset;
    r.d = llc + horizontal * s + vertical * t - origin - offset;
    r.tm = rng.gen_range(cam.time0, cam.time1);
    return r;
}

/* start of an iteration: the rays of queue `cur` are about to be extended */
/* Start of an iteration: resets the per-iteration counters and publishes the length the queue will have once k_regen has
 * topped it up -- k_regen only fills entries [n_old, count[cur]), so it can run beside the sort of the survivors. */
__global__ void k_iter_begin(WaveCtl *ctl, int cur, WaveParams wp) {
    if (threadIdx.x == 0) {
        ctl->fetch = 0;
        ctl->count[cur ^ 1] = 0;
        ctl->iter += 1;
        const uint32_t cnt = ctl->count[cur], first = ctl->gen_next;
        const uint32_t room = wp.pool > cnt ? wp.pool - cnt : 0u, left = wp.n_paths - first;
        const uint32_t add = room < left ? room : left;
        ctl->n_old = cnt;
        ctl->gen_first = first;
        ctl->count[cur] = cnt + add;
        ctl->gen_next = first + add;
        ctl->rays += (unsigned long long)(cnt + add);
    }
    if (threadIdx.x < 8) ctl->class_count[threadIdx.x] = 0;
}
/* Only the survivors of the last bounce are sorted: the new camera rays behind them are neighbouring pixels already.  The
 * sort runs over a host-side upper bound of the number of survivors; entries past them get the largest key, and since the
 * pair sort is stable and the values ascend they stay where they are (perm[k] = k there). */
__global__ void k_pad_keys(const WaveCtl *ctl, uint32_t *keys, uint32_t pad_to) {
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = ctl->n_old + blockIdx.x * blockDim.x + threadIdx.x; i < pad_to; i += stride) keys[i] = 0xffffffffu;
}
/* world.hit(r, 0.001, f64::MAX), main.rs:243 -- persistent warps, see extend_warp_loop */
#ifndef RT_EXTEND_MIN_BLOCKS
#define RT_EXTEND_MIN_BLOCKS 5 /* NEST 1 (boundaries walked by the loop: a variable t_min, a range end and a depth per lane): 96 registers,
                                 20 warps/SM -- cornell_smoke, ms in k_extend: 4: 26.0, 5: 24.9, 6: 25.5 (with the out-of-line walk: 29.0) */
#endif
#ifndef RT_EXTEND_MIN_BLOCKS0
#define RT_EXTEND_MIN_BLOCKS0 7 /* NEST 0 (no call in the loop): 72 registers, 28 warps/SM -- A/B on C3 (64 spp), ms in k_extend:
                                  4: 78.0, 5: 68.9, 6: 68.5, 7: 67.4, 8: 68.8; the other scenes agree (7 best on C1, C2, wwscene) */
#endif
struct QueueIO {
    static constexpr bool kConstTmin = true; /* world.hit(r, 0.001, f64::MAX), main.rs:243 */
    PathQueue q;
    HitBuf hb;
    WaveParams wp;
    const uint32_t *perm; /* sorted order -> queue entry (ray reordering without moving the records), or nullptr */
    uint32_t n_perm;      /* entries covered by perm (the survivors of the last bounce); the new camera rays behind them are walked in place */
    __device__ __forceinline__ void load(uint32_t idx, V3 &o, V3 &d, double &tm, double &tmin, double &tmax, uint32_t &a, uint32_t &b,
                                         uint32_t &pos, uint32_t &cur) {
        if (idx < n_perm) idx = __ldg(perm + idx);
        RayRec rr = ld_ray_stream(q.ray + idx);
        o = v3(rr.ox, rr.oy, rr.oz);
        d = v3(rr.dx, rr.dy, rr.dz);
        tm = rr.tm;
        tmin = 0.001;
        tmax = RT_F64_MAX;
        a = rr.pid; /* the path's Philox stream: (pixel, sample) follow from the path id */
        b = 0;
        pos = rr.rng & RT_POS_MASK;
        cur = idx;
    }
    __device__ __forceinline__ int refill_at() const { return perm ? RT_EXTEND_REFILL_SORTED : RT_EXTEND_REFILL; }
    /* pulls a queue entry that some warp will ask for a few microseconds from now into L2 */
    __device__ __forceinline__ void prefetch(uint32_t idx) {
        if (!perm) asm volatile("prefetch.global.L2 [%0];" ::"l"(q.ray + idx));
    }
    __device__ __forceinline__ StreamDraw stream(uint32_t a, uint32_t, uint32_t pos) const { /* the same stream, resolved */
        uint32_t pixel = wp.first_pixel + a % wp.npix_wave, sample = wp.sample_begin + a / wp.npix_wave;
        return StreamDraw{wp.seed, pixel, sample, RT_DOMAIN_PATH, pos};
    }
    __device__ __forceinline__ double draw(uint32_t a, uint32_t, uint32_t &pos) const { /* the path's stream resumes at slot `pos` */
        rt_rng rng;
        path_rng(rng, wp, a, pos);
        double u = rng.gen();
        pos = rng.pos;
        return u;
    }
    __device__ __forceinline__ void store(uint32_t idx, const Closest &c, uint32_t pos) {
        uint4 v;
        v.x = (uint32_t)__double2loint(c.t);
        v.y = (uint32_t)__double2hiint(c.t);
        v.z = c.op == RT_NO_HIT ? RT_NO_HIT : (c.op | (c.face << 29));
        v.w = pos | ((c.op == RT_NO_HIT ? 0u : c.cls) << 29); /* shade class of what was hit: k_classify needs no other load */
#if RT_STREAM_HINTS
        __stcs(reinterpret_cast<uint4 *>(hb.rec + idx), v);
#else
        *reinterpret_cast<uint4 *>(hb.rec + idx) = v;
#endif
    }
};
/* NEST 0: every ConstantMedium of the scene is bounded by one Sphere (intersected in line); NEST 1: some boundary has to be walked,
 * and the warp loop does it itself (Traversal<.., WALK_IN_LOOP>, any depth up to RT_MAX_MEDIUM_NESTING).  Picked at rtgpu_scene_create. */
template <int NEST, bool MESH>
__global__ void __launch_bounds__(kThreads, NEST == 0 ? RT_EXTEND_MIN_BLOCKS0 : RT_EXTEND_MIN_BLOCKS) k_extend(DScene sc, PathQueue q, HitBuf hb, WaveCtl *ctl, int cur,
                                                                           WaveParams wp, const uint32_t *perm, uint32_t n_perm) {
    QueueIO io;
    io.q = q;
    io.hb = hb;
    io.wp = wp;
    io.perm = perm;
    io.n_perm = perm ? n_perm : 0u;
    extend_warp_loop<NEST, MESH>(trav_of(sc), sc.n_units, ctl->count[cur], &ctl->fetch, io);
}

/* ---- ray reordering between bounces ---------------------------------------------------------------
 * Rays that start close together and point the same way walk nearly the same ops in the same order (the traversal
 * order is ray-independent), so a warp of neighbours in a sorted order diverges far less: the box rounds of k_extend
 * run with 5-6 of 32 lanes on bounce rays in arrival order and ~16 on camera rays.  The records do not move:
 *   - the shade kernels write a 32-bit key next to every ray they enqueue (ray_sort_key: origin cell, Morton order,
 *     then direction octant and a 2 x 3-bit direction) -- no extra pass over the queue;
 *   - cub::DeviceRadixSort sorts (key, entry index) pairs;
 *   - k_extend takes its rays through the sorted index (QueueIO::perm) and writes its results to the entry's own
 *     slot, so classify / shade are unaffected and the image does not depend on the order at all.
 * The grid of origin cells is fixed per render from the first hit points (mean +- 2.5 sigma per axis: the global fog
 * scatters a few rays very far away; min/max bounds would waste the key bits on empty space). */
struct SortGrid {
    float lo[3], inv[3];
};
#ifndef RT_SORT_CELL_BITS
#define RT_SORT_CELL_BITS 7 /* origin cells per axis = 2^bits */
#endif
#ifndef RT_SORT_SIGMAS
#define RT_SORT_SIGMAS 2.5 /* half-width of the cell grid in standard deviations of the first hit points */
#endif
#ifndef RT_SORT_DIR_BITS
#define RT_SORT_DIR_BITS 8 /* direction below the octant: 8, 10, 12 = major axis + 2 x (3, 4, 5) bits, 6 = 2 x 3 bits, 0 = octant only */
#endif
constexpr int kSortKeyBits = 3 * RT_SORT_CELL_BITS + 3 + RT_SORT_DIR_BITS; /* radix passes = ceil(bits / 8) */
__device__ __forceinline__ uint32_t spread_bits(uint32_t v) { /* the low (up to 10) bits of v -> every third bit */
    static_assert(RT_SORT_CELL_BITS <= 10, "spread_bits spreads at most 10 bits");
    v &= (1u << RT_SORT_CELL_BITS) - 1u;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ __forceinline__ uint32_t ray_sort_key(const SortGrid &g, V3 ro, V3 rd) {
    float o[3] = {(float)ro.x, (float)ro.y, (float)ro.z};
    float d[3] = {(float)rd.x, (float)rd.y, (float)rd.z};
    uint32_t cell[3];
    const float top = (float)((1 << RT_SORT_CELL_BITS) - 1);
    for (int c = 0; c < 3; ++c) cell[c] = (uint32_t)fminf(fmaxf((o[c] - g.lo[c]) * g.inv[c], 0.f), top); /* NaN -> 0 */
    uint32_t morton = spread_bits(cell[0]) | (spread_bits(cell[1]) << 1) | (spread_bits(cell[2]) << 2);
    uint32_t oct = (d[0] < 0.f ? 1u : 0u) | (d[1] < 0.f ? 2u : 0u) | (d[2] < 0.f ? 4u : 0u);
    float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
    float m = fmaxf(ax, fmaxf(ay, az));
    constexpr int kFine = RT_SORT_DIR_BITS >= 8 ? (RT_SORT_DIR_BITS - 2) / 2 : 3; /* bits per minor direction component */
    constexpr float kTop = (float)((1 << kFine) - 1);
    float sc = m > 0.f ? (kTop + 0.999f) / m : 0.f;
    uint32_t major = (ax >= ay && ax >= az) ? 0u : (ay >= az ? 1u : 2u);
    float u = major == 0 ? ay : ax, v = major == 2 ? ay : az;
    uint32_t dq = 0;
    if (RT_SORT_DIR_BITS >= 8) dq = (major << (2 * kFine)) | ((uint32_t)fminf(u * sc, kTop) << kFine) | (uint32_t)fminf(v * sc, kTop);
    else if (RT_SORT_DIR_BITS == 6) dq = ((uint32_t)fminf(ax * sc, 7.f) << 3) | (uint32_t)fminf(ay * sc, 7.f);
    return (((morton << 3) | oct) << RT_SORT_DIR_BITS) | dq;
}
/* moments of the first hit points of a render (every 8th ray of the first iteration: camera rays) */
__global__ void k_sort_moments(PathQueue q, HitBuf hb, const WaveCtl *ctl, int cur, double *moments /* sum xyz, sum xyz^2, n */) {
    uint32_t n = ctl->count[cur];
    double s1[3] = {0., 0., 0.}, s2[3] = {0., 0., 0.}, cnt = 0.;
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 8u; i < n; i += stride * 8u) {
        HitRec16 hr = ld_hit(hb.rec + i);
        if (hr.op == RT_NO_HIT) continue;
        RayRec rr = ld_ray(q.ray + i);
        double t = hr.t;
        double x = rr.ox + t * rr.dx, y = rr.oy + t * rr.dy, z = rr.oz + t * rr.dz;
        if (fabs(x) < 1e30 && fabs(y) < 1e30 && fabs(z) < 1e30) {
            s1[0] += x; s1[1] += y; s1[2] += z;
            s2[0] += x * x; s2[1] += y * y; s2[2] += z * z;
            cnt += 1.;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        for (int c = 0; c < 3; ++c) {
            s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o);
            s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], o);
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.) {
        for (int c = 0; c < 3; ++c) {
            atomicAdd(moments + c, s1[c]);
            atomicAdd(moments + 3 + c, s2[c]);
        }
        atomicAdd(moments + 6, cnt);
    }
}
__global__ void k_sort_grid(const double *moments, SortGrid *g) {
    for (int c = 0; c < 3; ++c) {
        double cnt = fmax(moments[6], 1.);
        double mean = moments[c] / cnt;
        double sd = sqrt(fmax(moments[3 + c] / cnt - mean * mean, 0.));
        g->lo[c] = (float)(mean - RT_SORT_SIGMAS * sd);
        float ext = (float)(2. * RT_SORT_SIGMAS * sd);
        g->inv[c] = ext > 0.f ? ((float)(1 << RT_SORT_CELL_BITS) - 0.001f) / ext : 0.f;
    }
}
__global__ void k_iota(uint32_t *p, uint32_t n) {
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = i;
}

/* Path regeneration: while samples remain, the queue of the coming iteration is topped up to `pool` entries with new
 * camera rays (main.rs:145-149 + Camera::get_ray).  Paths are generated in path-id order = sample-major, so the new
 * rays of one launch are neighbouring pixels of the same sample.  k_iter_begin has published the new length already. */
__global__ void __launch_bounds__(kThreads) k_regen(DScene sc, PathQueue q, const WaveCtl *ctl, int cur, WaveParams wp) {
    const uint32_t cnt = ctl->n_old, first = ctl->gen_first;
    const uint32_t add = ctl->count[cur] - cnt;
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < add; k += stride) {
        uint32_t pid = first + k;
        uint32_t pixel = wp.first_pixel + pid % wp.npix_wave;
        rt_rng rng;
        path_rng(rng, wp, pid, 0);
        Ray r = camera_ray(sc.cam, (int)(pixel % (uint32_t)wp.width), (int)(pixel / (uint32_t)wp.width), wp.width, wp.height, rng);
        st_ray(q.ray + cnt + k, r.o, r.d, r.tm, pid, rng.pos | (wp.max_depth << RT_POS_BITS));
        st_beta(q.st + cnt + k, v3(1., 1., 1.));
    }
}

/* ---- shading, split by material class ------------------------------------------------------------
 * One monolithic shade kernel was 130 KB of SASS executed by warps whose lanes sat on five different
 * materials: ncu showed 6.7 active lanes per instruction and `no_instruction` (instruction-cache miss)
 * as the top stall.  Hits are therefore binned by material class first (k_classify: one warp-aggregated
 * atomic per class), and each class runs its own small kernel over its own index list, so warps are
 * uniform and every kernel fits the instruction cache.
 *   class 0  path ends here: miss (background, main.rs:275-277) or DiffuseLight (main.rs:272-274)
 *   class 1  Lambertian over SolidColor / Checker-of-solids      class 2  Lambertian over any texture
 *   class 3  Metal            class 4  Dielectric                class 5  Isotropic (ConstantMedium) */
constexpr int kClasses = 6;

/* A block bins a tile of 256 x 8 hits: warp-aggregated shared-memory counters give every hit its rank inside the tile, then ONE global
 * atomic per class and tile reserves the tile's ranges of the class lists.  (One global atomic per warp and class -- 1.2 M of them on
 * six addresses per iteration -- kept this kernel at 0.3 ms for 0.3 GB of traffic: same-address atomics retire ~1 per L2 clock.) */
constexpr int kClassifyThreads = 256, kClassifyPerThread = 8;
__global__ void __launch_bounds__(kClassifyThreads) k_classify(HitBuf hb, WaveCtl *ctl, int cur, uint32_t *class_idx, uint32_t cap) {
    __shared__ uint32_t s_cnt[8], s_base[8];
    const uint32_t n = ctl->count[cur];
    const uint32_t lane = threadIdx.x & 31u;
    constexpr uint32_t kTile = kClassifyThreads * kClassifyPerThread;
    for (uint32_t tile0 = blockIdx.x * kTile; tile0 < n; tile0 += gridDim.x * kTile) {
        if (threadIdx.x < 8) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t cls[kClassifyPerThread], rank[kClassifyPerThread];
#pragma unroll
        for (int k = 0; k < kClassifyPerThread; ++k) {
            const uint32_t i = tile0 + k * kClassifyThreads + threadIdx.x;
            cls[k] = i < n ? hb.rec[i].rng >> 29 : 0xffu; /* written by k_extend: 0 for a miss, else the class of the material that was hit */
        }
#pragma unroll
        for (int k = 0; k < kClassifyPerThread; ++k) {
            const uint32_t peers = __match_any_sync(0xffffffffu, cls[k]);
            rank[k] = 0;
            if (cls[k] != 0xffu) {
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if ((int)lane == leader) base = atomicAdd(&s_cnt[cls[k]], (uint32_t)__popc(peers));
                rank[k] = __shfl_sync(peers, base, leader) + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            }
        }
        __syncthreads();
        if (threadIdx.x < 8) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(ctl->class_count + threadIdx.x, s_cnt[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kClassifyPerThread; ++k)
            if (cls[k] != 0xffu) class_idx[(size_t)cls[k] * cap + s_base[cls[k]] + rank[k]] = tile0 + k * kClassifyThreads + threadIdx.x;
    }
}

/* the body of ray_color (main.rs:233-278) for one segment of a path whose hit is of class CLS;
 * returns true when the path goes on */
template <int CLS>
__device__ __forceinline__ bool shade_segment(const DScene &sc, const Ray &r, double t, uint32_t hop, rt_rng &rng, V3 &beta,
                                              Ray &next, V3 &contrib) {
    if constexpr (CLS == 0) {
        if (hop == RT_NO_HIT) { /* main.rs:275-277 */
            contrib = beta * v3(sc.background[0], sc.background[1], sc.background[2]);
            return false;
        }
        /* DiffuseLight: emitted, no scatter (material/mod.rs:173-181; main.rs:272-274) */
        HitRec rec;
        make_record<true>(sc, r, t, hop & 0x1fffffffu, hop >> 29, rec);
        const rtgpu_material &m = sc.mats[rec.mat];
        V3 e = rec.front ? texture_value<RT_TEXMASK_ALL>(tex_of(sc), m.tex, rec.u, rec.v, rec.p) : v3(0., 0., 0.);
        contrib = beta * e;
        return false;
    } else {
        HitRec rec;
        make_record<CLS == 2 || CLS == 5>(sc, r, t, hop & 0x1fffffffu, hop >> 29, rec);
        const rtgpu_material &m = sc.mats[rec.mat];
        if constexpr (CLS == 3) { /* Metal, material/mod.rs:85-96: time reset to 0, no absorption test (Q6) */
            V3 reflected = reflect(to_unit(r.d), rec.n);
            next.o = rec.p;
            next.d = reflected + random_in_unit_sphere(rng) * m.p[3];
            next.tm = 0.;
            beta = beta * v3(m.p[0], m.p[1], m.p[2]);
            return true;
        } else if constexpr (CLS == 4) { /* Dielectric, material/mod.rs:120-147 */
            double refraction_ratio = rec.front ? 1. / m.p[0] : m.p[0];
            V3 unit_direction = to_unit(r.d);
            double cos_theta = fmin(dot(-unit_direction, rec.n), 1.);
            double sin_theta = sqrt(1. - cos_theta * cos_theta);
            bool cannot_refract = refraction_ratio * sin_theta > 1.;
            double random_double = rng.gen_range(0.0, 1.0);
            V3 direction = (cannot_refract || reflectance(cos_theta, refraction_ratio) > random_double)
                               ? reflect(unit_direction, rec.n)
                               : refract(unit_direction, rec.n, refraction_ratio);
            next.o = rec.p;
            next.d = direction;
            next.tm = r.tm;
            return true; /* attenuation (1,1,1) */
        } else if constexpr (CLS == 5) { /* Isotropic, material/mod.rs:206-214 */
            next.o = rec.p;
            next.d = random_in_unit_sphere(rng);
            next.tm = r.tm;
            beta = beta * texture_value<RT_TEXMASK_ALL>(tex_of(sc), m.tex, rec.u, rec.v, rec.p);
            return true;
        } else { /* Lambertian + MixturePdf(HittablePdf(lights), CosPdf), main.rs:263-271 */
            V3 attenuation = texture_value<(CLS == 1 ? RT_TEXMASK_SIMPLE : RT_TEXMASK_ALL)>(tex_of(sc), m.tex, rec.u, rec.v, rec.p);
            Onb uvw = onb_from_w(rec.n); /* CosPdf::new(rec.normal), pdf.rs:35-39 */
            V3 direction;
            double pdf_val;
            V3 unit_direction; /* direction.to_unit(): upstream forms it in CosPdf::value and again in scattering_pdf -- the same value */
            if (sc.n_light_roots == 0) {
                /* Q12: empty `lights` -> the cosine pdf alone (decision shared with the oracle) */
                direction = onb_local(uvw, random_cosine_direction(rng));
                unit_direction = to_unit(direction);
                double c = dot(unit_direction, uvw.w);
                pdf_val = c <= 0. ? 0. : c / RT_PI;
            } else {
                if (rng.gen_range(0.0, 1.0) < 0.5) /* MixturePdf::generate, pdf.rs:95-101 */
                    direction = lights_random(lights_of(sc), rec.p, rng);
                else
                    direction = onb_local(uvw, random_cosine_direction(rng));
                unit_direction = to_unit(direction);
                double c = dot(unit_direction, uvw.w);
                double cos_pdf = c <= 0. ? 0. : c / RT_PI;                            /* CosPdf::value, pdf.rs:46-53 */
                pdf_val = 0.5 * lights_pdf_value(lights_of(sc), rec.p, direction) + 0.5 * cos_pdf; /* pdf.rs:102-104 */
            }
            double cosine = dot(rec.n, unit_direction); /* Lambertian::scattering_pdf, material/mod.rs:58-65 */
            double spdf = cosine < 0. ? 0. : cosine / RT_PI;
            next.o = rec.p;
            next.d = direction;
            next.tm = r.tm;
            beta = beta * (attenuation * spdf / pdf_val); /* no pdf == 0 guard upstream either (Q10) */
            return true;
        }
    }
}

/* where the radiance of a finished path goes */
struct PathSink {
    long long *fix;     /* per pixel and channel: sum of contributions in 2^-32 units (exact, order-independent) */
    double *big;        /* per pixel and channel: contributions too large for the fixed-point sum (|c| >= 2^16) */
    uint32_t *flags;    /* per pixel: bit c = NaN, bit 3+c = +inf, bit 6+c = -inf seen in channel c */
    double *cx, *cy, *cz; /* per-path mode (rtgpu_sample_radiance): radiance of path pid, f64 */
    uint32_t *path_len;   /* per-path mode: segments traced */
};
#define RT_FIX_SCALE 4294967296.0 /* 2^32 */
#define RT_FIX_LIMIT 65536.0      /* |c| below this goes to the fixed-point sum: 32768 samples cannot overflow 63 bits */
__device__ __forceinline__ void sink_add(const PathSink &sk, uint32_t pixel, int ch, double c) {
    if (c == 0.) return;
    if (c != c) {
        atomicOr(sk.flags + pixel, 1u << ch);
    } else if (fabs(c) < RT_FIX_LIMIT) {
        atomicAdd(reinterpret_cast<unsigned long long *>(sk.fix + 3 * (size_t)pixel + ch), (unsigned long long)__double2ll_rn(c * RT_FIX_SCALE));
    } else if (fabs(c) <= 1.7976931348623157e308) {
        atomicAdd(sk.big + 3 * (size_t)pixel + ch, c);
    } else {
        atomicOr(sk.flags + pixel, (c > 0. ? 8u : 64u) << ch);
    }
}

template <int CLS>
#ifndef RT_SHADE_MIN_BLOCKS
#define RT_SHADE_MIN_BLOCKS RT_SHADE_BLOCKS_PER_SM
#endif
__global__ void __launch_bounds__(kShadeThreads, RT_SHADE_MIN_BLOCKS) k_shade(DScene sc, PathQueue qin, HitBuf hb, PathQueue qout, WaveCtl *ctl, int cur,
                                                    const uint32_t *class_idx, PathSink sk, WaveParams wp, const SortGrid *grid,
                                                    uint32_t *keys_out) {
    __shared__ uint32_t s_len[RTGPU_MAX_BOUNCE_STATS + 1]; /* paths by length, flushed once per block */
    for (uint32_t k = threadIdx.x; k <= RTGPU_MAX_BOUNCE_STATS; k += blockDim.x) s_len[k] = 0;
    __syncthreads();
    uint32_t n = ctl->class_count[CLS];
    uint32_t stride = gridDim.x * blockDim.x;
    uint32_t lane = threadIdx.x & 31u;
    for (uint32_t j0 = (blockIdx.x * blockDim.x + threadIdx.x) - lane; j0 < n; j0 += stride) {
        uint32_t j = j0 + lane;
        bool alive = false, ended = false;
        Ray next;
        V3 beta;
        uint32_t pid = 0, pos = 0, left = 0, traced = 0;
        if (j < n) {
            uint32_t i = class_idx[j];
            Ray r;
            RayRec rr = ld_ray(qin.ray + i);
            HitRec16 hr = ld_hit(hb.rec + i);
            r.o = v3(rr.ox, rr.oy, rr.oz);
            r.d = v3(rr.dx, rr.dy, rr.dz);
            r.tm = rr.tm;
            beta = ld_beta(qin.st + i);
            pid = rr.pid;
            left = rr.rng >> RT_POS_BITS; /* segments this path may still trace, this one included */
            rt_rng rng;
            path_rng(rng, wp, pid, hr.rng & RT_POS_MASK);
            V3 contrib = v3(0., 0., 0.);
            alive = shade_segment<CLS>(sc, r, hr.t, hr.op, rng, beta, next, contrib);
            pos = rng.pos;
            traced = wp.max_depth - left + 1u;
            if (alive && left <= 1u) { /* depth exhausted: the next ray_color call returns black (main.rs:240-242) */
                alive = false;
                contrib = v3(0., 0., 0.);
            }
            if (!alive) {
                ended = true;
                if (sk.cx) {
                    sk.cx[pid] = contrib.x;
                    sk.cy[pid] = contrib.y;
                    sk.cz[pid] = contrib.z;
                    if (sk.path_len) sk.path_len[pid] = traced;
                } else {
                    uint32_t pixel = pid % wp.npix_wave;
                    sink_add(sk, pixel, 0, contrib.x);
                    sink_add(sk, pixel, 1, contrib.y);
                    sink_add(sk, pixel, 2, contrib.z);
                }
            }
        }
        /* paths by length (rays per bounce follow from it): one atomic per distinct length in the warp */
        {
            uint32_t bin = ended ? (traced < RTGPU_MAX_BOUNCE_STATS ? traced : RTGPU_MAX_BOUNCE_STATS) : 0xffffffffu;
            uint32_t peers = __match_any_sync(0xffffffffu, bin);
            if (ended && (int)lane == __ffs(peers) - 1) atomicAdd(&s_len[bin], (uint32_t)__popc(peers));
        }
        if constexpr (CLS != 0) {
            /* compaction: one atomic per warp, survivors packed in lane order */
            uint32_t mask = __ballot_sync(0xffffffffu, alive);
            if (mask) {
                uint32_t base = 0;
                int leader = __ffs(mask) - 1;
                if ((int)lane == leader) base = atomicAdd(&ctl->count[cur ^ 1], (uint32_t)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (alive) {
                    uint32_t k = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
                    st_ray(qout.ray + k, next.o, next.d, next.tm, pid, pos | ((left - 1u) << RT_POS_BITS));
                    st_beta(qout.st + k, beta);
                    if (keys_out) keys_out[k] = ray_sort_key(*grid, next.o, next.d);
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k <= RTGPU_MAX_BOUNCE_STATS; k += blockDim.x)
        if (s_len[k]) atomicAdd(&ctl->len_hist[k], (unsigned long long)s_len[k]);
}

/* pixel_color of main.rs:144-151 for the samples of this call: fixed-point sum + large terms, NaN / inf as flagged;
 * written to (or added to) the caller's float buffer, accumulators cleared for the next call */
__global__ void k_finish(PathSink sk, float *accum, uint32_t npix, int add) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    uint32_t fl = sk.flags[p];
    sk.flags[p] = 0;
    for (int c = 0; c < 3; ++c) {
        size_t k = 3 * (size_t)p + c;
        double v = (double)sk.fix[k] * (1.0 / RT_FIX_SCALE) + sk.big[k];
        sk.fix[k] = 0;
        sk.big[k] = 0.;
        bool pinf = (fl >> (3 + c)) & 1u, ninf = (fl >> (6 + c)) & 1u;
        if (((fl >> c) & 1u) || (pinf && ninf)) v = CUDART_NAN;
        else if (pinf) v = CUDART_INF;
        else if (ninf) v = -CUDART_INF;
        float *a = accum + k;
        *a = add ? *a + (float)v : (float)v;
    }
}

/* The tail of a render: once every path has been generated and only a few thousand are still alive, another 30 - 45 iterations
 * of ~12 launches each would carry a few thousand rays apiece (0.2 - 0.3 ms per iteration for nothing).  k_tail finishes those
 * paths in ONE launch: a thread takes a queue entry and runs its path to the end -- the single-ray walk of the traversal program
 * (the same Traversal::step the medium boundaries use) and the same shade_segment<class>() as the class kernels, so every path
 * gets bit for bit what the iterations would have given it. */
template <int NEST>
__global__ void __launch_bounds__(kThreads) k_tail(DScene sc, PathQueue q, WaveCtl *ctl, int cur, PathSink sk, WaveParams wp) {
    const uint32_t n = ctl->count[cur];
    const uint32_t stride = gridDim.x * blockDim.x;
    const TravScene ts = trav_of(sc);
    unsigned long long my_rays = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        RayRec rr = ld_ray(q.ray + i);
        Ray r;
        r.o = v3(rr.ox, rr.oy, rr.oz);
        r.d = v3(rr.dx, rr.dy, rr.dz);
        r.tm = rr.tm;
        V3 beta = ld_beta(q.st + i);
        const uint32_t pid = rr.pid;
        uint32_t left = rr.rng >> RT_POS_BITS, pos = rr.rng & RT_POS_MASK;
        V3 contrib = v3(0., 0., 0.);
        for (;;) {
            /* world.hit(r, 0.001, f64::MAX), main.rs:243 */
            Traversal<RayRegs> tr;
            SavedRays stk;
            tr.start(ts, 0, r.o, r.d, r.tm, 0.001, RT_F64_MAX);
            StreamDraw draw{wp.seed, wp.first_pixel + pid % wp.npix_wave, wp.sample_begin + pid / wp.npix_wave, RT_DOMAIN_PATH, pos};
            while (!tr.done(ts)) tr.step<0, NEST>(ts, stk, draw);
            pos = draw.pos;
            ++my_rays;
            const uint32_t hop = tr.best.op == RT_NO_HIT ? RT_NO_HIT : (tr.best.op | (tr.best.face << 29));
            const uint32_t cls = tr.best.op == RT_NO_HIT ? 0u : tr.best.cls;
            rt_rng rng;
            path_rng(rng, wp, pid, pos);
            Ray next;
            bool alive;
            contrib = v3(0., 0., 0.);
            switch (cls) {
            case 0: alive = shade_segment<0>(sc, r, tr.best.t, hop, rng, beta, next, contrib); break;
            case 1: alive = shade_segment<1>(sc, r, tr.best.t, hop, rng, beta, next, contrib); break;
            case 2: alive = shade_segment<2>(sc, r, tr.best.t, hop, rng, beta, next, contrib); break;
            case 3: alive = shade_segment<3>(sc, r, tr.best.t, hop, rng, beta, next, contrib); break;
            case 4: alive = shade_segment<4>(sc, r, tr.best.t, hop, rng, beta, next, contrib); break;
            default: alive = shade_segment<5>(sc, r, tr.best.t, hop, rng, beta, next, contrib); break;
            }
            pos = rng.pos;
            if (alive && left <= 1u) { /* depth exhausted: the next ray_color call returns black (main.rs:240-242) */
                alive = false;
                contrib = v3(0., 0., 0.);
            }
            if (!alive) break;
            r = next;
            --left;
        }
        const uint32_t traced = wp.max_depth - left + 1u;
        if (sk.cx) {
            sk.cx[pid] = contrib.x;
            sk.cy[pid] = contrib.y;
            sk.cz[pid] = contrib.z;
            if (sk.path_len) sk.path_len[pid] = traced;
        } else {
            uint32_t pixel = pid % wp.npix_wave;
            sink_add(sk, pixel, 0, contrib.x);
            sink_add(sk, pixel, 1, contrib.y);
            sink_add(sk, pixel, 2, contrib.z);
        }
        atomicAdd(&ctl->len_hist[traced < RTGPU_MAX_BOUNCE_STATS ? traced : RTGPU_MAX_BOUNCE_STATS], 1ull);
    }
    /* rays traced here (the iterations count theirs in k_iter_begin) */
    for (int o = 16; o > 0; o >>= 1) my_rays += __shfl_xor_sync(0xffffffffu, my_rays, o);
    if ((threadIdx.x & 31) == 0 && my_rays) atomicAdd(&ctl->rays, my_rays);
}
__global__ void k_tail_done(WaveCtl *ctl, int cur) { ctl->count[cur] = 0; }

/* dumped-ray batch: BvhNode::hit through the SAME warp loop as k_extend, then the record of the winner */
struct BatchIO {
    static constexpr bool kConstTmin = false; /* a dumped ray brings its own t_min */
    const rtgpu_ray *rays;
    const uint64_t *keys;
    double *t;
    uint32_t *op, *rng_pos;
    __device__ __forceinline__ void load(uint32_t idx, V3 &o, V3 &d, double &tm, double &tmin, double &tmax, uint32_t &a, uint32_t &b,
                                         uint32_t &pos, uint32_t &cur) {
        rtgpu_ray q = rays[idx];
        o = v3(q.o[0], q.o[1], q.o[2]);
        d = v3(q.d[0], q.d[1], q.d[2]);
        tm = q.tm;
        tmin = q.t_min;
        tmax = q.t_max;
        uint64_t key = keys ? keys[idx] : (uint64_t)idx; /* stream BATCH(key) */
        a = (uint32_t)key;
        b = (uint32_t)(key >> 32);
        pos = 0;
        cur = idx;
    }
    __device__ __forceinline__ int refill_at() const { return RT_EXTEND_REFILL; }
    __device__ __forceinline__ void prefetch(uint32_t idx) { asm volatile("prefetch.global.L2 [%0];" ::"l"(rays + idx)); }
    __device__ __forceinline__ StreamDraw stream(uint32_t a, uint32_t b, uint32_t pos) const {
        return StreamDraw{((uint64_t)b << 32) | (uint64_t)a, 0u, 0u, RT_DOMAIN_BATCH, pos};
    }
    __device__ __forceinline__ double draw(uint32_t a, uint32_t b, uint32_t &pos) const {
        rt_rng rng;
        rng.init(((uint64_t)b << 32) | (uint64_t)a, 0, 0, RT_DOMAIN_BATCH, pos);
        double u = rng.gen();
        pos = rng.pos;
        return u;
    }
    __device__ __forceinline__ void store(uint32_t cur, const Closest &c, uint32_t pos) {
        t[cur] = c.t;
        op[cur] = c.op == RT_NO_HIT ? RT_NO_HIT : (c.op | (c.face << 29));
        rng_pos[cur] = pos;
    }
};
template <int NEST>
__global__ void __launch_bounds__(kThreads, NEST == 0 ? RT_EXTEND_MIN_BLOCKS0 : RT_EXTEND_MIN_BLOCKS) k_trace_extend(DScene sc, const rtgpu_ray *rays, const uint64_t *keys, double *t,
                                                                   uint32_t *op, uint32_t *rng_pos, uint32_t n, uint32_t *fetch) {
    BatchIO io;
    io.rays = rays;
    io.keys = keys;
    io.t = t;
    io.op = op;
    io.rng_pos = rng_pos;
    extend_warp_loop<NEST, true>(trav_of(sc), sc.n_units, n, fetch, io);
}
__global__ void __launch_bounds__(kThreads) k_trace_record(DScene sc, const rtgpu_ray *rays, const double *t, const uint32_t *op,
                                                           const uint32_t *rng_pos, rtgpu_hit *out, uint32_t n) {
    uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        rtgpu_ray q = rays[i];
        Ray r;
        r.o = v3(q.o[0], q.o[1], q.o[2]);
        r.d = v3(q.d[0], q.d[1], q.d[2]);
        r.tm = q.tm;
        rtgpu_hit h;
        memset(&h, 0, sizeof h);
        h.gid = -1;
        h.inst = -1;
        h.mat = -1;
        h.rng_pos = rng_pos[i];
        if (op[i] != RT_NO_HIT) {
            HitRec rec;
            make_record<true>(sc, r, t[i], op[i] & 0x1fffffffu, op[i] >> 29, rec);
            h.hit = 1;
            h.t = rec.t;
            h.p[0] = rec.p.x; h.p[1] = rec.p.y; h.p[2] = rec.p.z;
            h.n[0] = rec.n.x; h.n[1] = rec.n.y; h.n[2] = rec.n.z;
            h.u = rec.u;
            h.v = rec.v;
            h.front_face = rec.front ? 1 : 0;
            h.gid = rec.gid;
            h.face = rec.face;
            h.inst = rec.inst;
            h.mat = (int32_t)rec.mat;
        }
        out[i] = h;
    }
}

__global__ void k_camera_rays(DScene sc, WaveParams wp, rtgpu_ray *out, uint32_t *rng_pos) {
    uint32_t pid = blockIdx.x * blockDim.x + threadIdx.x;
    if (pid >= wp.n_paths) return;
    uint32_t pixel = wp.first_pixel + pid % wp.npix_wave;
    rt_rng rng;
    path_rng(rng, wp, pid, 0);
    Ray r = camera_ray(sc.cam, (int)(pixel % (uint32_t)wp.width), (int)(pixel / (uint32_t)wp.width), wp.width, wp.height, rng);
    rtgpu_ray q;
    q.o[0] = r.o.x; q.o[1] = r.o.y; q.o[2] = r.o.z;
    q.d[0] = r.d.x; q.d[1] = r.d.y; q.d[2] = r.d.z;
    q.tm = r.tm;
    q.t_min = 0.001;
    q.t_max = RT_F64_MAX;
    out[pid] = q;
    if (rng_pos) rng_pos[pid] = rng.pos;
}

/* write_color (main.rs:280-299) + the row flip of main.rs:197 */
__global__ void k_resolve(const float *accum, int w, int h, int spp, uint8_t *rgb8) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= w * h) return;
    int x = p % w, y = p / w;
    uint8_t *o = rgb8 + 3 * ((size_t)(h - y - 1) * w + x);
    for (int c = 0; c < 3; ++c) {
        double r = (double)accum[3 * (size_t)p + c];
        if (r != r) r = 0.;
        double v = sqrt(r / (double)spp);
        if (v < 0.0) v = 0.0;
        if (v > 0.999) v = 0.999; /* f64::clamp keeps NaN; `as u8` maps it to 0 */
        double f = floor(v * 255.999);
        o[c] = (f != f) ? 0 : (f <= 0. ? 0 : (f >= 255. ? 255 : (uint8_t)f));
    }
}

} // namespace

/* --------------------------------------------------------------------------------------------- */
/* per-device working memory of the wavefront: queues, hit buffer, contributions, counters.  It is
 * independent of the scene, so a destroyed scene parks it for the next scene on the same device
 * (an application that re-creates its scene every frame does not pay cudaMalloc of GBs each time). */
struct CtlSnapshot { /* what the host reads back, asynchronously, once per iteration */
    uint32_t count[2], gen_next, fetch;
};
constexpr int kSnapRing = 64;
constexpr int kSideStreams = 5;
constexpr int kRunAheadDefault = 2;
struct WaveRes {
    size_t cap = 0; /* paths the pool can hold */
    void *wave_mem = nullptr;
    PathQueue q[2];
    HitBuf hb;
    uint32_t *class_idx = nullptr; /* kClasses lists of queue indices, `cap` entries each */
    /* ray reordering: keys written by k_regen / the shade kernels, their sorted copy, iota and the sorted entry indices */
    uint32_t *sort_keys[2] = {nullptr, nullptr}, *sort_idx[2] = {nullptr, nullptr};
    double *sort_bounds = nullptr; /* 7 moments, then the SortGrid */
    bool iota_ready = false;
    void *sort_temp = nullptr;
    size_t sort_temp_bytes = 0;
    WaveCtl *ctl = nullptr;           /* device */
    CtlSnapshot *snap_host = nullptr; /* pinned ring */
    WaveCtl *ctl_host = nullptr;      /* pinned: final read of the counters */
    cudaEvent_t snap_ev[kSnapRing] = {};
    /* pixel sums (PathSink) */
    void *sink_mem = nullptr;
    size_t sink_pix = 0;
    long long *fix = nullptr;
    double *big = nullptr;
    uint32_t *flags = nullptr;
    /* per-path outputs of rtgpu_sample_radiance */
    void *perpath_mem = nullptr;
    size_t perpath_cap = 0;
    double *cx = nullptr, *cy = nullptr, *cz = nullptr;
    uint32_t *path_len = nullptr;
    float *accum_dev = nullptr;
    size_t accum_cap = 0;
    std::vector<cudaEvent_t> events;
    /* side streams: kernels of an iteration that do not depend on each other (k_regen beside the sort, the class kernels beside each
     * other) are issued on these so that the tail of one fills with the blocks of the next; joined back before anything depends on them */
    cudaStream_t side[kSideStreams] = {};
    cudaEvent_t fork_ev = nullptr, side_ev[kSideStreams] = {};
    void release() {
        if (wave_mem) cudaFree(wave_mem);
        if (sink_mem) cudaFree(sink_mem);
        if (perpath_mem) cudaFree(perpath_mem);
        if (ctl) cudaFree(ctl);
        if (snap_host) cudaFreeHost(snap_host);
        if (ctl_host) cudaFreeHost(ctl_host);
        if (accum_dev) cudaFree(accum_dev);
        if (sort_temp) cudaFree(sort_temp);
        for (cudaEvent_t e : snap_ev)
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        for (cudaStream_t st : side)
            if (st) cudaStreamDestroy(st);
        for (cudaEvent_t e : side_ev)
            if (e) cudaEventDestroy(e);
        if (fork_ev) cudaEventDestroy(fork_ev);
        *this = WaveRes();
    }
};

struct rtgpu_scene : WaveRes {
    int device = 0;
    int n_sm = 148; /* multiprocessors of the device (cudaDevAttrMultiProcessorCount) */
    DScene d;
    std::vector<void *> allocs;
    bool class_present[8] = {true, false, false, false, false, false, false, false};
    int medium_nesting = 0; /* deepest ConstantMedium-inside-a-boundary nesting of the program: picks the NEST of the walk kernels */
    bool has_mesh = true;   /* the program holds Triangle or Ring ops: picks the MESH build of k_extend */
    int overlap = 1; /* RTGPU_OVERLAP=0: everything on one stream */
    int run_ahead = kRunAheadDefault; /* RTGPU_RUN_AHEAD: iterations the host may be ahead of the device (1 .. 32) */
    uint32_t tail_max = 1u << 16; /* RTGPU_TAIL_MAX: at most this many live paths are finished by k_tail in one launch (0 = never) */
    /* Reordering the rays between bounces pays where a ray meets many ops: C3 (2 961 ops) 99 -> 67 ms in k_extend for 10 ms of sorting;
     * on the small programs the sort costs more than the walk gains (B200, same build: C2 Cornell, 13 ops, 13.8 ms sorted / 11.4 not;
     * C1 book-1 scene, 1 506 ops, 15.2 / 14.3; C4 OBJ mesh, 42 556 ops, 28.8 / 28.3).  RTGPU_SORT_RAYS=0 / 1 forces it off / on;
     * unset, programs of at least RTGPU_SORT_MIN_OPS ops are sorted. */
    int sort_mode = 1;
    uint32_t sort_min_ops = 2048;
    uint32_t sort_min = 1u << 20; /* RTGPU_SORT_MIN: smallest bounce worth a sort (A/B on the B200: 64 K .. 1 M within 1 % on C3, 1 M best on C1, C2, C4) */
};

namespace {
constexpr int kMaxDevices = 64;
std::mutex g_park_mutex;
WaveRes g_parked[kMaxDevices];
bool g_parked_valid[kMaxDevices];
} // namespace

namespace {

/* Scene arrays are staged into ONE host blob and copied with ONE cudaMemcpy into ONE allocation
 * (256-byte aligned sub-arrays): scene creation costs one malloc + one copy instead of eleven. */
struct SceneArena {
    std::vector<unsigned char> host;
    std::vector<std::pair<size_t, const void **>> fix; /* offset -> device pointer to patch */
    template <typename T>
    void add(const T *src, size_t n, const T **dst) {
        size_t off = (host.size() + 255) & ~(size_t)255;
        size_t bytes = (n ? n : 1) * sizeof(T);
        host.resize(off + bytes, 0);
        if (n) memcpy(host.data() + off, src, n * sizeof(T));
        fix.emplace_back(off, reinterpret_cast<const void **>(dst));
    }
    rtgpu_status commit(rtgpu_scene *s);
};

rtgpu_status validate(const rtgpu_scene_desc *d, int *medium_nesting, bool *has_mesh) {
    *medium_nesting = 0;
    *has_mesh = false;
    auto bad = [](const char *m) {
        g_err = m;
        return RTGPU_ERR_INVALID_SCENE;
    };
    if (!d || !d->ops || d->n_ops == 0) return bad("empty traversal program (the reference panics on an empty BvhNode)");
    if (d->n_ops >= 0x1fffffffu) return bad("traversal program too long");
    if (d->n_ctxs == 0) return bad("no contexts");
    for (uint32_t i = 0; i < d->n_ctxs; ++i) {
        if (d->ctxs[i].len > RT_MAX_CHAIN) return bad("more than RT_MAX_CHAIN wrappers around one primitive");
        if ((uint64_t)d->ctxs[i].start + d->ctxs[i].len > d->n_chain) return bad("context outside the chain array");
    }
    for (uint32_t i = 0; i < d->n_chain; ++i)
        if (d->chain[i] >= d->n_xforms) return bad("chain entry outside the xform array");
    int push = 0;
    std::vector<uint32_t> medium_ends; /* skip targets of the ConstantMedium ops whose boundary the scan is inside of */
    for (uint32_t i = 0; i < d->n_ops;) {
        while (!medium_ends.empty() && i >= medium_ends.back()) medium_ends.pop_back();
        const rtgpu_op &op = d->ops[i];
        uint32_t kind = op.kind_ctx & 0xff, ctx = op.kind_ctx >> 8;
        uint32_t width = (kind == RTGPU_OP_MSPHERE || kind == RTGPU_OP_TRI) ? 2 : 1;
        if (i + width > d->n_ops) return bad("two-slot op at the end of the program");
        if (kind == RTGPU_OP_TRI || kind == RTGPU_OP_RING) *has_mesh = true;
        switch (kind) {
        case RTGPU_OP_NODE:
            if (op.skip <= i || op.skip > d->n_ops) return bad("skip index out of range");
            break;
        case RTGPU_OP_MEDIUM:
            if (op.skip <= i || op.skip > d->n_ops) return bad("skip index out of range");
            if (!medium_ends.empty() && op.skip > medium_ends.back()) return bad("ConstantMedium boundaries overlap");
            medium_ends.push_back(op.skip);
            if (medium_ends.size() > RT_MAX_MEDIUM_NESTING) return bad("ConstantMedium boundaries nested deeper than RT_MAX_MEDIUM_NESTING");
            {   /* a top-level medium over ONE Sphere is intersected in line (both media of final_scene): it needs no boundary walk */
                const bool one_sphere = medium_ends.size() == 1 && op.skip == i + 2 && (d->ops[i + 1].kind_ctx & 0xff) == RTGPU_OP_SPHERE;
                const int need = one_sphere ? 0 : (int)medium_ends.size();
                if (need > *medium_nesting) *medium_nesting = need;
            }
            break;
        case RTGPU_OP_PUSH:
            if ((uint64_t)op.a + op.b > d->n_chain) return bad("PUSH chain out of range");
            if (op.skip <= i || op.skip > d->n_ops) return bad("skip index out of range");
            /* the nesting is counted here, not taken from desc->max_push_depth: the walk saves one ray per level */
            if (++push > RT_MAX_PUSH) return bad("wrapper nesting deeper than RT_MAX_PUSH");
            break;
        case RTGPU_OP_POP:
            if (--push < 0) return bad("POP without PUSH");
            break;
        default:
            if (kind > RTGPU_OP_RING) return bad("unknown op kind");
            break;
        }
        if (kind != RTGPU_OP_NODE && kind != RTGPU_OP_PUSH && kind != RTGPU_OP_POP) {
            if (ctx >= d->n_ctxs) return bad("context id out of range");
            if (op.a >= d->n_materials) return bad("material id out of range");
        }
        i += width;
    }
    if (push != 0) return bad("unbalanced PUSH/POP");
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const rtgpu_material &m = d->materials[i];
        if (m.kind > 4) return bad("unknown material kind");
        if ((m.kind == 0 || m.kind == 3 || m.kind == 4) && m.tex >= d->n_textures) return bad("texture id out of range");
    }
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const rtgpu_texture &t = d->textures[i];
        if (t.kind > 3) return bad("unknown texture kind");
        if (t.kind == 1 && (t.a < 0 || t.b < 0 || (uint32_t)t.a >= i || (uint32_t)t.b >= i)) return bad("checker sub-texture must precede it");
        if (t.kind == 2 && (t.a < 0 || (uint32_t)t.a >= d->n_perlins)) return bad("perlin id out of range");
        if (t.kind == 3 && t.a >= (int32_t)d->n_images) return bad("image id out of range");
    }
    for (uint32_t i = 0; i < d->n_images; ++i) {
        const rtgpu_image &im = d->images[i];
        if (im.width <= 0 || im.height <= 0 || im.offset < 0 ||
            (uint64_t)im.offset + (uint64_t)im.width * (uint64_t)im.height * 3 > d->n_texel_bytes)
            return bad("image outside the texel buffer");
    }
    if (d->n_light_roots > d->n_lights || (d->n_lights > 0 && d->n_light_roots == 0)) return bad("n_light_roots out of range");
    {
        /* nesting depth of every entry: roots are level 0; the members of a LIST lie behind it and one level deeper */
        std::vector<uint8_t> level(d->n_lights, 0);
        for (uint32_t i = 0; i < d->n_lights; ++i) {
            const rtgpu_light &l = d->lights[i];
            if (l.kind > RTGPU_LIGHT_LIST) return bad("unknown light kind");
            if (l.kind == RTGPU_LIGHT_LIST) {
                if (!(l.p[0] >= 0. && l.p[1] >= 1. && l.p[0] <= 1048576. && l.p[1] <= 1048576.)) return bad("bad light list range");
                uint32_t first = (uint32_t)l.p[0], count = (uint32_t)l.p[1];
                if ((double)first != l.p[0] || (double)count != l.p[1]) return bad("bad light list range");
                if (first <= i || first < d->n_light_roots || (uint64_t)first + count > d->n_lights) return bad("light list members out of range");
                if (level[i] + 1 > RTGPU_MAX_LIGHT_NESTING) return bad("light lists nested deeper than RTGPU_MAX_LIGHT_NESTING");
                for (uint32_t k = first; k < first + count; ++k) level[k] = (uint8_t)(level[i] + 1);
                continue;
            }
            int np = l.kind == RTGPU_LIGHT_SPHERE ? 4 : (l.kind == RTGPU_LIGHT_DEAD ? 0 : 5);
            for (int k = 0; k < np; ++k)
                if (!(l.p[k] == l.p[k])) return bad("NaN light parameter");
        }
    }
    return RTGPU_OK;
}

/* The caller's program (64-byte ops, include/rtgpu.h) -> the device's 32-byte units (rt_device.cuh): a BvhNode becomes
 * ONE unit with an f32 copy of its box, its f64 box goes to boxes64[ordinal]; every other slot is copied as two units;
 * skip indices are re-mapped; the unused skip word of a primitive gets the shade class of its material.  Two zero units pad the end (the walk always fetches units i and i+1). */
rtgpu_status repack_program(const rtgpu_scene_desc *d, const std::vector<uint8_t> &mat_class, std::vector<uint32_t> &units,
                            std::vector<double> &boxes64, float *extent, uint32_t *n_units) {
    std::vector<uint32_t> at((size_t)d->n_ops + 1);
    uint32_t u = 0, n_nodes = 0;
    for (uint32_t i = 0; i < d->n_ops;) {
        uint32_t kind = d->ops[i].kind_ctx & 0xffu;
        uint32_t width = (kind == RTGPU_OP_MSPHERE || kind == RTGPU_OP_TRI) ? 2 : 1;
        for (uint32_t k = 0; k < width; ++k) at[i + k] = u + 2 * k;
        if (kind == RTGPU_OP_NODE) {
            u += 1;
            ++n_nodes;
        } else {
            u += 2 * width;
        }
        i += width;
    }
    at[d->n_ops] = u;
    if (u >= 0x1fffffffu || n_nodes >= (1u << 24)) {
        g_err = "traversal program too long";
        return RTGPU_ERR_INVALID_SCENE;
    }
    units.assign(((size_t)u + 2) * 8, 0u);
    boxes64.assign((size_t)(n_nodes ? n_nodes : 1) * 6, 0.);
    double ext = (double)RT_CULL_MIN_EXTENT;
    uint32_t ordinal = 0;
    /* A PUSH hands its child a transformed ray, and the boxes under it live in the child's own space: the margin of the box
     * pre-test there only has to cover THOSE boxes (rt_cull.h: B = the largest coordinate of any box the ray is tested against).
     * A mesh modelled at unit size under Zoom / Translate would otherwise be tested with the margin of the whole scene -- hundreds
     * of times its own -- and a quarter of its boxes would fall back to the f64 test.  Each PUSH carries its subtree's extent. */
    struct OpenPush {
        uint32_t end;   /* op index where the wrapper ends */
        double ext;     /* largest |coordinate| of a NODE box seen under it so far */
        uint32_t *unit; /* the PUSH's repacked unit */
    };
    std::vector<OpenPush> open;
    auto close_pushes = [&](uint32_t upto) {
        while (!open.empty() && upto >= open.back().end) {
            OpenPush p = open.back();
            open.pop_back();
            double e = (p.ext == p.ext && p.ext <= 3e38) ? (double)nextafterf((float)p.ext, INFINITY) : (double)INFINITY;
            memcpy(p.unit + 4, &e, 8); /* d[0] of the PUSH op: unused by the op itself */
            if (!open.empty() && !(p.ext <= open.back().ext)) open.back().ext = p.ext; /* (an inner wrapper's boxes are not tested with the outer ray, but keeping the outer bound monotone costs nothing) */
        }
    };
    for (uint32_t i = 0; i < d->n_ops;) {
        close_pushes(i);
        const rtgpu_op &op = d->ops[i];
        uint32_t kind = op.kind_ctx & 0xffu;
        uint32_t width = (kind == RTGPU_OP_MSPHERE || kind == RTGPU_OP_TRI) ? 2 : 1;
        uint32_t *w = units.data() + (size_t)at[i] * 8;
        if (kind == RTGPU_OP_NODE) {
            w[0] = RTGPU_OP_NODE | (ordinal << 8);
            w[1] = at[op.skip];
            for (int k = 0; k < 6; ++k) {
                float f = (float)op.d[k]; /* nearest: rt_cull.h's bound assumes it */
                memcpy(w + 2 + k, &f, 4);
                boxes64[(size_t)ordinal * 6 + k] = op.d[k];
                double a = op.d[k] < 0. ? -op.d[k] : op.d[k];
                if (!(a <= ext)) ext = a; /* NaN or inf ends as an infinite extent: every box is then decided in f64 */
                if (!open.empty() && !(a <= open.back().ext)) open.back().ext = a;
            }
            ++ordinal;
        } else {
            memcpy(w, &op, sizeof(rtgpu_op) * width);
            if (kind == RTGPU_OP_MEDIUM || kind == RTGPU_OP_PUSH) {
                w[1] = at[op.skip];
                if (kind == RTGPU_OP_PUSH) open.push_back(OpenPush{op.skip, (double)RT_CULL_MIN_EXTENT, w});
            } else if (kind != RTGPU_OP_POP) {
                w[1] = mat_class[op.a]; /* a primitive has no skip: word 1 carries the shade class of its material */
            }
        }
        i += width;
    }
    close_pushes(d->n_ops);
    *extent = (ext == ext && ext <= 3e38) ? nextafterf((float)ext, INFINITY) : INFINITY;
    *n_units = u;
    return RTGPU_OK;
}

rtgpu_status SceneArena::commit(rtgpu_scene *s) {
    void *p = nullptr;
    {
        /* stream-ordered pool, kept warm: re-creating a scene every frame reuses the same block */
        cudaMemPool_t pool;
        CK(cudaDeviceGetDefaultMemPool(&pool, s->device));
        uint64_t keep = ~0ull;
        CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    CK(cudaMallocAsync(&p, host.size(), 0));
    s->allocs.push_back(p);
    CK(cudaMemcpyAsync(p, host.data(), host.size(), cudaMemcpyHostToDevice, 0));
    CK(cudaStreamSynchronize(0));
    for (auto &f : fix) *f.second = static_cast<unsigned char *>(p) + f.first;
    return RTGPU_OK;
}

int grid_blocks(const rtgpu_scene *s) { return s->n_sm * kBlocksPerSM; } /* persistent grids: a multiple of the SM count (148 on a B200) */

constexpr size_t kBytesPerPath = 2 * (sizeof(RayRec) + sizeof(StateRec)) + sizeof(HitRec16) + 4 * 4 + kClasses * 4; /* 248 */

rtgpu_status ensure_pool(rtgpu_scene *s, size_t n_paths) {
    if (!s->ctl) {
        CK(cudaMalloc(&s->ctl, sizeof(WaveCtl)));
        CK(cudaMallocHost(&s->snap_host, sizeof(CtlSnapshot) * kSnapRing));
        CK(cudaMallocHost(&s->ctl_host, sizeof(WaveCtl)));
        for (int i = 0; i < kSnapRing; ++i) CK(cudaEventCreateWithFlags(&s->snap_ev[i], cudaEventDisableTiming));
    }
    if (n_paths <= s->cap) return RTGPU_OK;
    if (s->wave_mem) cudaFree(s->wave_mem);
    s->wave_mem = nullptr;
    s->cap = 0;
    size_t n = (n_paths + 255) & ~(size_t)255;
    CK(cudaMalloc(&s->wave_mem, n * kBytesPerPath + 256));
    char *p = static_cast<char *>(s->wave_mem);
    auto take_u = [&](uint32_t *&dst) { dst = reinterpret_cast<uint32_t *>(p); p += n * 4; };
    for (int k = 0; k < 2; ++k) {
        PathQueue &q = s->q[k];
        q.ray = reinterpret_cast<RayRec *>(p);
        p += n * sizeof(RayRec);
        q.st = reinterpret_cast<StateRec *>(p);
        p += n * sizeof(StateRec);
    }
    s->hb.rec = reinterpret_cast<HitRec16 *>(p);
    p += n * sizeof(HitRec16);
    s->class_idx = reinterpret_cast<uint32_t *>(p);
    p += n * 4 * kClasses;
    take_u(s->sort_keys[0]); take_u(s->sort_keys[1]); take_u(s->sort_idx[0]); take_u(s->sort_idx[1]);
    s->sort_bounds = reinterpret_cast<double *>(p + ((8 - ((size_t)p & 7)) & 7)); /* 7 doubles + SortGrid: inside the 256 spare bytes */
    s->iota_ready = false;
    {
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, s->sort_keys[0], s->sort_keys[1], s->sort_idx[0], s->sort_idx[1], (int)n, 0, kSortKeyBits);
        if (need > s->sort_temp_bytes) {
            if (s->sort_temp) cudaFree(s->sort_temp);
            s->sort_temp = nullptr;
            s->sort_temp_bytes = 0;
            CK(cudaMalloc(&s->sort_temp, need));
            s->sort_temp_bytes = need;
        }
    }
    s->cap = n;
    return RTGPU_OK;
}
rtgpu_status ensure_sink(rtgpu_scene *s, size_t npix, cudaStream_t stream) {
    if (npix <= s->sink_pix) return RTGPU_OK;
    if (s->sink_mem) cudaFree(s->sink_mem);
    s->sink_mem = nullptr;
    s->sink_pix = 0;
    size_t bytes = npix * (3 * 8 + 3 * 8 + 4);
    CK(cudaMalloc(&s->sink_mem, bytes));
    CK(cudaMemsetAsync(s->sink_mem, 0, bytes, stream)); /* k_finish leaves them cleared from then on */
    char *p = static_cast<char *>(s->sink_mem);
    s->fix = reinterpret_cast<long long *>(p);
    p += npix * 24;
    s->big = reinterpret_cast<double *>(p);
    p += npix * 24;
    s->flags = reinterpret_cast<uint32_t *>(p);
    s->sink_pix = npix;
    return RTGPU_OK;
}
rtgpu_status ensure_perpath(rtgpu_scene *s, size_t n) {
    if (n <= s->perpath_cap) return RTGPU_OK;
    if (s->perpath_mem) cudaFree(s->perpath_mem);
    s->perpath_mem = nullptr;
    s->perpath_cap = 0;
    CK(cudaMalloc(&s->perpath_mem, n * 28));
    char *p = static_cast<char *>(s->perpath_mem);
    s->cx = reinterpret_cast<double *>(p);
    s->cy = s->cx + n;
    s->cz = s->cy + n;
    s->path_len = reinterpret_cast<uint32_t *>(s->cz + n);
    s->perpath_cap = n;
    return RTGPU_OK;
}

struct WaveTiming {
    double ms_extend = 0, ms_shade = 0, ms_other = 0;
    uint64_t launches = 0, launches_extend = 0, iterations = 0;
};

/* Traces wp.n_paths paths with at most wp.pool of them in flight; radiance goes to `sk`.  Everything is enqueued on
 * `stream`; the host only looks at asynchronous snapshots of the control block to know when to stop. */
rtgpu_status run_paths(rtgpu_scene *s, const WaveParams &wp, const PathSink &sk, cudaStream_t stream, bool timing, WaveTiming *wt,
                       uint64_t *rays, uint64_t *rays_per_bounce, bool allow_tail) {
    rtgpu_status st = ensure_pool(s, wp.pool);
    if (st != RTGPU_OK) return st;
    if (wp.n_paths == 0 || wp.max_depth == 0) return RTGPU_OK; /* depth 0: every ray_color call returns black (main.rs:240-242) */
    CK(cudaMemsetAsync(s->ctl, 0, sizeof(WaveCtl), stream));
    const bool overlap = s->overlap != 0;
    if (overlap && !s->fork_ev) {
        CK(cudaEventCreateWithFlags(&s->fork_ev, cudaEventDisableTiming));
        for (int k = 0; k < kSideStreams; ++k) {
            CK(cudaStreamCreateWithFlags(&s->side[k], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&s->side_ev[k], cudaEventDisableTiming));
        }
    }
    const int blocks = grid_blocks(s);
    const uint32_t cap = (uint32_t)s->cap;
    bool sorting = s->sort_mode != 0 && wp.n_paths >= s->sort_min && wp.max_depth > 1;
    SortGrid *grid = reinterpret_cast<SortGrid *>(s->sort_bounds + 8);
    if (sorting && !s->iota_ready) {
        k_iota<<<blocks, 256, 0, stream>>>(s->sort_idx[0], cap);
        s->iota_ready = true;
    }
    if (sorting) { /* until the first hit points are known every key is 0 (the first iteration is camera rays in pixel order) */
        CK(cudaMemsetAsync(grid, 0, sizeof(SortGrid), stream));
    }
    size_t ev = 0;
    auto mark = [&]() -> rtgpu_status {
        if (!timing) return RTGPU_OK;
        if (ev == s->events.size()) {
            cudaEvent_t e;
            CK(cudaEventCreate(&e));
            s->events.push_back(e);
        }
        CK(cudaEventRecord(s->events[ev++], stream));
        return RTGPU_OK;
    };
    /* upper bound of the number of survivors that enter the coming iteration, as far as the host knows */
    uint32_t bound = wp.pool < wp.n_paths ? wp.pool : wp.n_paths;
    bool gen_done = false; /* host knows that every path has been generated */
    uint64_t alive_ub = wp.n_paths; /* paths that can still be alive, as far as the host knows (an upper bound) */
    const uint64_t max_iter = (uint64_t)wp.n_paths / (bound ? bound : 1) * (uint64_t)wp.max_depth + 2ull * wp.max_depth + 64;
    int cur = 0;
    uint64_t it = 0, seen = 0;
    bool finished = false, tail_ran = false;
    std::vector<int> phase_of_event; /* 0 other, 1 extend, 2 shade: what ran between event k and k+1 */
    for (; it < max_iter && !finished; ++it) {
        PathQueue &qin = s->q[cur], &qout = s->q[cur ^ 1];
        if ((st = mark()) != RTGPU_OK) return st;
        if (allow_tail && gen_done && alive_ub <= s->tail_max && it > 0) { /* the few paths that are left run to their end in one launch */
            if (s->medium_nesting > 1) k_tail<2><<<blocks, kThreads, 0, stream>>>(s->d, qin, s->ctl, cur, sk, wp);
            else if (s->medium_nesting == 1) k_tail<1><<<blocks, kThreads, 0, stream>>>(s->d, qin, s->ctl, cur, sk, wp);
            else k_tail<0><<<blocks, kThreads, 0, stream>>>(s->d, qin, s->ctl, cur, sk, wp);
            k_tail_done<<<1, 1, 0, stream>>>(s->ctl, cur);
            CK(cudaGetLastError());
            if (wt) wt->launches += 2;
            phase_of_event.push_back(0);
            finished = true;
            tail_ran = true;
            break;
        }
        k_iter_begin<<<1, 32, 0, stream>>>(s->ctl, cur, wp);
        const bool sort_now = sorting && it >= 1 && bound >= s->sort_min;
        bool regen_aside = false;
        if (!gen_done) { /* new camera rays go behind the survivors: beside the sort of the survivors when there is one */
            regen_aside = overlap && sort_now;
            if (regen_aside) {
                CK(cudaEventRecord(s->fork_ev, stream));
                CK(cudaStreamWaitEvent(s->side[0], s->fork_ev, 0));
            }
            k_regen<<<blocks, kThreads, 0, regen_aside ? s->side[0] : stream>>>(s->d, qin, s->ctl, cur, wp);
            if (regen_aside) CK(cudaEventRecord(s->side_ev[0], s->side[0]));
            if (wt) wt->launches += 1;
        }
        const uint32_t *perm = nullptr;
        if (sort_now) {
            k_pad_keys<<<blocks, 256, 0, stream>>>(s->ctl, s->sort_keys[0], bound);
            size_t tb = s->sort_temp_bytes;
            CK(cub::DeviceRadixSort::SortPairs(s->sort_temp, tb, s->sort_keys[0], s->sort_keys[1], s->sort_idx[0], s->sort_idx[1], (int)bound, 0,
                                               kSortKeyBits, stream));
            perm = s->sort_idx[1]; /* (the library's sort kernels are not counted in stats->launches: those are this file's) */
            if (wt) wt->launches += 1;
        }
        if (regen_aside) CK(cudaStreamWaitEvent(stream, s->side_ev[0], 0));
        phase_of_event.push_back(0);
        if ((st = mark()) != RTGPU_OK) return st;
        if (s->medium_nesting >= 1) {
            if (s->has_mesh) k_extend<1, true><<<blocks, kThreads, 0, stream>>>(s->d, qin, s->hb, s->ctl, cur, wp, perm, bound);
            else k_extend<1, false><<<blocks, kThreads, 0, stream>>>(s->d, qin, s->hb, s->ctl, cur, wp, perm, bound);
        } else {
            if (s->has_mesh) k_extend<0, true><<<blocks, kThreads, 0, stream>>>(s->d, qin, s->hb, s->ctl, cur, wp, perm, bound);
            else k_extend<0, false><<<blocks, kThreads, 0, stream>>>(s->d, qin, s->hb, s->ctl, cur, wp, perm, bound);
        }
        CK(cudaGetLastError()); /* a failed extend launch must not let the shade kernels read stale hits */
        phase_of_event.push_back(1);
        if ((st = mark()) != RTGPU_OK) return st;
        if (sorting && it == 0) { /* the origin grid of this render's keys, from the first hit points */
            CK(cudaMemsetAsync(s->sort_bounds, 0, 7 * sizeof(double), stream));
            k_sort_moments<<<blocks, kThreads, 0, stream>>>(qin, s->hb, s->ctl, cur, s->sort_bounds);
            k_sort_grid<<<1, 1, 0, stream>>>(s->sort_bounds, grid);
            if (wt) wt->launches += 2;
        }
        {
            k_classify<<<blocks, 256, 0, stream>>>(s->hb, s->ctl, cur, s->class_idx, cap);
            /* the class kernels read disjoint lists and append to the out queue through one atomic cursor: independent of each other.
             * Class 1 (the largest) stays on the caller's stream, the others go to side streams and are joined below. */
            int n_side = 0;
            if (overlap) CK(cudaEventRecord(s->fork_ev, stream));
#define RT_SHADE(C)                                                                                                                     \
    do {                                                                                                                                \
        cudaStream_t cs = stream;                                                                                                       \
        if (overlap && C != 1) {                                                                                                        \
            cs = s->side[n_side];                                                                                                       \
            CK(cudaStreamWaitEvent(cs, s->fork_ev, 0));                                                                                 \
        }                                                                                                                               \
        k_shade<C><<<s->n_sm * RT_SHADE_BLOCKS_PER_SM, kShadeThreads, 0, cs>>>(s->d, qin, s->hb, qout, s->ctl, cur,                     \
                                                                                s->class_idx + (size_t)C * cap, sk, wp, grid,            \
                                                                                sorting ? s->sort_keys[0] : nullptr);                    \
        if (overlap && C != 1) {                                                                                                        \
            CK(cudaEventRecord(s->side_ev[n_side], cs));                                                                                \
            ++n_side;                                                                                                                   \
        }                                                                                                                               \
    } while (0)
            int n_launched = 0;
            if (s->class_present[1]) { RT_SHADE(1); ++n_launched; }
            if (s->class_present[5]) { RT_SHADE(5); ++n_launched; }
            if (s->class_present[2]) { RT_SHADE(2); ++n_launched; }
            RT_SHADE(0);
            if (s->class_present[4]) { RT_SHADE(4); ++n_launched; }
            if (s->class_present[3]) { RT_SHADE(3); ++n_launched; }
            for (int k = 0; k < n_side; ++k) CK(cudaStreamWaitEvent(stream, s->side_ev[k], 0));
            if (wt) wt->launches += 4 + (uint64_t)n_launched; /* iter_begin, extend, classify, class 0 + the others */
#undef RT_SHADE
        }
        phase_of_event.push_back(2);
        /* snapshot of the control block: tells the host, a few iterations late, how many paths are left */
        const int slot = (int)(it % kSnapRing);
        CK(cudaMemcpyAsync(s->snap_host + slot, s->ctl, sizeof(CtlSnapshot), cudaMemcpyDeviceToHost, stream));
        CK(cudaEventRecord(s->snap_ev[slot], stream));
        cur ^= 1;
        /* consume the snapshots that have arrived (never waits, except when the ring is about to wrap) */
        while (seen <= it) {
            const int sl = (int)(seen % kSnapRing);
            /* the host stays at most kRunAhead iterations ahead of the device: the queue bound it sizes the sort with stays
             * fresh and a finished render is not followed by a long tail of empty launches; the device always has work queued */
            cudaError_t q = (it - seen >= (uint64_t)s->run_ahead) ? cudaEventSynchronize(s->snap_ev[sl]) : cudaEventQuery(s->snap_ev[sl]);
            if (q == cudaErrorNotReady) break;
            CK(q);
            const CtlSnapshot &sn = s->snap_host[sl];
            const int next_q = (int)((seen + 1) & 1); /* queue filled by the shade kernels of iteration `seen` */
            const uint32_t survivors = sn.count[next_q], left = wp.n_paths - sn.gen_next;
            gen_done = left == 0;
            /* How many survivors will the coming iterations carry?  Never more than what is alive now plus what can still be
             * generated; in the steady state of a regenerating pool about what the last known iteration carried.  This only
             * sizes the sort: survivors beyond it are walked in place, like the new camera rays, and the result is the same. */
            uint64_t b2 = (uint64_t)survivors + left;
            alive_ub = b2;
            if (b2 > wp.pool) b2 = wp.pool;
            uint64_t est = (uint64_t)survivors + survivors / 16 + 1024;
            bound = (uint32_t)(est < b2 ? est : b2);
            if (gen_done && survivors == 0) finished = true;
            ++seen;
        }
    }
    if (wt) {
        wt->launches_extend += it;
        wt->iterations += it;
    }
    (void)tail_ran;
    CK(cudaGetLastError());
    if ((st = mark()) != RTGPU_OK) return st;
    CK(cudaMemcpyAsync(s->ctl_host, s->ctl, sizeof(WaveCtl), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    const WaveCtl &c = *s->ctl_host;
    if (c.gen_next != wp.n_paths || c.count[cur] != 0) {
        g_err = "render did not finish within its iteration budget";
        return RTGPU_ERR_CUDA;
    }
    *rays += c.rays;
    if (rays_per_bounce) {
        /* rays of bounce b = paths that traced more than b segments */
        unsigned long long longer = 0;
        for (int L = RTGPU_MAX_BOUNCE_STATS; L >= 1; --L) {
            longer += c.len_hist[L];
            rays_per_bounce[L - 1] += longer;
        }
    }
    if (wt && timing) {
        for (size_t k = 0; k + 1 < ev && k < phase_of_event.size(); ++k) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, s->events[k], s->events[k + 1]));
            (phase_of_event[k] == 1 ? wt->ms_extend : phase_of_event[k] == 2 ? wt->ms_shade : wt->ms_other) += ms;
        }
    }
    return RTGPU_OK;
}

} // namespace

namespace {
/* Scratch of the entry points that have no scene to keep it in (rtgpu_resolve) or are parity probes (rtgpu_trace_batch,
 * rtgpu_camera_rays): stream-ordered allocations from the device's default pool, which is told to keep what is freed -- the second
 * call of a kind finds its buffers in the pool instead of going through cudaMalloc / cudaFree. */
cudaError_t scratch_alloc(void **p, size_t bytes) {
    static thread_local int tuned_device = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (tuned_device != dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        tuned_device = dev;
    }
    return cudaMallocAsync(p, bytes ? bytes : 1, 0);
}
template <typename T>
cudaError_t scratch_alloc(T **p, size_t bytes) { return scratch_alloc(reinterpret_cast<void **>(p), bytes); }
void scratch_free(void *p) {
    if (p) cudaFreeAsync(p, 0);
}
} // namespace

extern "C" {

const char *rtgpu_last_error(void) { return g_err.c_str(); }

int rtgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

rtgpu_status rtgpu_scene_create(const rtgpu_scene_desc *desc, int device, rtgpu_scene **out) {
    if (!out) {
        g_err = "out is NULL";
        return RTGPU_ERR_INVALID_ARG;
    }
    *out = nullptr;
    int medium_nesting = 0;
    bool has_mesh = true;
    rtgpu_status st = validate(desc, &medium_nesting, &has_mesh);
    if (st != RTGPU_OK) return st;
    if (device >= 0) CK(cudaSetDevice(device));
    int cur = 0;
    CK(cudaGetDevice(&cur));
    rtgpu_scene *s = new (std::nothrow) rtgpu_scene();
    if (!s) return RTGPU_ERR_OOM;
    s->device = cur;
    s->medium_nesting = medium_nesting;
    s->has_mesh = has_mesh;
    if (cudaDeviceGetAttribute(&s->n_sm, cudaDevAttrMultiProcessorCount, cur) != cudaSuccess || s->n_sm < 1) {
        cudaGetLastError();
        s->n_sm = 148;
    }
    {
        std::lock_guard<std::mutex> lock(g_park_mutex);
        if (cur >= 0 && cur < kMaxDevices && g_parked_valid[cur]) {
            static_cast<WaveRes &>(*s) = g_parked[cur];
            g_parked[cur] = WaveRes();
            g_parked_valid[cur] = false;
        }
    }
    if (const char *e = getenv("RTGPU_OVERLAP")) s->overlap = atoi(e);
    if (const char *e = getenv("RTGPU_SORT_MIN_OPS")) s->sort_min_ops = (uint32_t)atoll(e);
    s->sort_mode = desc->n_ops >= s->sort_min_ops ? 1 : 0;
    if (const char *e = getenv("RTGPU_SORT_RAYS")) s->sort_mode = atoi(e);
    if (const char *e = getenv("RTGPU_SORT_MIN")) s->sort_min = (uint32_t)atoll(e);
    if (const char *e = getenv("RTGPU_TAIL_MAX")) s->tail_max = (uint32_t)atoll(e);
    if (const char *e = getenv("RTGPU_RUN_AHEAD")) {
        int v = atoi(e);
        s->run_ahead = v < 1 ? 1 : (v > 32 ? 32 : v);
    }
    memset(&s->d, 0, sizeof s->d);
    SceneArena arena;
    /* shade class per material (see k_classify) */
    std::vector<uint8_t> mc(desc->n_materials ? desc->n_materials : 1, 2);
    {
        auto simple_tex = [&](uint32_t t) {
            const rtgpu_texture &x = desc->textures[t];
            if (x.kind == 0) return true;
            return x.kind == 1 && desc->textures[x.a].kind == 0 && desc->textures[x.b].kind == 0;
        };
        for (uint32_t i = 0; i < desc->n_materials; ++i) {
            const rtgpu_material &m = desc->materials[i];
            switch (m.kind) {
            case 0: mc[i] = simple_tex(m.tex) ? 1 : 2; break;
            case 1: mc[i] = 3; break;
            case 2: mc[i] = 4; break;
            case 3: mc[i] = 0; break;
            default: mc[i] = 5; break;
            }
            s->class_present[mc[i]] = true;
        }
    }
    std::vector<uint32_t> units;
    std::vector<double> boxes64;
    st = repack_program(desc, mc, units, boxes64, &s->d.extent, &s->d.n_units);
    if (st != RTGPU_OK) {
        rtgpu_scene_destroy(s);
        return st;
    }
    const float extent = s->d.extent;
    const uint32_t n_units = s->d.n_units;
    arena.add(reinterpret_cast<const uint4 *>(units.data()), units.size() / 4, &s->d.prog);
    arena.add(reinterpret_cast<const double2 *>(boxes64.data()), boxes64.size() / 2, &s->d.boxes64);
    arena.add(desc->xforms, desc->n_xforms, &s->d.xforms);
    arena.add(desc->chain, desc->n_chain, &s->d.chain);
    arena.add(desc->ctxs, desc->n_ctxs, &s->d.ctxs);
    arena.add(desc->materials, desc->n_materials, &s->d.mats);
    arena.add(desc->textures, desc->n_textures, &s->d.texs);
    arena.add(desc->perlins, desc->n_perlins, &s->d.perlins);
    arena.add(desc->images, desc->n_images, &s->d.images);
    arena.add(desc->texels, (size_t)desc->n_texel_bytes, &s->d.texels);
    arena.add(desc->lights, desc->n_lights, &s->d.lights);
    st = arena.commit(s);
    if (st != RTGPU_OK) {
        rtgpu_scene_destroy(s);
        return st;
    }
    s->d.n_units = n_units;
    s->d.extent = extent;
    s->d.n_lights = desc->n_lights;
    s->d.n_light_roots = desc->n_light_roots;
    s->d.cam = desc->camera;
    for (int c = 0; c < 3; ++c) s->d.background[c] = desc->background[c];
    *out = s;
    return RTGPU_OK;
}

rtgpu_status rtgpu_scene_destroy(rtgpu_scene *s) {
    if (!s) return RTGPU_OK;
    cudaSetDevice(s->device);
    for (void *p : s->allocs) cudaFreeAsync(p, 0);
    {
        std::lock_guard<std::mutex> lock(g_park_mutex);
        WaveRes &w = *s;
        if (s->device >= 0 && s->device < kMaxDevices && !g_parked_valid[s->device] && w.cap > 0) {
            g_parked[s->device] = w; /* keep the working memory for the next scene on this device */
            g_parked_valid[s->device] = true;
            w = WaveRes();
        }
    }
    s->release();
    delete s;
    return RTGPU_OK;
}

namespace {
struct EventPair { /* destroyed on every exit path */
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
};
} // namespace

rtgpu_status rtgpu_render(rtgpu_scene *s, const rtgpu_render_params *p, float *accum, rtgpu_stats *stats) {
    if (!s || !p || !accum) {
        g_err = "NULL argument";
        return RTGPU_ERR_INVALID_ARG;
    }
    if (p->width < 2 || p->height < 2 || p->spp_end < p->spp_begin || p->max_depth < 0 || p->max_depth > 4095) {
        g_err = "bad render parameters (width/height >= 2: u = (x+xi)/(W-1), main.rs:147; max_depth <= 4095)";
        return RTGPU_ERR_INVALID_ARG;
    }
    if (!(s->d.cam.time0 < s->d.cam.time1)) {
        g_err = "camera shutter is empty: gen_range(time0..time1) panics upstream (camera.rs:71)";
        return RTGPU_ERR_INVALID_SCENE;
    }
    CK(cudaSetDevice(s->device));
    cudaStream_t stream = static_cast<cudaStream_t>(p->stream);
    size_t npix = (size_t)p->width * (size_t)p->height;
    if (npix > 0x7fffffffu) {
        g_err = "image too large";
        return RTGPU_ERR_INVALID_ARG;
    }
    /* paths kept in flight: the caller's cap, else 15 Mi (3.9 GB of pool at 248 B per path); never more than the job.  A larger
     * pool is faster -- sorted rays are closer neighbours (C3, 192 spp: 1 354 / 1 478 / 1 523 / 1 558 / 1 595 / 1 625 Mrays/s at
     * 4 / 8 / 12 / 16 / 32 / 123 Mi paths) -- the default stays below 4 GB */
    size_t pool = p->max_paths_in_flight > 0 ? (size_t)p->max_paths_in_flight : ((size_t)15 << 20);
    if (pool < 4096) pool = 4096;
    if (pool > 0x7ffffff0u) pool = 0x7ffffff0u;
    /* one pass covers at most 32768 samples per pixel (the fixed-point pixel sums cannot overflow) and < 2^32 paths */
    size_t spp_per_pass = 0xfffffff0u / npix;
    if (spp_per_pass > 32768) spp_per_pass = 32768;
    if (const char *e = getenv("RTGPU_SPP_PER_PASS")) { /* tests: several passes on a small image */
        long long v = atoll(e);
        if (v >= 1 && (size_t)v < spp_per_pass) spp_per_pass = (size_t)v;
    }
    if (spp_per_pass < 1) spp_per_pass = 1;

    float *acc_dev = accum;
    if (!p->accum_on_device) {
        if (s->accum_cap < npix * 3) {
            if (s->accum_dev) cudaFree(s->accum_dev);
            s->accum_dev = nullptr;
            s->accum_cap = 0;
            CK(cudaMalloc(&s->accum_dev, npix * 3 * sizeof(float)));
            s->accum_cap = npix * 3;
        }
        acc_dev = s->accum_dev;
        if (p->accumulate) CK(cudaMemcpyAsync(acc_dev, accum, npix * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
    }
    rtgpu_status st = ensure_sink(s, npix, stream);
    if (st != RTGPU_OK) return st;
    EventPair ev;
    CK(cudaEventCreate(&ev.a));
    CK(cudaEventCreate(&ev.b));
    CK(cudaEventRecord(ev.a, stream));
    WaveTiming wt;
    uint64_t rays = 0;
    uint64_t per_bounce[RTGPU_MAX_BOUNCE_STATS];
    memset(per_bounce, 0, sizeof per_bounce);
    PathSink sk;
    memset(&sk, 0, sizeof sk);
    sk.fix = s->fix;
    sk.big = s->big;
    sk.flags = s->flags;
    bool first = true;
    if (p->spp_end == p->spp_begin && !p->accumulate) CK(cudaMemsetAsync(acc_dev, 0, npix * 3 * sizeof(float), stream));
    for (int64_t sb = p->spp_begin; sb < p->spp_end; sb += (int64_t)spp_per_pass) {
        int64_t ns = p->spp_end - sb < (int64_t)spp_per_pass ? p->spp_end - sb : (int64_t)spp_per_pass;
        WaveParams wp;
        wp.seed = p->seed;
        wp.first_pixel = 0;
        wp.npix_wave = (uint32_t)npix;
        wp.sample_begin = (uint32_t)sb;
        wp.n_paths = (uint32_t)(npix * (size_t)ns);
        wp.width = p->width;
        wp.height = p->height;
        wp.max_depth = (uint32_t)p->max_depth;
        wp.pool = (uint32_t)(pool < wp.n_paths ? pool : wp.n_paths);
        st = run_paths(s, wp, sk, stream, stats != nullptr, &wt, &rays, per_bounce, true);
        if (st != RTGPU_OK) return st;
        k_finish<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(sk, acc_dev, (uint32_t)npix, (first && !p->accumulate) ? 0 : 1);
        wt.launches += 1;
        first = false;
    }
    CK(cudaGetLastError());
    if (!p->accum_on_device) CK(cudaMemcpyAsync(accum, acc_dev, npix * 3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CK(cudaEventRecord(ev.b, stream));
    CK(cudaStreamSynchronize(stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev.a, ev.b));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->rays = rays;
        stats->samples = (uint64_t)npix * (uint64_t)(p->spp_end - p->spp_begin);
        stats->ms_total = ms;
        stats->ms_extend = wt.ms_extend;
        stats->ms_shade = wt.ms_shade;
        stats->ms_other = wt.ms_other;
        stats->launches = wt.launches;
        stats->launches_extend = wt.launches_extend;
        stats->pool_paths = s->cap;
        stats->pool_bytes = s->cap * kBytesPerPath + s->sink_pix * 52;
        memcpy(stats->rays_per_bounce, per_bounce, sizeof per_bounce);
    }
    return RTGPU_OK;
}

rtgpu_status rtgpu_trace_batch(rtgpu_scene *s, const rtgpu_ray *rays, int64_t n, const uint64_t *keys, rtgpu_hit *out,
                               double *ms_kernel) {
    if (!s || (n > 0 && (!rays || !out)) || n < 0) {
        g_err = "bad argument";
        return RTGPU_ERR_INVALID_ARG;
    }
    if (ms_kernel) *ms_kernel = 0.;
    if (n == 0) return RTGPU_OK;
    CK(cudaSetDevice(s->device));
    if (n > 0x7fffffff) {
        g_err = "batch too large";
        return RTGPU_ERR_INVALID_ARG;
    }
    rtgpu_ray *d_rays = nullptr;
    uint64_t *d_keys = nullptr;
    rtgpu_hit *d_out = nullptr;
    double *d_t = nullptr;
    uint32_t *d_u32 = nullptr; /* op[n], rng_pos[n], fetch counter */
    rtgpu_status st = RTGPU_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto body = [&]() -> rtgpu_status {
        CK(scratch_alloc(&d_rays, (size_t)n * sizeof(rtgpu_ray)));
        CK(scratch_alloc(&d_out, (size_t)n * sizeof(rtgpu_hit)));
        CK(scratch_alloc(&d_t, (size_t)n * sizeof(double)));
        CK(scratch_alloc(&d_u32, ((size_t)2 * n + 1) * sizeof(uint32_t)));
        CK(cudaMemset(d_u32 + 2 * n, 0, sizeof(uint32_t)));
        CK(cudaMemcpy(d_rays, rays, (size_t)n * sizeof(rtgpu_ray), cudaMemcpyHostToDevice));
        if (keys) {
            CK(scratch_alloc(&d_keys, (size_t)n * sizeof(uint64_t)));
            CK(cudaMemcpy(d_keys, keys, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice));
        }
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, 0));
        if (s->medium_nesting == 0)
            k_trace_extend<0><<<grid_blocks(s), kThreads>>>(s->d, d_rays, d_keys, d_t, d_u32, d_u32 + n, (uint32_t)n, d_u32 + 2 * n);
        else
            k_trace_extend<1><<<grid_blocks(s), kThreads>>>(s->d, d_rays, d_keys, d_t, d_u32, d_u32 + n, (uint32_t)n, d_u32 + 2 * n);
        CK(cudaEventRecord(e1, 0));
        k_trace_record<<<grid_blocks(s), kThreads>>>(s->d, d_rays, d_t, d_u32, d_u32 + n, d_out, (uint32_t)n);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms_kernel) *ms_kernel = ms;
        CK(cudaMemcpy(out, d_out, (size_t)n * sizeof(rtgpu_hit), cudaMemcpyDeviceToHost));
        return RTGPU_OK;
    };
    st = body();
    scratch_free(d_rays);
    scratch_free(d_keys);
    scratch_free(d_out);
    scratch_free(d_t);
    scratch_free(d_u32);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return st;
}

rtgpu_status rtgpu_camera_rays(rtgpu_scene *s, int32_t width, int32_t height, int64_t first_pixel, int64_t n, int32_t k,
                               uint64_t seed, rtgpu_ray *out, uint32_t *rng_pos) {
    if (!s || !out || n <= 0 || width < 2 || height < 2 || first_pixel < 0 || first_pixel + n > (int64_t)width * height) {
        g_err = "bad argument";
        return RTGPU_ERR_INVALID_ARG;
    }
    CK(cudaSetDevice(s->device));
    rtgpu_ray *d_out = nullptr;
    uint32_t *d_pos = nullptr;
    auto body = [&]() -> rtgpu_status {
        CK(scratch_alloc(&d_out, (size_t)n * sizeof(rtgpu_ray)));
        CK(scratch_alloc(&d_pos, (size_t)n * sizeof(uint32_t)));
        WaveParams wp;
        wp.seed = seed;
        wp.first_pixel = (uint32_t)first_pixel;
        wp.npix_wave = (uint32_t)n;
        wp.sample_begin = (uint32_t)k;
        wp.n_paths = (uint32_t)n;
        wp.width = width;
        wp.height = height;
        k_camera_rays<<<(unsigned)((n + 127) / 128), 128>>>(s->d, wp, d_out, d_pos);
        CK(cudaGetLastError());
        CK(cudaMemcpy(out, d_out, (size_t)n * sizeof(rtgpu_ray), cudaMemcpyDeviceToHost));
        if (rng_pos) CK(cudaMemcpy(rng_pos, d_pos, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        return RTGPU_OK;
    };
    rtgpu_status st = body();
    scratch_free(d_out);
    scratch_free(d_pos);
    return st;
}

rtgpu_status rtgpu_sample_radiance(rtgpu_scene *s, int32_t width, int32_t height, int64_t first_pixel, int64_t n, int32_t k,
                                   int32_t max_depth, uint64_t seed, double *out_rgb, uint32_t *out_rays) {
    if (!s || !out_rgb || n <= 0 || n > 0x7fffff00 || width < 2 || height < 2 || first_pixel < 0 ||
        first_pixel + n > (int64_t)width * height || max_depth < 0 || max_depth > 4095) {
        g_err = "bad argument";
        return RTGPU_ERR_INVALID_ARG;
    }
    CK(cudaSetDevice(s->device));
    rtgpu_status st = ensure_perpath(s, (size_t)n);
    if (st != RTGPU_OK) return st;
    WaveParams wp;
    wp.seed = seed;
    wp.first_pixel = (uint32_t)first_pixel;
    wp.npix_wave = (uint32_t)n;
    wp.sample_begin = (uint32_t)k;
    wp.n_paths = (uint32_t)n;
    wp.width = width;
    wp.height = height;
    wp.max_depth = (uint32_t)max_depth;
    wp.pool = (uint32_t)(n < 4096 ? 4096 : n);
    PathSink sk;
    memset(&sk, 0, sizeof sk);
    sk.cx = s->cx;
    sk.cy = s->cy;
    sk.cz = s->cz;
    sk.path_len = s->path_len;
    CK(cudaMemsetAsync(s->perpath_mem, 0, (size_t)s->perpath_cap * 28, 0));
    uint64_t rays = 0;
    st = run_paths(s, wp, sk, 0, false, nullptr, &rays, nullptr, false); /* the parity entry point stays on the class kernels */
    if (st != RTGPU_OK) return st;
    std::vector<double> tmp((size_t)n * 3);
    CK(cudaMemcpy(tmp.data(), s->cx, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tmp.data() + n, s->cy, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(tmp.data() + 2 * n, s->cz, (size_t)n * 8, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < n; ++i) {
        out_rgb[3 * i] = tmp[(size_t)i];
        out_rgb[3 * i + 1] = tmp[(size_t)(n + i)];
        out_rgb[3 * i + 2] = tmp[(size_t)(2 * n + i)];
    }
    if (out_rays) CK(cudaMemcpy(out_rays, s->path_len, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return RTGPU_OK;
}

rtgpu_status rtgpu_resolve(const float *accum, int32_t width, int32_t height, int32_t spp, int32_t accum_on_device, uint8_t *rgb8) {
    if (!accum || !rgb8 || width <= 0 || height <= 0 || spp <= 0) {
        g_err = "bad argument";
        return RTGPU_ERR_INVALID_ARG;
    }
    size_t npix = (size_t)width * (size_t)height;
    float *d_acc = nullptr;
    uint8_t *d_rgb = nullptr;
    auto body = [&]() -> rtgpu_status {
        const float *src = accum;
        if (!accum_on_device) {
            CK(scratch_alloc(&d_acc, npix * 3 * sizeof(float)));
            CK(cudaMemcpy(d_acc, accum, npix * 3 * sizeof(float), cudaMemcpyHostToDevice));
            src = d_acc;
        }
        CK(scratch_alloc(&d_rgb, npix * 3));
        k_resolve<<<(unsigned)((npix + 255) / 256), 256>>>(src, width, height, spp, d_rgb);
        CK(cudaGetLastError());
        CK(cudaMemcpy(rgb8, d_rgb, npix * 3, cudaMemcpyDeviceToHost));
        return RTGPU_OK;
    };
    rtgpu_status st = body();
    scratch_free(d_acc);
    scratch_free(d_rgb);
    return st;
}


/* ---- spp sharding over the GPUs of one box -----------------------------------------------------------------------
 * Samples of a pixel are i.i.d. and the reference already sums them (main.rs:144-151): device g of G renders ALL pixels
 * for sample indices [b + g*n/G, b + (g+1)*n/G), one host thread per device as upstream has one per scanline block
 * (main.rs:109-162), and the per-device float sums meet in ONE ncclReduce(sum) on the first device (main.rs:171-183).
 * NCCL is looked up at run time (dlopen): librtgpu.so carries no link-time dependency on it, and a process that has
 * already loaded a libnccl.so.2 (PyTorch) shares that copy. */
} // extern "C"

#include <dlfcn.h>

#include <thread>

namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load() {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        Reduce = reinterpret_cast<decltype(Reduce)>(dlsym(lib, "ncclReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        return CommInitAll && CommDestroy && GroupStart && GroupEnd && Reduce;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;
constexpr int kNcclFloat32 = 7, kNcclSum = 0; /* ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0) */
} // namespace

extern "C" {

rtgpu_status rtgpu_multi_render(const rtgpu_scene_desc *desc, const rtgpu_render_params *params, int32_t n_devices,
                                const int32_t *devices, float *accum_root, rtgpu_stats *stats, double *ms_reduce) {
    if (!desc || !params || !accum_root || n_devices < 1 || n_devices > kMaxDevices) {
        g_err = "bad argument";
        return RTGPU_ERR_INVALID_ARG;
    }
    if (params->spp_end < params->spp_begin || params->accumulate || params->accum_on_device) {
        g_err = "rtgpu_multi_render writes host memory and does not accumulate";
        return RTGPU_ERR_INVALID_ARG;
    }
    int have = rtgpu_device_count();
    std::vector<int> devs((size_t)n_devices);
    for (int g = 0; g < n_devices; ++g) {
        devs[g] = devices ? devices[g] : g;
        if (devs[g] < 0 || devs[g] >= have) {
            g_err = "device ordinal out of range";
            return RTGPU_ERR_INVALID_ARG;
        }
    }
    const size_t n_float = (size_t)3 * (size_t)params->width * (size_t)params->height;
    std::vector<void *> comms((size_t)n_devices, nullptr);
    if (n_devices > 1) {
        std::lock_guard<std::mutex> lock(g_nccl_mutex);
        if (!g_nccl.load()) {
            g_err = "libnccl.so.2 not found: rtgpu_multi_render needs NCCL for more than one device";
            return RTGPU_ERR_NCCL;
        }
        int rc = g_nccl.CommInitAll(comms.data(), n_devices, devs.data());
        if (rc != 0) {
            g_err = std::string("ncclCommInitAll: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "failed");
            return RTGPU_ERR_NCCL;
        }
    }
    std::vector<float *> accum((size_t)n_devices, nullptr);
    std::vector<rtgpu_scene *> scenes((size_t)n_devices, nullptr);
    std::vector<rtgpu_stats> st((size_t)n_devices);
    std::vector<rtgpu_status> rc((size_t)n_devices, RTGPU_OK);
    std::vector<std::string> msg((size_t)n_devices);
    const long long n_spp = (long long)params->spp_end - params->spp_begin;
    {
        std::vector<std::thread> pool;
        for (int g = 0; g < n_devices; ++g)
            pool.emplace_back([&, g]() {
                auto fail = [&](rtgpu_status code) {
                    rc[g] = code;
                    msg[g] = rtgpu_last_error(); /* this thread's message */
                };
                if (cudaSetDevice(devs[g]) != cudaSuccess) {
                    g_err = "cudaSetDevice failed";
                    return fail(RTGPU_ERR_CUDA);
                }
                rtgpu_status r = rtgpu_scene_create(desc, devs[g], &scenes[g]);
                if (r != RTGPU_OK) return fail(r);
                if (cudaMalloc(&accum[g], n_float * sizeof(float)) != cudaSuccess) {
                    g_err = "out of device memory";
                    return fail(RTGPU_ERR_OOM);
                }
                rtgpu_render_params p = *params;
                p.spp_begin = params->spp_begin + (int32_t)(g * n_spp / n_devices);
                p.spp_end = params->spp_begin + (int32_t)((g + 1) * n_spp / n_devices);
                p.device = devs[g];
                p.stream = nullptr;
                p.accum_on_device = 1;
                p.accumulate = 0;
                r = rtgpu_render(scenes[g], &p, accum[g], &st[g]);
                if (r != RTGPU_OK) return fail(r);
            });
        for (auto &t : pool) t.join();
    }
    rtgpu_status result = RTGPU_OK;
    for (int g = 0; g < n_devices && result == RTGPU_OK; ++g)
        if (rc[g] != RTGPU_OK) {
            result = rc[g];
            g_err = "device " + std::to_string(devs[g]) + ": " + msg[g];
        }
    double red_ms = 0.;
    if (result == RTGPU_OK && n_devices > 1) {
        auto t0 = std::chrono::steady_clock::now();
        int nrc = g_nccl.GroupStart();
        for (int g = 0; g < n_devices && nrc == 0; ++g) {
            cudaSetDevice(devs[g]);
            nrc = g_nccl.Reduce(accum[g], accum[g], n_float, kNcclFloat32, kNcclSum, 0, comms[g], (cudaStream_t)0);
        }
        int erc = g_nccl.GroupEnd();
        if (nrc == 0) nrc = erc;
        for (int g = 0; g < n_devices; ++g) {
            cudaSetDevice(devs[g]);
            if (cudaStreamSynchronize(0) != cudaSuccess && nrc == 0) nrc = -1;
        }
        red_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (nrc != 0) {
            g_err = std::string("ncclReduce: ") + (nrc > 0 && g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "failed");
            result = RTGPU_ERR_NCCL;
        }
    }
    if (result == RTGPU_OK) {
        cudaSetDevice(devs[0]);
        if (cudaMemcpy(accum_root, accum[0], n_float * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
            g_err = "copy of the reduced image failed";
            result = RTGPU_ERR_CUDA;
        }
    }
    for (int g = 0; g < n_devices; ++g) {
        cudaSetDevice(devs[g]);
        if (accum[g]) cudaFree(accum[g]);
        if (scenes[g]) rtgpu_scene_destroy(scenes[g]);
        if (comms[g]) g_nccl.CommDestroy(comms[g]);
    }
    if (result == RTGPU_OK && stats) {
        memset(stats, 0, sizeof *stats);
        for (int g = 0; g < n_devices; ++g) {
            stats->rays += st[g].rays;
            stats->samples += st[g].samples;
            stats->launches += st[g].launches;
            stats->launches_extend += st[g].launches_extend;
            stats->ms_extend += st[g].ms_extend;
            stats->ms_shade += st[g].ms_shade;
            stats->ms_other += st[g].ms_other;
            if (st[g].ms_total > stats->ms_total) stats->ms_total = st[g].ms_total; /* the slowest device */
            if (st[g].pool_paths > stats->pool_paths) stats->pool_paths = st[g].pool_paths;
            if (st[g].pool_bytes > stats->pool_bytes) stats->pool_bytes = st[g].pool_bytes;
            for (int b = 0; b < RTGPU_MAX_BOUNCE_STATS; ++b) stats->rays_per_bounce[b] += st[g].rays_per_bounce[b];
        }
    }
    if (ms_reduce) *ms_reduce = red_ms;
    return result;
}

} // extern "C"
