/*
 * lower.cpp -- object graph (include/rt_graph.h) -> flat GPU arrays (include/rtgpu.h).
 *
 * The reference visits children in a fixed, ray-independent order: BvhNode::hit tests its box,
 * then left, then right with t_max shrunk to the left hit (hittable/bvh/mod.rs:86-101);
 * HittableList::hit scans in insertion order with a running closest_so_far
 * (hittable/mod.rs:90-100).  Both are "pre-order walk with one running best t", so the whole
 * graph -- nested BvhNodes, lists, wrappers, media -- unrolls into ONE linear program:
 *
 *   NODE(box, skip)   box missed -> jump to skip, else fall through into the subtree
 *   <primitive>       hit(r, t_min, best_t); on a hit best_t = t
 *   PUSH(chain)/POP   ray into / out of the object space of a run of wrappers
 *   MEDIUM(skip)      ConstantMedium; its boundary is the sub-program [i+1, skip)
 *
 * A stackless walk of this array reproduces the reference's visiting order, tie-breaking
 * (Q1: later visit wins on equal t), singleton double visits (Q8) and per-node box tests
 * (Q15/Q16) by construction.
 */
#include "rt_host.h"

#include <cmath>
#include <cstring>

namespace rt {

namespace {

struct Lowerer {
    const rt_graph &g;
    Lowered &out;
    bool elide;
    std::vector<uint32_t> stack;     // xform ids around the op being emitted, outermost first
    std::vector<int8_t> has_medium;  // memo: -1 unknown, 0/1
    uint32_t push_depth = 0;
    int depth = 0;

    Lowerer(const rt_graph &g_, Lowered &o, bool e) : g(g_), out(o), elide(e), has_medium((size_t)g_.n_objects, -1) {}

    const rt_object &obj(int32_t id) const {
        if (id < 0 || id >= g.n_objects) throw Error{"lower: object id out of range"};
        return g.objects[id];
    }
    static bool is_wrapper(int kind) {
        return kind == RT_OBJ_TRANSLATE || kind == RT_OBJ_ROTATE_Y || kind == RT_OBJ_ZOOM || kind == RT_OBJ_FLIP_FACE;
    }
    bool contains_medium(int32_t id, int d = 0) {
        if (d > 100000) throw Error{"lower: object graph too deep (cycle?)"};
        const rt_object &o = obj(id);
        if (has_medium[id] >= 0) return has_medium[id] != 0;
        bool r = false;
        switch (o.kind) {
        case RT_OBJ_MEDIUM: r = true; break;
        case RT_OBJ_TRANSLATE: case RT_OBJ_ROTATE_Y: case RT_OBJ_ZOOM: case RT_OBJ_FLIP_FACE:
            r = contains_medium(o.c0, d + 1);
            break;
        case RT_OBJ_BVH_NODE: r = contains_medium(o.c0, d + 1) || contains_medium(o.c1, d + 1); break;
        case RT_OBJ_LIST:
            for (int i = 0; i < o.c1 && !r; ++i) r = contains_medium(g.list_items[o.c0 + i], d + 1);
            break;
        default: break;
        }
        has_medium[id] = r ? 1 : 0;
        return r;
    }

    uint32_t ctx_cache_size = ~0u;
    uint32_t ctx_cache_id = 0;
    uint64_t stack_version = 0, ctx_cache_version = ~0ull;
    uint32_t current_ctx() {
        if (stack.empty()) return 0; // context 0 = no wrappers
        if (ctx_cache_version == stack_version) return ctx_cache_id;
        rtgpu_ctx c;
        c.start = (uint32_t)out.chain.size();
        c.len = (uint32_t)stack.size();
        out.chain.insert(out.chain.end(), stack.begin(), stack.end());
        out.ctxs.push_back(c);
        ctx_cache_id = (uint32_t)out.ctxs.size() - 1;
        ctx_cache_version = stack_version;
        if (ctx_cache_id >= (1u << 24)) throw Error{"lower: too many wrapper contexts"};
        return ctx_cache_id;
    }

    uint32_t push_op(uint32_t kind, uint32_t ctx, uint32_t a, uint32_t b) {
        rtgpu_op op;
        std::memset(&op, 0, sizeof op);
        op.kind_ctx = kind | (ctx << 8);
        op.a = a;
        op.b = b;
        out.ops.push_back(op);
        return (uint32_t)out.ops.size() - 1;
    }
    void push_raw(const double d[8]) { // second slot of a two-slot op
        rtgpu_op op;
        std::memcpy(&op, d, sizeof op);
        out.ops.push_back(op);
    }
    uint32_t mat_of(const rt_object &o) const {
        if (o.mat < 0 || o.mat >= g.n_materials) throw Error{"lower: material id out of range"};
        return (uint32_t)o.mat;
    }

    int medium_depth = 0; // ConstantMedium boundaries the op being emitted is inside of
    void emit(int32_t id, bool in_boundary) {
        if (++depth > 100000) throw Error{"lower: object graph too deep (cycle?)"};
        const rt_object &o = obj(id);
        const double *p = o.p;
        switch (o.kind) {
        case RT_OBJ_SPHERE: {
            uint32_t i = push_op(RTGPU_OP_SPHERE, current_ctx(), mat_of(o), (uint32_t)id);
            for (int k = 0; k < 4; ++k) out.ops[i].d[k] = p[k];
            break;
        }
        case RT_OBJ_MOVING_SPHERE: {
            uint32_t i = push_op(RTGPU_OP_MSPHERE, current_ctx(), mat_of(o), (uint32_t)id);
            double *d = out.ops[i].d;
            d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; // center0
            d[3] = p[8];                            // radius
            d[4] = p[6]; d[5] = p[7];               // time0, time1
            double s1[8] = {p[3], p[4], p[5], 0, 0, 0, 0, 0}; // center1
            push_raw(s1);
            break;
        }
        case RT_OBJ_XY_RECT: case RT_OBJ_XZ_RECT: case RT_OBJ_YZ_RECT: {
            uint32_t kind = o.kind == RT_OBJ_XY_RECT ? RTGPU_OP_RECT_XY : (o.kind == RT_OBJ_XZ_RECT ? RTGPU_OP_RECT_XZ : RTGPU_OP_RECT_YZ);
            uint32_t i = push_op(kind, current_ctx(), mat_of(o), (uint32_t)id);
            for (int k = 0; k < 5; ++k) out.ops[i].d[k] = p[k];
            break;
        }
        case RT_OBJ_BOX: {
            uint32_t i = push_op(RTGPU_OP_BOX, current_ctx(), mat_of(o), (uint32_t)id);
            for (int k = 0; k < 6; ++k) out.ops[i].d[k] = p[k];
            break;
        }
        case RT_OBJ_TRIANGLE: {
            uint32_t i = push_op(RTGPU_OP_TRI, current_ctx(), mat_of(o), (uint32_t)id);
            for (int k = 0; k < 6; ++k) out.ops[i].d[k] = p[k];
            // Triangle::get_normal (triangle.rs:30-32), which upstream recomputes in every hit();
            // a pure function of the vertices, so the host computes it once with the same
            // operation order (this file is compiled with -ffp-contract=off).
            Vec3 a(p[0], p[1], p[2]), b(p[3], p[4], p[5]), c(p[6], p[7], p[8]);
            Vec3 n = cross(b - a, c - a).to_unit();
            double s1[8] = {p[6], p[7], p[8], n.x, n.y, n.z, 0, 0};
            push_raw(s1);
            break;
        }
        case RT_OBJ_RING: {
            uint32_t i = push_op(RTGPU_OP_RING, current_ctx(), mat_of(o), (uint32_t)id);
            out.ops[i].d[0] = (p[0] - p[1]) * (p[0] - p[1]); // dis_min, ring.rs:27
            out.ops[i].d[1] = (p[0] + p[1]) * (p[0] + p[1]); // dis_max, ring.rs:28
            break;
        }
        case RT_OBJ_MEDIUM: {
            // upstream's ConstantMedium<H> takes any Hittable as its boundary, another ConstantMedium included
            // (constantmedium.rs:14-43); the device walks two levels (RT_MAX_MEDIUM_NESTING)
            if (medium_depth >= 2) throw Error{"lower: ConstantMedium boundaries nested deeper than 2 are not supported"};
            uint32_t i = push_op(RTGPU_OP_MEDIUM, current_ctx(), mat_of(o), (uint32_t)id);
            out.ops[i].d[0] = -1. / p[0]; // neg_inv_density, constantmedium.rs:29,39
            ++medium_depth;
            emit(o.c0, true);
            --medium_depth;
            out.ops[i].skip = (uint32_t)out.ops.size();
            break;
        }
        case RT_OBJ_TRANSLATE: case RT_OBJ_ROTATE_Y: case RT_OBJ_ZOOM: case RT_OBJ_FLIP_FACE: {
            // maximal run of wrappers, outermost first
            std::vector<uint32_t> run;
            int32_t cur = id;
            bool moves_ray = false;
            for (;;) {
                const rt_object &w = obj(cur);
                if (!is_wrapper(w.kind)) break;
                rtgpu_xform xf;
                std::memset(&xf, 0, sizeof xf);
                xf.gid = (uint32_t)cur;
                switch (w.kind) {
                case RT_OBJ_TRANSLATE:
                    xf.kind = RTGPU_XF_TRANSLATE;
                    xf.p[0] = w.p[0]; xf.p[1] = w.p[1]; xf.p[2] = w.p[2];
                    moves_ray = true;
                    break;
                case RT_OBJ_ROTATE_Y: { // RotateY::new, hittable/mod.rs:189-191
                    xf.kind = RTGPU_XF_ROTATE_Y;
                    double radians = w.p[0] * (3.14159265358979323846264338327950288 / 180.0);
                    xf.p[0] = std::sin(radians);
                    xf.p[1] = std::cos(radians);
                    moves_ray = true;
                    break;
                }
                case RT_OBJ_ZOOM:
                    xf.kind = RTGPU_XF_ZOOM;
                    xf.p[0] = w.p[0];
                    moves_ray = true;
                    break;
                default:
                    xf.kind = RTGPU_XF_FLIP;
                    break;
                }
                out.xforms.push_back(xf);
                run.push_back((uint32_t)out.xforms.size() - 1);
                cur = w.c0;
                if (run.size() > 100000) throw Error{"lower: wrapper chain too long (cycle?)"};
            }
            uint32_t push_idx = 0;
            if (moves_ray) {
                uint32_t first = (uint32_t)out.chain.size();
                out.chain.insert(out.chain.end(), run.begin(), run.end());
                push_idx = push_op(RTGPU_OP_PUSH, 0, first, (uint32_t)run.size());
                ++push_depth;
                if (push_depth > out.max_push_depth) out.max_push_depth = push_depth;
            }
            stack.insert(stack.end(), run.begin(), run.end());
            ++stack_version;
            emit(cur, in_boundary);
            stack.resize(stack.size() - run.size());
            ++stack_version;
            if (moves_ray) {
                push_op(RTGPU_OP_POP, 0, 0, 0);
                out.ops[push_idx].skip = (uint32_t)out.ops.size();
                --push_depth;
            }
            break;
        }
        case RT_OBJ_BVH_NODE: {
            uint32_t i = push_op(RTGPU_OP_NODE, 0, 0, 0);
            for (int k = 0; k < 6; ++k) out.ops[i].d[k] = p[k];
            emit(o.c0, in_boundary);
            // `left == right` (bvh/mod.rs:44-47): the second visit re-accepts the same record for
            // deterministic subtrees; a ConstantMedium would draw again (Q8), so it is kept.
            if (!(elide && o.c0 == o.c1 && !contains_medium(o.c0))) emit(o.c1, in_boundary);
            out.ops[i].skip = (uint32_t)out.ops.size();
            break;
        }
        case RT_OBJ_LIST:
            if (o.c0 < 0 || o.c1 < 0 || o.c0 + o.c1 > g.n_list_items) throw Error{"lower: bad list range"};
            for (int k = 0; k < o.c1; ++k) emit(g.list_items[o.c0 + k], in_boundary);
            break;
        default:
            throw Error{"lower: unknown object kind"};
        }
        --depth;
    }
};

} // namespace

rtgpu_scene_desc Lowered::desc() const {
    rtgpu_scene_desc d;
    std::memset(&d, 0, sizeof d);
    d.ops = ops.data();
    d.xforms = xforms.data();
    d.chain = chain.data();
    d.ctxs = ctxs.data();
    d.materials = materials.data();
    d.textures = textures.data();
    d.perlins = perlins.data();
    d.images = images.data();
    d.texels = texels.data();
    d.lights = lights.data();
    d.n_ops = (uint32_t)ops.size();
    d.n_xforms = (uint32_t)xforms.size();
    d.n_chain = (uint32_t)chain.size();
    d.n_ctxs = (uint32_t)ctxs.size();
    d.n_materials = (uint32_t)materials.size();
    d.n_textures = (uint32_t)textures.size();
    d.n_perlins = (uint32_t)perlins.size();
    d.n_images = (uint32_t)images.size();
    d.n_lights = (uint32_t)lights.size();
    d.n_light_roots = n_light_roots;
    d.max_push_depth = max_push_depth;
    d.n_texel_bytes = texels.size();
    d.camera = camera;
    for (int c = 0; c < 3; ++c) d.background[c] = background[c];
    return d;
}

Lowered lower(const rt_graph &g, bool elide_singletons) {
    Lowered out;
    // context 0 = "no wrappers"
    out.ctxs.push_back(rtgpu_ctx{0, 0});
    Lowerer lw(g, out, elide_singletons);
    lw.ctx_cache_id = 0;
    lw.ctx_cache_version = lw.stack_version;
    if (g.world < 0 || g.world >= g.n_objects) throw Error{"lower: scene has no world object"};
    lw.emit(g.world, false);
    if (out.ops.empty()) throw Error{"lower: empty world"};

    for (int i = 0; i < g.n_materials; ++i) {
        const rt_material &m = g.materials[i];
        rtgpu_material o;
        std::memset(&o, 0, sizeof o);
        o.kind = (uint32_t)m.kind;
        o.tex = (uint32_t)(m.tex < 0 ? 0 : m.tex);
        if ((m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_DIFFUSE_LIGHT || m.kind == RT_MAT_ISOTROPIC) &&
            (m.tex < 0 || m.tex >= g.n_textures))
            throw Error{"lower: texture id out of range"};
        for (int k = 0; k < 4; ++k) o.p[k] = m.p[k];
        if (m.kind == RT_MAT_METAL) o.p[3] = m.p[3] < 1. ? m.p[3] : 1.;
        out.materials.push_back(o);
    }
    for (int i = 0; i < g.n_textures; ++i) {
        const rt_texture &t = g.textures[i];
        rtgpu_texture o;
        std::memset(&o, 0, sizeof o);
        o.kind = (uint32_t)t.kind;
        o.a = t.a;
        o.b = t.b;
        for (int k = 0; k < 4; ++k) o.p[k] = t.p[k];
        if (t.kind == RT_TEX_CHECKER && (t.a < 0 || t.a >= g.n_textures || t.b < 0 || t.b >= g.n_textures))
            throw Error{"lower: checker sub-texture out of range"};
        if (t.kind == RT_TEX_CHECKER && (t.a >= i || t.b >= i)) throw Error{"lower: checker sub-textures must precede it"};
        if (t.kind == RT_TEX_NOISE && (t.a < 0 || t.a >= g.n_perlins)) throw Error{"lower: perlin id out of range"};
        if (t.kind == RT_TEX_IMAGE && t.a >= g.n_images) throw Error{"lower: image id out of range"};
        out.textures.push_back(o);
    }
    static_assert(sizeof(rtgpu_perlin) == sizeof(rt_perlin), "perlin layouts must match");
    out.perlins.resize((size_t)g.n_perlins);
    if (g.n_perlins) std::memcpy(out.perlins.data(), g.perlins, sizeof(rt_perlin) * (size_t)g.n_perlins);
    for (int i = 0; i < g.n_images; ++i) {
        rtgpu_image im;
        im.width = g.images[i].width;
        im.height = g.images[i].height;
        im.offset = g.images[i].offset;
        if (im.width <= 0 || im.height <= 0 || im.offset < 0 ||
            im.offset + (int64_t)im.width * im.height * 3 > g.n_texel_bytes)
            throw Error{"lower: image outside the texel buffer"};
        out.images.push_back(im);
    }
    out.texels.assign(g.texels, g.texels + g.n_texel_bytes);
    /* `lights` (main.rs:89) is a HittableList: its members become lights[0 .. n_light_roots); a member that is itself a
     * HittableList is live upstream (hittable/mod.rs:121-132) and becomes a LIST entry whose members follow, level by level */
    {
        struct Pending { size_t entry; int32_t list_obj; int depth; };
        std::vector<Pending> todo;
        auto lower_member = [&](int32_t id, int depth) {
            if (id < 0 || id >= g.n_objects) throw Error{"lower: light id out of range"};
            const rt_object &o = g.objects[id];
            rtgpu_light l;
            std::memset(&l, 0, sizeof l);
            l.gid = (uint32_t)id;
            switch (o.kind) {
            case RT_OBJ_XY_RECT: l.kind = RTGPU_LIGHT_RECT_XY; break;
            case RT_OBJ_XZ_RECT: l.kind = RTGPU_LIGHT_RECT_XZ; break;
            case RT_OBJ_YZ_RECT: l.kind = RTGPU_LIGHT_RECT_YZ; break;
            case RT_OBJ_SPHERE: l.kind = RTGPU_LIGHT_SPHERE; break;
            case RT_OBJ_LIST: l.kind = RTGPU_LIGHT_LIST; break;
            default: l.kind = RTGPU_LIGHT_DEAD; break; // trait defaults, hittable/mod.rs:62-67 (wrappers do not forward either)
            }
            if (l.kind == RTGPU_LIGHT_LIST) {
                if (depth >= RTGPU_MAX_LIGHT_NESTING) throw Error{"lower: HittableLists nested deeper than RTGPU_MAX_LIGHT_NESTING inside `lights`"};
                if (o.c0 < 0 || o.c1 < 0 || o.c0 + o.c1 > g.n_list_items) throw Error{"lower: bad list range"};
                // HittableList::random picks with gen_range(0..len): an empty list panics upstream (hittable/mod.rs:130)
                if (o.c1 == 0) throw Error{"lower: an empty HittableList inside `lights` panics upstream (gen_range on an empty range)"};
                todo.push_back(Pending{out.lights.size(), id, depth});
            } else {
                int np = l.kind == RTGPU_LIGHT_SPHERE ? 4 : (l.kind == RTGPU_LIGHT_DEAD ? 0 : 5);
                for (int k = 0; k < np; ++k) l.p[k] = o.p[k];
            }
            out.lights.push_back(l);
        };
        for (int i = 0; i < g.n_lights; ++i) lower_member(g.lights[i], 0);
        out.n_light_roots = (uint32_t)out.lights.size();
        for (size_t k = 0; k < todo.size(); ++k) { // (todo grows while nested lists are found)
            Pending pd = todo[k];
            const rt_object &lo = g.objects[pd.list_obj];
            out.lights[pd.entry].p[0] = (double)out.lights.size();
            out.lights[pd.entry].p[1] = (double)lo.c1;
            for (int m = 0; m < lo.c1; ++m) lower_member(g.list_items[lo.c0 + m], pd.depth + 1);
            if (out.lights.size() > (1u << 20)) throw Error{"lower: too many lights"};
        }
    }
    const rt_camera &c = g.camera;
    for (int k = 0; k < 3; ++k) {
        out.camera.origin[k] = c.origin[k];
        out.camera.lower_left_corner[k] = c.lower_left_corner[k];
        out.camera.horizontal[k] = c.horizontal[k];
        out.camera.vertical[k] = c.vertical[k];
        out.camera.u[k] = c.u[k];
        out.camera.v[k] = c.v[k];
        out.background[k] = g.background[k];
    }
    out.camera.lens_radius = c.lens_radius;
    out.camera.time0 = c.time0;
    out.camera.time1 = c.time1;
    return out;
}

} // namespace rt
