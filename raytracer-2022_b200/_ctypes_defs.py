"""ctypes mirrors of include/rt_graph.h and include/rtgpu.h (field order and sizes must match)."""
import ctypes as C

c_double3 = C.c_double * 3


class rt_texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("pad_", C.c_int32), ("p", C.c_double * 4)]


class rt_material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tex", C.c_int32), ("p", C.c_double * 4)]


class rt_object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mat", C.c_int32), ("c0", C.c_int32), ("c1", C.c_int32), ("p", C.c_double * 12)]


class rt_perlin(C.Structure):
    _fields_ = [
        ("randvec", (C.c_double * 3) * 256),
        ("perm_x", C.c_int32 * 256),
        ("perm_y", C.c_int32 * 256),
        ("perm_z", C.c_int32 * 256),
    ]


class rt_image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("offset", C.c_int64)]


class rt_camera(C.Structure):
    _fields_ = [
        ("lookfrom", c_double3), ("lookat", c_double3), ("vup", c_double3),
        ("vfov", C.c_double), ("aspect_ratio", C.c_double), ("aperture", C.c_double), ("focus_dist", C.c_double),
        ("origin", c_double3), ("lower_left_corner", c_double3), ("horizontal", c_double3), ("vertical", c_double3),
        ("u", c_double3), ("v", c_double3), ("w", c_double3),
        ("lens_radius", C.c_double), ("time0", C.c_double), ("time1", C.c_double),
    ]


class rt_graph(C.Structure):
    _fields_ = [
        ("textures", C.POINTER(rt_texture)),
        ("materials", C.POINTER(rt_material)),
        ("objects", C.POINTER(rt_object)),
        ("list_items", C.POINTER(C.c_int32)),
        ("perlins", C.POINTER(rt_perlin)),
        ("images", C.POINTER(rt_image)),
        ("texels", C.POINTER(C.c_uint8)),
        ("lights", C.POINTER(C.c_int32)),
        ("n_textures", C.c_int32), ("n_materials", C.c_int32), ("n_objects", C.c_int32), ("n_list_items", C.c_int32),
        ("n_perlins", C.c_int32), ("n_images", C.c_int32), ("n_lights", C.c_int32),
        ("world", C.c_int32),
        ("n_texel_bytes", C.c_int64),
        ("camera", rt_camera),
        ("background", c_double3),
    ]


# ---- rtgpu.h ------------------------------------------------------------------------------------
class rtgpu_op(C.Structure):
    _fields_ = [("kind_ctx", C.c_uint32), ("skip", C.c_uint32), ("a", C.c_uint32), ("b", C.c_uint32), ("d", C.c_double * 6)]


class rtgpu_xform(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("gid", C.c_uint32), ("p", C.c_double * 3)]


class rtgpu_ctx(C.Structure):
    _fields_ = [("start", C.c_uint32), ("len", C.c_uint32)]


class rtgpu_material(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("tex", C.c_uint32), ("p", C.c_double * 4)]


class rtgpu_texture(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("a", C.c_int32), ("b", C.c_int32), ("pad_", C.c_uint32), ("p", C.c_double * 4)]


class rtgpu_image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("offset", C.c_int64)]


class rtgpu_light(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("gid", C.c_uint32), ("p", C.c_double * 5)]


class rtgpu_camera(C.Structure):
    _fields_ = [
        ("origin", c_double3), ("lower_left_corner", c_double3), ("horizontal", c_double3), ("vertical", c_double3),
        ("u", c_double3), ("v", c_double3),
        ("lens_radius", C.c_double), ("time0", C.c_double), ("time1", C.c_double),
    ]


class rtgpu_scene_desc(C.Structure):
    _fields_ = [
        ("ops", C.POINTER(rtgpu_op)),
        ("xforms", C.POINTER(rtgpu_xform)),
        ("chain", C.POINTER(C.c_uint32)),
        ("ctxs", C.POINTER(rtgpu_ctx)),
        ("materials", C.POINTER(rtgpu_material)),
        ("textures", C.POINTER(rtgpu_texture)),
        ("perlins", C.c_void_p),
        ("images", C.POINTER(rtgpu_image)),
        ("texels", C.POINTER(C.c_uint8)),
        ("lights", C.POINTER(rtgpu_light)),
        ("n_ops", C.c_uint32), ("n_xforms", C.c_uint32), ("n_chain", C.c_uint32), ("n_ctxs", C.c_uint32),
        ("n_materials", C.c_uint32), ("n_textures", C.c_uint32), ("n_perlins", C.c_uint32), ("n_images", C.c_uint32),
        ("n_lights", C.c_uint32),
        ("max_push_depth", C.c_uint32), ("n_light_roots", C.c_uint32), ("pad_", C.c_uint32),
        ("n_texel_bytes", C.c_uint64),
        ("camera", rtgpu_camera),
        ("background", c_double3),
    ]


class rtgpu_ray(C.Structure):
    _fields_ = [("o", c_double3), ("d", c_double3), ("tm", C.c_double), ("t_min", C.c_double), ("t_max", C.c_double)]


class rtgpu_hit(C.Structure):
    _fields_ = [
        ("t", C.c_double), ("p", c_double3), ("n", c_double3), ("u", C.c_double), ("v", C.c_double),
        ("hit", C.c_int32), ("front_face", C.c_int32), ("gid", C.c_int32), ("face", C.c_int32),
        ("inst", C.c_int32), ("mat", C.c_int32), ("rng_pos", C.c_uint32), ("pad_", C.c_uint32),
    ]


class rtgpu_render_params(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("spp_begin", C.c_int32), ("spp_end", C.c_int32),
        ("max_depth", C.c_int32), ("device", C.c_int32),
        ("seed", C.c_uint64),
        ("max_paths_in_flight", C.c_int64),
        ("stream", C.c_void_p),
        ("accumulate", C.c_int32), ("accum_on_device", C.c_int32),
    ]


RTGPU_MAX_BOUNCE_STATS = 64


class rtgpu_stats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64), ("samples", C.c_uint64),
        ("ms_total", C.c_double), ("ms_extend", C.c_double), ("ms_shade", C.c_double), ("ms_other", C.c_double),
        ("launches", C.c_uint64), ("launches_extend", C.c_uint64),
        ("rays_per_bounce", C.c_uint64 * RTGPU_MAX_BOUNCE_STATS),
        ("pool_paths", C.c_uint64), ("pool_bytes", C.c_uint64),
    ]


# numpy dtypes of the batch records (same layout as the structs above)
import numpy as np  # noqa: E402

RAY_DTYPE = np.dtype([("o", "f8", 3), ("d", "f8", 3), ("tm", "f8"), ("t_min", "f8"), ("t_max", "f8")])
HIT_DTYPE = np.dtype([
    ("t", "f8"), ("p", "f8", 3), ("n", "f8", 3), ("u", "f8"), ("v", "f8"),
    ("hit", "i4"), ("front_face", "i4"), ("gid", "i4"), ("face", "i4"),
    ("inst", "i4"), ("mat", "i4"), ("rng_pos", "u4"), ("pad_", "u4"),
])
assert RAY_DTYPE.itemsize == C.sizeof(rtgpu_ray) == 72
assert HIT_DTYPE.itemsize == C.sizeof(rtgpu_hit) == 104
assert C.sizeof(rtgpu_op) == 64

# enums
TEX_SOLID, TEX_CHECKER, TEX_NOISE, TEX_IMAGE = range(4)
MAT_LAMBERTIAN, MAT_METAL, MAT_DIELECTRIC, MAT_DIFFUSE_LIGHT, MAT_ISOTROPIC = range(5)
(OBJ_SPHERE, OBJ_MOVING_SPHERE, OBJ_XY_RECT, OBJ_XZ_RECT, OBJ_YZ_RECT, OBJ_BOX, OBJ_TRIANGLE, OBJ_MEDIUM,
 OBJ_TRANSLATE, OBJ_ROTATE_Y, OBJ_FLIP_FACE, OBJ_ZOOM, OBJ_BVH_NODE, OBJ_LIST, OBJ_RING) = range(15)
(OP_NODE, OP_SPHERE, OP_MSPHERE, OP_RECT_XY, OP_RECT_XZ, OP_RECT_YZ, OP_BOX, OP_TRI, OP_MEDIUM, OP_PUSH,
 OP_POP, OP_RING) = range(12)
RTGPU_OK, RTGPU_ERR_INVALID_SCENE, RTGPU_ERR_INVALID_ARG, RTGPU_ERR_CUDA, RTGPU_ERR_OOM = range(5)
