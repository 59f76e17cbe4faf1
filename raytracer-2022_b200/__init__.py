"""raytracer-2022_b200 -- B200-native `ray_color` hot path of Jerx2y/Raytracer-2022.

Python here is orchestration only: ctypes bindings over
  * librthost.so  -- C++ host: the reference's scene-building API, BVH build, OBJ/JPEG loading,
                     lowering to flat arrays (raytracer-2022_b200/host/)
  * librtgpu.so   -- hand-written sm_100a CUDA kernels behind the C ABI of include/rtgpu.h
                     (raytracer-2022_b200/csrc/)
There is no CPU fallback: GpuScene raises if librtgpu.so is missing or no CUDA device exists.
The directory name contains a hyphen, so load this package with importlib (see rtb200.py at the
repo root).
"""
import ctypes as C
import os

import numpy as np

from . import _ctypes_defs as D
from ._ctypes_defs import HIT_DTYPE, RAY_DTYPE  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
ASSET_DIR = os.path.join(REPO_ROOT, "assets")
HOST_LIB_PATH = os.path.join(_HERE, "librthost.so")
GPU_LIB_PATH = os.environ.get("RTGPU_LIB_PATH") or os.path.join(_HERE, "librtgpu.so")  # override: A/B builds only

_host = None
_gpu = None


class RtError(RuntimeError):
    pass


def host_lib():
    global _host
    if _host is None:
        if not os.path.exists(HOST_LIB_PATH):
            raise RtError(f"{HOST_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(HOST_LIB_PATH)
        L.rtg_last_error.restype = C.c_char_p
        L.rtg_new.restype = C.c_void_p
        L.rtg_new.argtypes = [C.c_uint64]
        L.rtg_free.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.rtg_add_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, C.c_int]
        L.rtg_add_noise_texture.argtypes = [C.c_void_p, C.c_double]
        L.rtg_add_image_texture_file.argtypes = [C.c_void_p, C.c_char_p]
        L.rtg_add_image_texture_rgb8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.rtg_add_material.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, C.c_int]
        L.rtg_add_object.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int]
        L.rtg_add_list.argtypes = [C.c_void_p, ip, C.c_int]
        L.rtg_bvh.argtypes = [C.c_void_p, ip, C.c_int, C.c_double, C.c_double]
        L.rtg_load_obj_triangles.argtypes = [C.c_void_p, C.c_char_p, C.c_int, ip, ip]
        L.rtg_bounding_box.argtypes = [C.c_void_p, C.c_int32, C.c_double, C.c_double, dp]
        L.rtg_set_world.argtypes = [C.c_void_p, C.c_int32]
        L.rtg_set_lights.argtypes = [C.c_void_p, ip, C.c_int]
        L.rtg_set_camera.argtypes = [C.c_void_p, dp, dp, dp] + [C.c_double] * 6
        L.rtg_set_background.argtypes = [C.c_void_p, dp]
        L.rtg_build_config.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_char_p]
        L.rtg_rng_pos.restype = C.c_uint32
        L.rtg_rng_pos.argtypes = [C.c_void_p]
        L.rtg_graph.argtypes = [C.c_void_p, C.POINTER(D.rt_graph)]
        L.rtg_lower.restype = C.c_void_p
        L.rtg_lower.argtypes = [C.POINTER(D.rt_graph), C.c_int]
        L.rtg_lowered_free.argtypes = [C.c_void_p]
        L.rtg_lowered_desc.argtypes = [C.c_void_p, C.POINTER(D.rtgpu_scene_desc)]
        L.rtg_decode_jpeg.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p, C.c_size_t]
        L.rtg_encode_jpeg.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _host = L
    return _host


GPU_SYMBOLS = [
    "rtgpu_last_error", "rtgpu_device_count", "rtgpu_scene_create", "rtgpu_scene_destroy", "rtgpu_render",
    "rtgpu_trace_batch", "rtgpu_camera_rays", "rtgpu_sample_radiance", "rtgpu_resolve", "rtgpu_multi_render",
]


def gpu_lib():
    """librtgpu.so.  Raises when the CUDA extension has not been built -- there is no other path."""
    global _gpu
    if _gpu is None:
        if not os.path.exists(GPU_LIB_PATH):
            raise RtError(f"{GPU_LIB_PATH} is missing: the CUDA extension must be built (no CPU fallback)")
        L = C.CDLL(GPU_LIB_PATH)
        L.rtgpu_last_error.restype = C.c_char_p
        L.rtgpu_scene_create.argtypes = [C.POINTER(D.rtgpu_scene_desc), C.c_int, C.POINTER(C.c_void_p)]
        L.rtgpu_scene_destroy.argtypes = [C.c_void_p]
        L.rtgpu_render.argtypes = [C.c_void_p, C.POINTER(D.rtgpu_render_params), C.c_void_p, C.POINTER(D.rtgpu_stats)]
        L.rtgpu_trace_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.rtgpu_camera_rays.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_uint64,
                                        C.c_void_p, C.c_void_p]
        L.rtgpu_sample_radiance.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                            C.c_uint64, C.c_void_p, C.c_void_p]
        L.rtgpu_resolve.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        if hasattr(L, "rtgpu_multi_render"):  # (A/B builds of rt_kernels.cu alone, selected with RTGPU_LIB_PATH, may predate it)
            L.rtgpu_multi_render.argtypes = [C.POINTER(D.rtgpu_scene_desc), C.POINTER(D.rtgpu_render_params), C.c_int32, C.c_void_p,
                                             C.c_void_p, C.POINTER(D.rtgpu_stats), C.POINTER(C.c_double)]
        _gpu = L
    return _gpu


def _dbl(vals):
    arr = (C.c_double * len(vals))(*[float(v) for v in vals])
    return arr, len(vals)


def _ids(vals):
    vals = [int(v) for v in vals]
    return (C.c_int32 * max(len(vals), 1))(*vals), len(vals)


class Scene:
    """Host-side object graph; one method per constructor of the reference's scene API."""

    def __init__(self, seed=2022):
        self.L = host_lib()
        self.h = self.L.rtg_new(C.c_uint64(seed))
        self.seed = seed
        if not self.h:
            raise RtError(self._err())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.rtg_free(self.h)
                self.h = None
        except Exception:
            pass

    def _err(self):
        return (self.L.rtg_last_error() or b"").decode()

    def _chk(self, r):
        if r < 0:
            raise RtError(self._err())
        return r

    # textures (texture/mod.rs)
    def solid_color(self, c):
        a, n = _dbl(c)
        return self._chk(self.L.rtg_add_texture(self.h, D.TEX_SOLID, -1, -1, a, n))

    def checker_texture(self, c1, c2):
        return self._chk(self.L.rtg_add_texture(self.h, D.TEX_CHECKER, self.solid_color(c1), self.solid_color(c2), None, 0))

    def noise_texture(self, scale):
        return self._chk(self.L.rtg_add_noise_texture(self.h, float(scale)))

    def image_texture(self, path):
        return self._chk(self.L.rtg_add_image_texture_file(self.h, os.fsencode(path)))

    def image_texture_rgb8(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        h, w, _ = arr.shape
        return self._chk(self.L.rtg_add_image_texture_rgb8(self.h, w, h, arr.ctypes.data))

    # materials (material/mod.rs)
    def _tex(self, c_or_tex):
        return c_or_tex if isinstance(c_or_tex, (int, np.integer)) else self.solid_color(c_or_tex)

    def lambertian(self, albedo):
        return self._chk(self.L.rtg_add_material(self.h, D.MAT_LAMBERTIAN, self._tex(albedo), None, 0))

    def metal(self, albedo, fuzz):
        a, n = _dbl(list(albedo) + [fuzz])
        return self._chk(self.L.rtg_add_material(self.h, D.MAT_METAL, -1, a, n))

    def dielectric(self, ir):
        a, n = _dbl([ir])
        return self._chk(self.L.rtg_add_material(self.h, D.MAT_DIELECTRIC, -1, a, n))

    def diffuse_light(self, emit):
        return self._chk(self.L.rtg_add_material(self.h, D.MAT_DIFFUSE_LIGHT, self._tex(emit), None, 0))

    def isotropic(self, albedo):
        return self._chk(self.L.rtg_add_material(self.h, D.MAT_ISOTROPIC, self._tex(albedo), None, 0))

    # hittables (hittable/)
    def _obj(self, kind, mat, c0, c1, p):
        a, n = _dbl(p)
        return self._chk(self.L.rtg_add_object(self.h, kind, mat, c0, c1, a, n))

    def sphere(self, center, radius, mat):
        return self._obj(D.OBJ_SPHERE, mat, -1, -1, list(center) + [radius])

    def moving_sphere(self, c0, c1, t0, t1, radius, mat):
        return self._obj(D.OBJ_MOVING_SPHERE, mat, -1, -1, list(c0) + list(c1) + [t0, t1, radius])

    def xy_rect(self, x0, x1, y0, y1, k, mat):
        return self._obj(D.OBJ_XY_RECT, mat, -1, -1, [x0, x1, y0, y1, k])

    def xz_rect(self, x0, x1, z0, z1, k, mat):
        return self._obj(D.OBJ_XZ_RECT, mat, -1, -1, [x0, x1, z0, z1, k])

    def yz_rect(self, y0, y1, z0, z1, k, mat):
        return self._obj(D.OBJ_YZ_RECT, mat, -1, -1, [y0, y1, z0, z1, k])

    def boxes(self, p0, p1, mat):
        return self._obj(D.OBJ_BOX, mat, -1, -1, list(p0) + list(p1))

    def triangle(self, a, b, c, mat):
        return self._obj(D.OBJ_TRIANGLE, mat, -1, -1, list(a) + list(b) + list(c))

    def ring(self, r, t, mat):
        return self._obj(D.OBJ_RING, mat, -1, -1, [r, t])

    def constant_medium(self, boundary, density, albedo):
        return self._obj(D.OBJ_MEDIUM, self.isotropic(albedo), boundary, -1, [density])

    def translate(self, child, offset):
        return self._obj(D.OBJ_TRANSLATE, -1, child, -1, list(offset))

    def rotate_y(self, child, angle):
        return self._obj(D.OBJ_ROTATE_Y, -1, child, -1, [angle])

    def flip_face(self, child):
        return self._obj(D.OBJ_FLIP_FACE, -1, child, -1, [])

    def zoom(self, child, rate):
        return self._obj(D.OBJ_ZOOM, -1, child, -1, [rate])

    def hittable_list(self, ids):
        a, n = _ids(ids)
        return self._chk(self.L.rtg_add_list(self.h, a, n))

    def bvh_node(self, ids, time0=0.0, time1=1.0):
        a, n = _ids(ids)
        return self._chk(self.L.rtg_bvh(self.h, a, n, time0, time1))

    def load_obj_triangles(self, path, mat):
        first, count = C.c_int32(), C.c_int32()
        self._chk(self.L.rtg_load_obj_triangles(self.h, os.fsencode(path), mat, C.byref(first), C.byref(count)))
        return list(range(first.value, first.value + count.value))

    def bounding_box(self, obj, time0=0.0, time1=1.0):
        out = (C.c_double * 6)()
        r = self._chk(self.L.rtg_bounding_box(self.h, obj, time0, time1, out))
        return np.array(out[:]) if r else None

    def set_world(self, obj):
        self._chk(self.L.rtg_set_world(self.h, obj))

    def set_lights(self, ids):
        a, n = _ids(ids)
        self._chk(self.L.rtg_set_lights(self.h, a, n))

    def set_camera(self, lookfrom, lookat, vup, vfov, aspect_ratio, aperture=0.0, focus_dist=10.0, time0=0.0, time1=1.0):
        self._chk(self.L.rtg_set_camera(self.h, _dbl(lookfrom)[0], _dbl(lookat)[0], _dbl(vup)[0], vfov, aspect_ratio,
                                        aperture, focus_dist, time0, time1))

    def set_background(self, rgb):
        self.L.rtg_set_background(self.h, _dbl(rgb)[0])

    def build_config(self, name, aspect_ratio=1.0, asset_dir=ASSET_DIR):
        """One of the BASELINE.json config scenes (scene.rs function + camera + BvhNode::new_list)."""
        self._chk(self.L.rtg_build_config(self.h, name.encode(), float(aspect_ratio), os.fsencode(asset_dir)))
        return self

    def rng_pos(self):
        return self.L.rtg_rng_pos(self.h)

    def graph(self):
        g = D.rt_graph()
        self.L.rtg_graph(self.h, C.byref(g))
        return g

    def lower(self, elide_singletons=True):
        return Lowered(self, elide_singletons)


class Lowered:
    """Flat arrays of include/rtgpu.h produced by host/lower.cpp."""

    def __init__(self, scene, elide_singletons=True):
        self.L = host_lib()
        self.scene = scene  # keeps the graph alive
        g = scene.graph()
        self.h = self.L.rtg_lower(C.byref(g), 1 if elide_singletons else 0)
        if not self.h:
            raise RtError((self.L.rtg_last_error() or b"").decode())
        self.desc = D.rtgpu_scene_desc()
        self.L.rtg_lowered_desc(self.h, C.byref(self.desc))

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.rtg_lowered_free(self.h)
                self.h = None
        except Exception:
            pass

    def ops(self):
        n = self.desc.n_ops
        buf = (D.rtgpu_op * n).from_address(C.addressof(self.desc.ops.contents))
        return np.frombuffer(buf, dtype=np.dtype([("kind_ctx", "u4"), ("skip", "u4"), ("a", "u4"), ("b", "u4"), ("d", "f8", 6)])).copy()

    def nbytes(self):
        d = self.desc
        return (d.n_ops * 64 + d.n_xforms * C.sizeof(D.rtgpu_xform) + d.n_chain * 4 + d.n_ctxs * 8 +
                d.n_materials * C.sizeof(D.rtgpu_material) + d.n_textures * C.sizeof(D.rtgpu_texture) +
                d.n_perlins * C.sizeof(D.rt_perlin) + d.n_images * 16 + d.n_texel_bytes + d.n_lights * C.sizeof(D.rtgpu_light))


class GpuScene:
    """A scene resident in HBM (rtgpu_scene_create).  All compute goes through librtgpu.so."""

    def __init__(self, lowered, device=0):
        self.G = gpu_lib()
        self.lowered = lowered
        self.device = device
        h = C.c_void_p()
        st = self.G.rtgpu_scene_create(C.byref(lowered.desc), device, C.byref(h))
        if st != D.RTGPU_OK:
            raise RtError(f"rtgpu_scene_create: status {st}: {(self.G.rtgpu_last_error() or b'').decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.G.rtgpu_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st, what):
        if st != D.RTGPU_OK:
            raise RtError(f"{what}: status {st}: {(self.G.rtgpu_last_error() or b'').decode()}")

    def render(self, width, height, spp_begin, spp_end, max_depth=50, seed=2022, accum=None, accum_ptr=None,
               accumulate=False, max_paths_in_flight=0, stream=None):
        """rtgpu_render.  accum: float32 host array (3*w*h) or, with accum_ptr, a device pointer."""
        p = D.rtgpu_render_params()
        p.width, p.height, p.spp_begin, p.spp_end = width, height, spp_begin, spp_end
        p.max_depth, p.device, p.seed = max_depth, self.device, seed
        p.max_paths_in_flight = max_paths_in_flight
        p.stream = stream
        p.accumulate = 1 if accumulate else 0
        stats = D.rtgpu_stats()
        if accum_ptr is not None:
            p.accum_on_device = 1
            ptr = C.c_void_p(accum_ptr)
            out = None
        else:
            p.accum_on_device = 0
            if accum is None:
                accum = np.zeros((height, width, 3), dtype=np.float32)
            assert accum.dtype == np.float32 and accum.size == 3 * width * height and accum.flags["C_CONTIGUOUS"]
            ptr = C.c_void_p(accum.ctypes.data)
            out = accum
        self._chk(self.G.rtgpu_render(self.h, C.byref(p), ptr, C.byref(stats)), "rtgpu_render")
        return out, stats

    def trace_batch(self, rays, keys=None):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = rays.shape[0]
        out = np.zeros(n, dtype=HIT_DTYPE)
        kp = None
        if keys is not None:
            keys = np.ascontiguousarray(keys, dtype=np.uint64)
            assert keys.shape[0] == n
            kp = C.c_void_p(keys.ctypes.data)
        ms = C.c_double()
        self._chk(self.G.rtgpu_trace_batch(self.h, C.c_void_p(rays.ctypes.data), n, kp, C.c_void_p(out.ctypes.data), C.byref(ms)),
                  "rtgpu_trace_batch")
        return out, ms.value

    def camera_rays(self, width, height, first_pixel, n, k, seed=2022):
        out = np.zeros(n, dtype=RAY_DTYPE)
        pos = np.zeros(n, dtype=np.uint32)
        self._chk(self.G.rtgpu_camera_rays(self.h, width, height, first_pixel, n, k, seed, C.c_void_p(out.ctypes.data),
                                           C.c_void_p(pos.ctypes.data)), "rtgpu_camera_rays")
        return out, pos

    def sample_radiance(self, width, height, first_pixel, n, k, max_depth=50, seed=2022):
        out = np.zeros((n, 3), dtype=np.float64)
        rays = np.zeros(n, dtype=np.uint32)
        self._chk(self.G.rtgpu_sample_radiance(self.h, width, height, first_pixel, n, k, max_depth, seed,
                                               C.c_void_p(out.ctypes.data), C.c_void_p(rays.ctypes.data)), "rtgpu_sample_radiance")
        return out, rays

    def resolve(self, accum, width, height, spp, accum_ptr=None):
        out = np.zeros((height, width, 3), dtype=np.uint8)
        if accum_ptr is not None:
            st = self.G.rtgpu_resolve(C.c_void_p(accum_ptr), width, height, spp, 1, C.c_void_p(out.ctypes.data))
        else:
            accum = np.ascontiguousarray(accum, dtype=np.float32)
            st = self.G.rtgpu_resolve(C.c_void_p(accum.ctypes.data), width, height, spp, 0, C.c_void_p(out.ctypes.data))
        self._chk(st, "rtgpu_resolve")
        return out


def multi_render(lowered, width, height, spp_begin, spp_end, n_devices, devices=None, max_depth=50, seed=2022, max_paths_in_flight=0):
    """rtgpu_multi_render: the frame spp-sharded over `n_devices` GPUs of this box in ONE process, one NCCL reduce to the first.
    Returns (accum float32 H x W x 3 on the host, stats summed over the devices, ms of the reduce)."""
    G = gpu_lib()
    p = D.rtgpu_render_params()
    p.width, p.height, p.spp_begin, p.spp_end = width, height, spp_begin, spp_end
    p.max_depth, p.device, p.seed = max_depth, -1, seed
    p.max_paths_in_flight = max_paths_in_flight
    accum = np.zeros((height, width, 3), dtype=np.float32)
    stats = D.rtgpu_stats()
    ms = C.c_double()
    dv = None
    if devices is not None:
        arr = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        dv = C.cast(arr, C.c_void_p)
    st = G.rtgpu_multi_render(C.byref(lowered.desc), C.byref(p), n_devices, dv, C.c_void_p(accum.ctypes.data), C.byref(stats), C.byref(ms))
    if st != D.RTGPU_OK:
        raise RtError(f"rtgpu_multi_render: status {st}: {(G.rtgpu_last_error() or b'').decode()}")
    return accum, stats, ms.value


def encode_jpeg(path, rgb8, quality=100):
    """RGB8 (H x W x 3, top row first) -> baseline JPEG file; quality 100 is what main.rs:213-221 writes."""
    L = host_lib()
    arr = np.ascontiguousarray(rgb8, dtype=np.uint8)
    h, w, _ = arr.shape
    if L.rtg_encode_jpeg(os.fsencode(path), C.c_void_p(arr.ctypes.data), w, h, int(quality)) != 0:
        raise RtError((L.rtg_last_error() or b"").decode())


def decode_jpeg(path):
    L = host_lib()
    w, h = C.c_int(), C.c_int()
    if L.rtg_decode_jpeg(os.fsencode(path), C.byref(w), C.byref(h), None, 0) != 0:
        raise RtError((L.rtg_last_error() or b"").decode())
    out = np.zeros((h.value, w.value, 3), dtype=np.uint8)
    L.rtg_decode_jpeg(os.fsencode(path), C.byref(w), C.byref(h), C.c_void_p(out.ctypes.data), out.size)
    return out
