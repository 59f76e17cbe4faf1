"""ctypes binding of oracle/liboracle.so -- TESTS ONLY (the oracle is the checker, never the product)."""
import ctypes as C
import os

import numpy as np

import rtb200 as rt

D = rt.D
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "liboracle.so")

ORC_RAY = rt.RAY_DTYPE
ORC_HIT = rt.HIT_DTYPE


class orc_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("samples", C.c_uint64), ("seconds", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(ORACLE_PATH)
        L.orc_last_error.restype = C.c_char_p
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.POINTER(D.rt_graph)]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_trace_batch.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_void_p,
                                 C.POINTER(orc_stats)]
        L.orc_render_region.argtypes = [C.c_void_p] + [C.c_int] * 9 + [C.c_uint64, C.c_int, C.c_void_p, C.POINTER(orc_stats)]
        L.orc_sample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p]
        L.orc_camera_ray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p,
                                     C.POINTER(C.c_uint32)]
        L.orc_dump_rays.restype = C.c_int64
        L.orc_dump_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int64, C.c_void_p]
        L.orc_write_color.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_lights_pdf_value.restype = C.c_double
        L.orc_lights_pdf_value.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lights_random.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_texture_value.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_perlin_noise.restype = C.c_double
        L.orc_perlin_noise.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_perlin_turb.restype = C.c_double
        L.orc_perlin_turb.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_reflectance.restype = C.c_double
        L.orc_reflectance.argtypes = [C.c_double, C.c_double]
        L.orc_aabb_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_camera_derive.argtypes = [C.POINTER(D.rt_camera)]
        L.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_stream_gen.restype = C.c_double
        L.orc_stream_gen.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_stream_index.restype = C.c_uint32
        L.orc_stream_index.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_bvh_build.restype = C.c_int64
        L.orc_bvh_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_uint32, C.c_uint32,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_bvh_listing.restype = C.c_int64
        L.orc_bvh_listing.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def _v3(v):
    return np.ascontiguousarray(v, dtype=np.float64)


def make_rays(o, d, tm=0.0, t_min=0.001, t_max=np.finfo(np.float64).max):
    o = np.atleast_2d(np.asarray(o, dtype=np.float64))
    d = np.atleast_2d(np.asarray(d, dtype=np.float64))
    n = max(o.shape[0], d.shape[0])
    r = np.zeros(n, dtype=ORC_RAY)
    r["o"] = o
    r["d"] = d
    r["tm"] = tm
    r["t_min"] = t_min
    r["t_max"] = t_max
    return r


class OracleScene:
    def __init__(self, scene):
        self.L = lib()
        self.scene = scene
        g = scene.graph()
        self.h = self.L.orc_scene_create(C.byref(g))
        if not self.h:
            raise RuntimeError((self.L.orc_last_error() or b"").decode())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def trace_batch(self, rays, keys=None):
        rays = np.ascontiguousarray(rays, dtype=ORC_RAY)
        out = np.zeros(rays.shape[0], dtype=ORC_HIT)
        kp = None
        if keys is not None:
            keys = np.ascontiguousarray(keys, dtype=np.uint64)
            kp = C.c_void_p(keys.ctypes.data)
        r = self.L.orc_trace_batch(self.h, rays.shape[0], C.c_void_p(rays.ctypes.data), kp, C.c_void_p(out.ctypes.data))
        assert r == 0
        return out

    def render(self, w, h, spp_begin, spp_end, max_depth=50, seed=2022, threads=8, region=None):
        accum = np.zeros((h, w, 3), dtype=np.float64)
        st = orc_stats()
        if region is None:
            r = self.L.orc_render(self.h, w, h, spp_begin, spp_end, max_depth, seed, threads, C.c_void_p(accum.ctypes.data), C.byref(st))
        else:
            x0, x1, y0, y1 = region
            r = self.L.orc_render_region(self.h, w, h, x0, x1, y0, y1, spp_begin, spp_end, max_depth, seed, threads,
                                         C.c_void_p(accum.ctypes.data), C.byref(st))
        assert r == 0, self.L.orc_last_error()
        return accum, st

    def sample(self, w, h, x, y, k, max_depth=50, seed=2022):
        rgb = np.zeros(3, dtype=np.float64)
        rays = self.L.orc_sample(self.h, w, h, x, y, k, max_depth, seed, C.c_void_p(rgb.ctypes.data))
        return rgb, rays

    def camera_ray(self, w, h, x, y, k, seed=2022):
        out = np.zeros(1, dtype=ORC_RAY)
        pos = C.c_uint32()
        self.L.orc_camera_ray(self.h, w, h, x, y, k, seed, C.c_void_p(out.ctypes.data), C.byref(pos))
        return out[0], pos.value

    def dump_rays(self, w, h, spp, bounce, cap, max_depth=50, seed=2022):
        out = np.zeros(cap, dtype=ORC_RAY)
        n = self.L.orc_dump_rays(self.h, w, h, spp, max_depth, seed, bounce, cap, C.c_void_p(out.ctypes.data))
        return out[:n]

    def lights_pdf_value(self, o, v):
        o, v = _v3(o), _v3(v)
        return self.L.orc_lights_pdf_value(self.h, C.c_void_p(o.ctypes.data), C.c_void_p(v.ctypes.data))

    def lights_random(self, o, key):
        o = _v3(o)
        out = np.zeros(3)
        self.L.orc_lights_random(self.h, C.c_void_p(o.ctypes.data), key, C.c_void_p(out.ctypes.data))
        return out

    def texture_value(self, tex, u, v, p):
        p = _v3(p)
        out = np.zeros(3)
        self.L.orc_texture_value(self.h, tex, u, v, C.c_void_p(p.ctypes.data), C.c_void_p(out.ctypes.data))
        return out

    def perlin_noise(self, perlin, p):
        p = _v3(p)
        return self.L.orc_perlin_noise(self.h, perlin, C.c_void_p(p.ctypes.data))

    def perlin_turb(self, perlin, p, depth=7):
        p = _v3(p)
        return self.L.orc_perlin_turb(self.h, perlin, C.c_void_p(p.ctypes.data), depth)

    def bvh_build(self, ids, t0, t1, seed, s0, pos):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        cap = 4 * len(ids) + 8
        out = np.zeros(cap, dtype=np.int32)
        boxes = np.zeros((cap, 6), dtype=np.float64)
        after = C.c_uint32()
        n = self.L.orc_bvh_build(self.h, C.c_void_p(ids.ctypes.data), len(ids), t0, t1, seed, s0, pos,
                                 C.c_void_p(out.ctypes.data), cap, C.c_void_p(boxes.ctypes.data), C.byref(after))
        assert n >= 0
        nn = int((out[:n] == -1).sum())
        return out[:n], boxes[:nn], after.value

    def bvh_listing(self, root, cap=1 << 20):
        out = np.zeros(cap, dtype=np.int32)
        boxes = np.zeros((cap, 6), dtype=np.float64)
        n = self.L.orc_bvh_listing(self.h, root, C.c_void_p(out.ctypes.data), cap, C.c_void_p(boxes.ctypes.data))
        assert n >= 0
        nn = int((out[:n] == -1).sum())
        return out[:n], boxes[:nn]


def write_color(rgb_sum, spp):
    rgb_sum = _v3(rgb_sum)
    out = np.zeros(3, dtype=np.uint8)
    lib().orc_write_color(C.c_void_p(rgb_sum.ctypes.data), spp, C.c_void_p(out.ctypes.data))
    return out


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(C.c_void_p(c.ctypes.data), C.c_void_p(k.ctypes.data), C.c_void_p(o.ctypes.data))
    return o
