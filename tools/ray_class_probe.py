"""k_trace_extend throughput by kind of ray on the book-2 final scene: camera rays, rays born inside the geometry, and rays born
far out in the global fog (what deep bounces mostly are: isotropic scatters up to 5000 units away).  With a
-DRT_DEBUG_COUNTERS library (RTGPU_LIB_PATH) it also prints the op mix per ray."""
import ctypes as C, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import rtb200 as rt

s = rt.Scene(seed=2022).build_config("final_scene", 1.0)
gpu = rt.GpuScene(s.lower(), 0)
G = rt.gpu_lib()
dbg = hasattr(G, "rtgpu_debug_counters")
rng = np.random.default_rng(1)
n = 1 << 22
cam, _ = gpu.camera_rays(800, 800, 0, 640000, 0)

def rays(origins):
    r = np.zeros(len(origins), dtype=cam.dtype)
    r["o"] = origins
    d = rng.normal(size=(len(origins), 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d
    r["tm"] = rng.uniform(0, 1, len(origins))
    r["t_min"] = 0.001
    r["t_max"] = 1.7976931348623157e308
    return r

def ball(radius, m):
    p = rng.normal(size=(m, 3)); p /= np.linalg.norm(p, axis=1, keepdims=True)
    return p * radius * rng.uniform(0, 1, (m, 1)) ** (1 / 3)

sets = {
    "camera": np.tile(cam, (n // len(cam) + 1))[:n],
    "inside the geometry (box [-400,50,-100]..[500,500,500])": rays(rng.uniform([-400, 50, -100], [500, 500, 500], size=(n, 3))),
    "fog, within 1500 of the origin": rays(ball(1500, n)),
    "fog, within 4500 of the origin": rays(ball(4500, n)),
}
buf = (C.c_ulonglong * 32)()
for name, r in sets.items():
    gpu.trace_batch(r[:1024])
    if dbg: G.rtgpu_debug_counters(buf, 1)
    hits, ms = gpu.trace_batch(r)
    line = f"{name:58s} {n / ms / 1e3:8.1f} Mrays/s  hit {hits['hit'].mean():.3f}"
    if dbg:
        G.rtgpu_debug_counters(buf, 1)
        line += f"  boxes/ray {2 * buf[8] / n:.1f} (lanes/round {buf[8] / max(buf[0], 1):.1f}) sphere {buf[1] / n:.2f} box {buf[2] / n:.2f} misc {buf[4] / n:.2f}"
    print(line)
