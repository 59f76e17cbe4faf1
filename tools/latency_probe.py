"""Kernel time of rtgpu_trace_batch for tiny batches: the latency floor of one k_trace_extend launch."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import rtb200 as rt

name = sys.argv[1] if len(sys.argv) > 1 else "final_scene"
s = rt.Scene(seed=2022).build_config(name, 1.0)
gpu = rt.GpuScene(s.lower(), 0)
rng = np.random.default_rng(1)
cam, _ = gpu.camera_rays(800, 800, 0, 640000, 0)
sel = rng.permutation(640000)
# random rays from inside the scene volume (like deep-bounce rays)
m = 65536
inner = np.zeros(m, dtype=cam.dtype)
inner["o"] = rng.uniform([-400, 50, -100], [500, 500, 500], size=(m, 3))
d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
inner["d"] = d
inner["tm"] = rng.uniform(0, 1, m)
inner["t_min"] = 0.001
inner["t_max"] = 1.7976931348623157e308
for label, src in (("camera", cam[sel]), ("inner", inner)):
    for n in (1, 32, 1024, 4096, 32768, 65536):
        if n > len(src): continue
        best = 1e9
        for rep in range(4):
            _, ms = gpu.trace_batch(src[:n])
            best = min(best, ms)
        print(f"{name} {label:6s} n={n:6d}  kernel {best*1e3:9.1f} us")
