import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rtb200 as rt
s = rt.Scene(seed=2022).build_config("cornell_box", 1.0)
gpu = rt.GpuScene(s.lower(), 0)
g, r = gpu.sample_radiance(64, 64, 0, 4096, 3, max_depth=50, seed=4242)
print(g[:3], r[:8])
