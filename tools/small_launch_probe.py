"""How long does a small k_extend launch take?  Renders final_scene at a few tiny sizes with RTGPU_TRACE_BOUNCES=1
(per-bounce table on stderr): the time of a launch against its ray count shows the fixed cost of a launch."""
import os, sys
os.environ["RTGPU_TRACE_BOUNCES"] = "1"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rtb200 as rt

s = rt.Scene(seed=2022).build_config("final_scene", 1.0)
gpu = rt.GpuScene(s.lower(), 0)
for w, spp in ((64, 1), (128, 1), (256, 1), (400, 1), (800, 1)):
    gpu.render(w, w, 0, spp)
    sys.stderr.write(f"---- {w}x{w} x {spp} spp\n")
    gpu.render(w, w, 0, spp)
