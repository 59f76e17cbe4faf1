"""Per-phase scheduling statistics of the extend warp loop (needs build_variants/dbg.so: -DRT_DEBUG_COUNTERS)."""
import ctypes as C, os, sys
os.environ["RTGPU_LIB_PATH"] = "build_variants/dbg.so"
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rtb200 as rt
name = sys.argv[1] if len(sys.argv) > 1 else "final_scene"
w, h, spp = {"final_scene": (800, 800, 4), "obj_cornell": (600, 600, 4), "random_scene": (400, 225, 32)}[name]
s = rt.Scene(seed=2022).build_config(name, w / h)
gpu = rt.GpuScene(s.lower(), 0)
G = rt.gpu_lib()
buf = (C.c_ulonglong * 16)()
G.rtgpu_debug_counters(buf, 1)
acc, st = gpu.render(w, h, 0, spp)
G.rtgpu_debug_counters(buf, 1)
names = ["NODE", "SPHERE", "BOX", "TRI", "MISC"]
tot_rounds = sum(buf[k] for k in range(5))
print(f"{name}: rays {st.rays}, rounds {tot_rounds}, avg lanes holding a ray per round {buf[5] / max(tot_rounds, 1):.1f}")
for k in range(5):
    if buf[k]:
        print(f"  {names[k]:7s} rounds {buf[k]:12d} ({buf[k] / tot_rounds:.3f})  ops {buf[8 + k]:12d}  lanes/round {buf[8 + k] / buf[k]:.2f}  ops/ray {buf[8 + k] / st.rays:.1f}")
