"""Per-phase scheduling statistics of the extend warp loop (needs a -DRT_DEBUG_COUNTERS variant: build_variants/dbg.so
or the library named by RTGPU_LIB_PATH)."""
import ctypes as C, os, sys
os.environ.setdefault("RTGPU_LIB_PATH", "build_variants/dbg.so")
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rtb200 as rt
name = sys.argv[1] if len(sys.argv) > 1 else "final_scene"
w, h, spp = {"final_scene": (800, 800, 24), "obj_cornell": (600, 600, 8), "random_scene": (400, 225, 64)}[name]
s = rt.Scene(seed=2022).build_config(name, w / h)
gpu = rt.GpuScene(s.lower(), 0)
G = rt.gpu_lib()
buf = (C.c_ulonglong * 32)()
G.rtgpu_debug_counters(buf, 1)
acc, st = gpu.render(w, h, 0, spp)
G.rtgpu_debug_counters(buf, 1)
names = ["NODE", "SPHERE", "BOX", "TRI", "MISC"]
print(f"{name}: rays {st.rays}")
print(f"  NODE rounds {buf[0]}  lanes/round {buf[8] / max(buf[0], 1):.2f}  node-lane-rounds per ray {buf[8] / st.rays:.1f}")
for k in range(1, 5):
    print(f"  {names[k]:7s} leaf ops {buf[k]:12d}  per ray {buf[k] / st.rays:.2f}")
tot = buf[16] + buf[17] + buf[18]
print(f"  warp clocks: refill/start {buf[16] / tot:.3f}  NODE rounds {buf[17] / tot:.3f}  leaf section {buf[18] / tot:.3f}   (per ray {tot / st.rays * 32:.0f} warp-clocks x 32 lanes)")
for k in range(1, 5):
    if buf[25 + k]:
        print(f"  {names[k]:7s} case ran {buf[25 + k]:10d} times in a warp, {buf[20 + k] / buf[25 + k]:7.0f} clocks each, {buf[k] / buf[25 + k]:.2f} lanes; share of all warp clocks {buf[20 + k] / tot:.3f}")
print(f"  NODE round: {buf[17] / max(buf[0], 1):.0f} clocks per round")
