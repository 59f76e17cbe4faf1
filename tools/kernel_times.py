"""Per-kernel device times of one render (CUDA events inside rtgpu_render) for quick A/B runs on the GPU box."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import rtb200 as rt

cfgs = {"final_scene": (800, 800, 8), "random_scene": (400, 225, 64), "cornell_box": (600, 600, 16), "obj_cornell": (600, 600, 8), "wwscene": (1280, 720, 8),
        "cornell_smoke": (600, 600, 16), "cornell_book3": (600, 600, 16)}
names = sys.argv[1:] or ["final_scene"]
for name in names:
    w, h, spp = cfgs[name]
    spp = int(os.environ.get("KT_SPP", spp))  # KT_SPP=24: the wave size bench.py uses
    s = rt.Scene(seed=2022).build_config(name, w / h)
    gpu = rt.GpuScene(s.lower(), 0)
    pool = int(os.environ.get("KT_POOL", 0))  # paths in flight (0 = library default)
    gpu.render(w, h, 0, spp, max_paths_in_flight=pool)  # also sizes the pool
    best = None
    for rep in range(3):
        acc, st = gpu.render(w, h, 0, spp, max_paths_in_flight=pool)
        if best is None or st.ms_total < best.ms_total:
            best = st
    st = best
    print(f"{name} [{w}x{h} x {spp} spp, pool {st.pool_paths} paths / {st.pool_bytes / 1e9:.2f} GB, {st.launches_extend} iterations]: rays {st.rays} total {st.ms_total:.2f} ms extend {st.ms_extend:.2f} shade {st.ms_shade:.2f} other {st.ms_other:.2f} "
          f"-> {st.rays / st.ms_total / 1e3:.1f} Mrays/s (extend-only {st.rays / st.ms_extend / 1e3:.1f}, shade-only {st.rays / st.ms_shade / 1e3:.1f})")
