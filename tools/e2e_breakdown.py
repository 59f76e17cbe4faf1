import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import rtb200 as rt
s = rt.Scene(seed=2022).build_config("final_scene", 1.0)
lw = s.lower()
w = h = 800
accum = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
host = torch.empty((h, w, 3), dtype=torch.float32, pin_memory=True)
g0 = rt.GpuScene(lw, 0)
g0.render(w, h, 0, 24, accum_ptr=accum.data_ptr())
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g = rt.GpuScene(lw, 0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    _, st = g.render(w, h, 0, 24, accum_ptr=accum.data_ptr())
    torch.cuda.synchronize(); t2 = time.perf_counter()
    host.copy_(accum, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
    g.close(); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} render {1e3*(t2-t1):.2f} (device {st.ms_total:.2f}) d2h {1e3*(t3-t2):.2f} destroy {1e3*(t4-t3):.2f} ms")
