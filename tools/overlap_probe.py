"""Does overlapping independent waves hide the per-launch drain?  T threads, each with its own GpuScene (own wave buffers, own
stream), render spp/T samples of the same frame concurrently; compare the wall time against one wave of spp samples."""
import sys, threading, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import rtb200 as rt

w = h = 800
total_spp = int(sys.argv[1]) if len(sys.argv) > 1 else 96
s = rt.Scene(seed=2022).build_config("final_scene", 1.0)
low = s.lower()
for T in (1, 2, 3, 4, 6, 8):
    spp = total_spp // T
    gpus = [rt.GpuScene(low, 0) for _ in range(T)]
    streams = [torch.cuda.Stream() for _ in range(T)]
    accs = [torch.zeros((h, w, 3), dtype=torch.float32, device="cuda") for _ in range(T)]
    rays = [0] * T

    def work(k):
        _, st = gpus[k].render(w, h, k * spp, (k + 1) * spp, accum_ptr=accs[k].data_ptr(), stream=streams[k].cuda_stream,
                               max_paths_in_flight=w * h * spp)
        rays[k] = st.rays

    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{T} concurrent waves x {spp} spp: {best * 1e3:8.2f} ms  {sum(rays) / best / 1e6:8.1f} Mrays/s")
    for g in gpus: g.close()
