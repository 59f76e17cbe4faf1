"""tools/parity_report.py [out.json] -- what the parity tests observe on the B200, as numbers (tracked in profiles/).

Per config scene: (1) 2^18 bounce-1 rays of the oracle's own paths through rtgpu_trace_batch: rays whose hit/miss, object id,
face, wrapper, material, front_face or draw count differ from the oracle (must be 0), rays whose t differs at all, and the
largest relative difference of t (media: CUDA log vs glibc log; spheres: none); (2) one-sample radiance of 64x64 paths:
fraction within 1e-6, fraction with the same number of segments, total rays on both sides."""
import json
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np

import rtb200 as rt
from oracle_lib import OracleScene

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_r02.json"
report = {"what": __doc__.strip().split("\n\n")[1] if "\n\n" in __doc__ else "", "scenes": {}}
for name in ["random_scene", "cornell_box", "final_scene", "obj_cornell", "cornell_smoke", "cornell_book3", "wwscene"]:
    s = rt.Scene(seed=2022).build_config(name, 1.0)
    orc, gpu = OracleScene(s), rt.GpuScene(s.lower(), device=0)
    rays = orc.dump_rays(512, 512, 4, 1, 1 << 18)
    keys = np.arange(rays.shape[0], dtype=np.uint64) * 7919 + 13
    g, _ = gpu.trace_batch(rays, keys)
    o = orc.trace_batch(rays, keys)
    m = o["hit"] == 1
    id_mismatch = int((g["hit"] != o["hit"]).sum())
    for f in ("gid", "face", "inst", "mat", "front_face"):
        id_mismatch += int((g[f][m] != o[f][m]).sum())
    id_mismatch += int((g["rng_pos"] != o["rng_pos"]).sum())
    both = m & (g["hit"] == 1)
    t_o, t_g = o["t"][both], g["t"][both]
    fin = np.isfinite(t_o) & np.isfinite(t_g)
    rel = np.abs(t_g - t_o)[fin] / np.maximum(np.abs(t_o)[fin], 1e-300)
    is_medium = o["front_face"][both] == 1
    w = h = 64
    n = w * h
    gr, grays = gpu.sample_radiance(w, h, 0, n, 3, max_depth=50, seed=4242)
    orr = np.zeros((n, 3))
    orays = np.zeros(n, dtype=np.int64)
    for p in range(n):
        orr[p], orays[p] = orc.sample(w, h, p % w, p // w, 3, max_depth=50, seed=4242)
    finite = np.isfinite(orr).all(axis=1) & np.isfinite(gr).all(axis=1)
    close = np.abs(gr - orr).max(axis=1) <= 1e-6 * np.maximum(1.0, np.abs(orr).max(axis=1))
    report["scenes"][name] = {
        "batch_rays": int(rays.shape[0]), "hits": int(m.sum()), "id_or_draw_mismatches": id_mismatch,
        "t_not_bit_identical": int((t_g != t_o).sum()), "t_max_rel_diff": float(rel.max()) if rel.size else 0.0,
        "paths": n, "paths_radiance_within_1e-6": float((close | ~finite).mean()), "paths_same_length": float((grays == orays).mean()),
        "rays_gpu": int(grays.sum()), "rays_oracle": int(orays.sum()),
    }
    print(name, report["scenes"][name], flush=True)
json.dump(report, open(out, "w"), indent=1)
