"""How much does ray coherence matter for k_extend?  Same rays, different orders (ms of the extend kernel only)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import rtb200 as rt
from oracle_lib import OracleScene

name = sys.argv[1] if len(sys.argv) > 1 else "final_scene"
s = rt.Scene(seed=2022).build_config(name, 1.0)
gpu = rt.GpuScene(s.lower(), 0)
orc = OracleScene(s)
N = 1 << 21
w = h = 512
cam, _ = gpu.camera_rays(w, h, 0, w * h, 0)
sec = orc.dump_rays(256, 256, 16, 2, N)  # bounce-2 rays in scan order (already somewhat coherent in origin)
print("secondary rays", len(sec))
rng = np.random.default_rng(0)


def run(label, rays):
    rays = np.ascontiguousarray(rays)
    gpu.trace_batch(rays[:1024])
    best = 1e9
    for _ in range(3):
        _, ms = gpu.trace_batch(rays)
        best = min(best, ms)
    print(f"{label:42s} n={len(rays):8d} extend {best:8.3f} ms  {len(rays) / best / 1e3:8.1f} Mrays/s")


run("camera rays, scan order", np.tile(cam, 8))
run("camera rays, shuffled", np.tile(cam, 8)[rng.permutation(8 * len(cam))])
one = np.tile(cam[w * h // 2 + w // 2], N)
run("one camera ray repeated", one)
run("secondary rays, dump order", sec)
perm = rng.permutation(len(sec))
run("secondary rays, shuffled", sec[perm])
# sorted by origin cell + direction octant on the host
o = sec["o"]; d = sec["d"]
lo = o.mean(0) - 2.5 * o.std(0); ext = 5 * o.std(0)
cell = np.clip(((o - lo) / ext * 64).astype(np.int64), 0, 63)
octant = (d[:, 0] < 0) * 1 + (d[:, 1] < 0) * 2 + (d[:, 2] < 0) * 4
key = ((cell[:, 0] * 64 + cell[:, 1]) * 64 + cell[:, 2]) * 8 + octant
run("secondary rays, sorted cell64+octant", sec[np.argsort(key, kind="stable")])
key2 = octant * (1 << 20) + ((cell[:, 0] * 64 + cell[:, 1]) * 64 + cell[:, 2])
run("secondary rays, sorted octant+cell64", sec[np.argsort(key2, kind="stable")])
# more keys: how fine must origin / direction be, and which goes first?
def cells(bits):
    return np.clip(((o - lo) / ext * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
def lin(c, bits):
    return (c[:, 0] << (2 * bits)) | (c[:, 1] << bits) | c[:, 2]
dn = d / np.linalg.norm(d, axis=1, keepdims=True)
theta = np.clip(((np.arccos(np.clip(dn[:, 1], -1, 1)) / np.pi) * 16).astype(np.int64), 0, 15)
phi = np.clip((((np.arctan2(dn[:, 2], dn[:, 0]) + np.pi) / (2 * np.pi)) * 32).astype(np.int64), 0, 31)
dir9 = theta * 32 + phi
for label, k in (("cell32, then direction 16x32", lin(cells(5), 5) * 512 + dir9),
                 ("cell16, then direction 16x32", lin(cells(4), 4) * 512 + dir9),
                 ("cell8, then direction 16x32", lin(cells(3), 3) * 512 + dir9),
                 ("direction 16x32, then cell32", dir9 * (1 << 15) + lin(cells(5), 5)),
                 ("octant, cell16, direction 16x32", (octant * 4096 + lin(cells(4), 4)) * 512 + dir9),
                 ("cell128 (7 bits), octant", lin(cells(7), 7) * 8 + octant)):
    run("secondary rays, sorted " + label, sec[np.argsort(k, kind="stable")])
one = np.tile(sec[12345], N)
run("one secondary ray repeated", one)
# cost by hit type
hits, _ = gpu.trace_batch(sec)
print("hit fraction", hits["hit"].mean())
