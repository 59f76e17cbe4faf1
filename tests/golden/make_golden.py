"""Regenerates the golden fixtures in this directory from the CPU oracle.

The reference (Rust) cannot be run in this image and ships no golden vectors, so these fixtures
pin the ORACLE (a later change to oracle/ or to the host scene builders that alters any of
these numbers fails tests/test_golden.py), and they give the GPU tests a fixed target that
does not depend on the oracle being rebuilt on the GPU box.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import rtb200 as rt  # noqa: E402
from oracle_lib import OracleScene  # noqa: E402

SCENES = {"cornell_box": 1.0, "random_scene": 16 / 9, "final_scene": 1.0, "cornell_smoke": 1.0, "cornell_book3": 1.0, "wwscene": 16 / 9}


def main():
    for name, aspect in SCENES.items():
        s = rt.Scene(seed=2022).build_config(name, aspect)
        orc = OracleScene(s)
        w, h = (48, 27) if aspect != 1.0 else (32, 32)
        cam = np.array([orc.camera_ray(w, h, p % w, p // w, 1, seed=77)[0] for p in range(w * h)], dtype=rt.RAY_DTYPE)
        sec = orc.dump_rays(w, h, 2, 2, 1024, seed=77)
        rays = np.concatenate([cam, sec])
        keys = (np.arange(rays.shape[0], dtype=np.uint64) * np.uint64(2654435761)) + np.uint64(12345)
        hits = orc.trace_batch(rays, keys)
        rad = np.zeros((w * h, 3))
        nrays = np.zeros(w * h, dtype=np.int32)
        for p in range(w * h):
            rad[p], nrays[p] = orc.sample(w, h, p % w, p // w, 0, max_depth=50, seed=77)
        accum, st = orc.render(w, h, 0, 8, max_depth=50, seed=77, threads=4)
        ops = s.lower().ops()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), width=w, height=h, aspect=aspect, rays=rays, keys=keys, hits=hits,
                            sample_radiance=rad, sample_rays=nrays, accum_8spp=accum.astype(np.float32), render_rays=np.int64(st.rays),
                            n_ops=np.int64(len(ops)), ops_crc=np.int64(int(np.frombuffer(ops.tobytes(), dtype=np.uint32).astype(np.uint64).sum())))
        print(name, rays.shape[0], "rays,", int(hits["hit"].sum()), "hits,", st.rays, "render rays")


if __name__ == "__main__":
    main()
