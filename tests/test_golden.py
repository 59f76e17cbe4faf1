"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the CPU oracle).

CPU part: the oracle and the host scene builders still reproduce the committed vectors (pins the oracle,
the seeded scene functions, the BVH build and the lowering against accidental change).
GPU part (-m gpu): the CUDA path against the same committed vectors, through the C ABI.
"""
import glob
import os

import numpy as np
import pytest

import rtb200 as rt
from oracle_lib import OracleScene

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
NAMES = [os.path.basename(f)[:-4] for f in FILES]


def load(name):
    g = np.load(os.path.join(HERE, "golden", name + ".npz"))
    s = rt.Scene(seed=2022).build_config(name, float(g["aspect"]))
    return g, s


def check_hits(h, ref, rel):
    assert np.array_equal(h["hit"], ref["hit"])
    m = ref["hit"] == 1
    for f in ("gid", "face", "inst", "mat", "front_face"):
        assert np.array_equal(h[f][m], ref[f][m]), f
    assert np.array_equal(h["rng_pos"], ref["rng_pos"])
    assert np.allclose(h["t"][m], ref["t"][m], rtol=rel, atol=0)
    assert np.allclose(h["p"][m], ref["p"][m], rtol=rel, atol=1e-9)
    assert np.allclose(h["n"][m], ref["n"][m], rtol=0, atol=rel)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    g, s = load(name)
    ops = s.lower().ops()
    assert len(ops) == int(g["n_ops"])
    assert int(np.frombuffer(ops.tobytes(), dtype=np.uint32).astype(np.uint64).sum()) == int(g["ops_crc"])  # same scene, same tree
    orc = OracleScene(s)
    check_hits(orc.trace_batch(g["rays"], g["keys"]), g["hits"], 1e-12)
    w, h = int(g["width"]), int(g["height"])
    for p in range(0, w * h, 37):
        rgb, n = orc.sample(w, h, p % w, p // w, 0, max_depth=50, seed=77)
        assert n == g["sample_rays"][p]
        assert np.allclose(rgb, g["sample_radiance"][p], rtol=1e-9, atol=1e-12, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_reproduces_golden(name):
    g, s = load(name)
    gpu = rt.GpuScene(s.lower(), device=0)
    hits, _ = gpu.trace_batch(g["rays"], g["keys"])
    check_hits(hits, g["hits"], 1e-5)  # north_star: ids exact, t and normals within 1e-5 relative
    w, h = int(g["width"]), int(g["height"])
    rad, nrays = gpu.sample_radiance(w, h, 0, w * h, 0, max_depth=50, seed=77)
    ref = g["sample_radiance"]
    close = np.abs(rad - ref).max(axis=1) <= 1e-6 * np.maximum(1.0, np.abs(ref).max(axis=1))
    assert close.mean() >= 0.97 and (nrays == g["sample_rays"]).mean() >= 0.97
    accum, st = gpu.render(w, h, 0, 8, max_depth=50, seed=77)
    assert abs(int(st.rays) - int(g["render_rays"])) <= 0.01 * int(g["render_rays"])
    diff = np.abs(accum - g["accum_8spp"])
    assert (diff <= 1e-4 * np.maximum(1.0, np.abs(g["accum_8spp"]))).mean() >= 0.97
