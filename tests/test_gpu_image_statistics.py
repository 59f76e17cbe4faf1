"""Statistical image parity (BASELINE.json north_star: the RNG streams of the reference cannot be
reproduced, so the rendered image must be statistically equal at matched spp).

Ground truth GT = CPU oracle at 1024 spp.  In display space (after /spp, sqrt, clamp -- write_color
before quantisation, main.rs:295-297):
  * RMSE(gpu@16, GT) / RMSE(oracle@16, GT) within [0.9, 1.1]      (same noise level at matched spp)
  * mean signed error of gpu@1024 vs GT per channel < 0.5/255      (no bias)
  * relative mean-luminance difference in linear space < 1 %
The GPU and the oracle use DIFFERENT seeds here, so the streams are independent.
"""
import numpy as np
import pytest

import rtb200 as rt
from oracle_lib import OracleScene

pytestmark = pytest.mark.gpu


def display(accum, spp):
    x = np.nan_to_num(np.asarray(accum, dtype=np.float64), nan=0.0) / spp
    return np.clip(np.sqrt(np.maximum(x, 0.0)), 0.0, 0.999)


@pytest.mark.parametrize("name,aspect,w,h", [("cornell_box", 1.0, 64, 64), ("random_scene", 16 / 9, 96, 54), ("final_scene", 1.0, 64, 64),
                                              ("cornell_smoke", 1.0, 48, 48)])
def test_image_statistics(name, aspect, w, h):
    s = rt.Scene(seed=2022).build_config(name, aspect)
    orc = OracleScene(s)
    gpu = rt.GpuScene(s.lower(), device=0)
    hi, lo = 1024, 16
    gt, _ = orc.render(w, h, 0, hi, seed=101, threads=16)
    o_lo, _ = orc.render(w, h, 0, lo, seed=202, threads=16)
    g_lo, _ = gpu.render(w, h, 0, lo, seed=303)
    g_hi, st = gpu.render(w, h, 0, hi, seed=404)
    GT = display(gt, hi)
    rmse_g = np.sqrt(np.mean((display(g_lo, lo) - GT) ** 2))
    rmse_o = np.sqrt(np.mean((display(o_lo, lo) - GT) ** 2))
    ratio = rmse_g / rmse_o
    bias = (display(g_hi, hi) - GT).mean(axis=(0, 1))
    lum = lambda a: (np.nan_to_num(a, nan=0.0).astype(np.float64) @ np.array([0.2126, 0.7152, 0.0722])).mean()
    rel_lum = abs(lum(g_hi) - lum(gt)) / lum(gt)
    print(f"{name}: RMSE gpu@{lo} {rmse_g:.4f} oracle@{lo} {rmse_o:.4f} ratio {ratio:.3f}; bias@{hi} {bias * 255}; lum diff {rel_lum:.4%}")
    assert 0.9 <= ratio <= 1.1
    assert np.all(np.abs(bias) < 0.5 / 255)
    assert rel_lum < 0.01
