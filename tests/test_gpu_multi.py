"""Multi-GPU inside the boundary (rtgpu_multi_render: spp sharding + one ncclReduce in librtgpu.so) and the native driver
rt_render, which is the reference's main.rs flow on top of it.  Needs a B200: -m gpu."""
import os
import subprocess

import numpy as np
import pytest

import rtb200 as rt

pytestmark = pytest.mark.gpu

PKG = os.path.dirname(os.path.abspath(rt.HOST_LIB_PATH))


def _devices():
    return rt.gpu_lib().rtgpu_device_count()


def test_multi_render_on_one_device_is_rtgpu_render():
    s = rt.Scene(seed=5).build_config("cornell_box", 1.0)
    low = s.lower()
    w = h = 64
    ref, st0 = rt.GpuScene(low, device=0).render(w, h, 0, 8, seed=5)
    a, st, ms = rt.multi_render(low, w, h, 0, 8, 1, seed=5)
    assert np.array_equal(a, ref) and st.rays == st0.rays and st.samples == w * h * 8 and ms == 0.0


@pytest.mark.skipif("_devices() < 2")
def test_multi_render_shards_samples_and_reduces():
    s = rt.Scene(seed=5).build_config("final_scene", 1.0)
    low = s.lower()
    w = h = 96
    ref, st0 = rt.GpuScene(low, device=0).render(w, h, 0, 16, seed=5)
    for g in (2, min(_devices(), 8)):
        a, st, ms = rt.multi_render(low, w, h, 0, 16, g, seed=5)
        assert st.rays == st0.rays and st.samples == st0.samples  # the same sample set, whoever renders it
        assert np.allclose(a, ref, rtol=1e-5, atol=1e-5)          # float sums in another order
        assert ms > 0.0


def test_native_driver_writes_the_same_image(tmp_path):
    """rt_render --devices 1 (and 2 when the box has them): byte-identical PPM to the Python path on one device."""
    exe = os.path.join(PKG, "rt_render")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    w, h, spp, seed = 64, 48, 8, 7
    s = rt.Scene(seed=seed).build_config("cornell_book3", w / h)
    gpu = rt.GpuScene(s.lower(), device=0)
    acc, _ = gpu.render(w, h, 0, spp, seed=seed)
    img = gpu.resolve(acc, w, h, spp)
    for g in [1] + ([2] if _devices() >= 2 else []):
        out = tmp_path / f"out{g}.ppm"
        r = subprocess.run([exe, "--scene", "cornell_book3", "--width", str(w), "--height", str(h), "--spp", str(spp), "--seed", str(seed),
                            "--devices", str(g), "--assets", rt.ASSET_DIR, "--out", str(out)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "All Work Done." in r.stdout
        data = out.read_bytes()
        header = f"P6\n{w} {h}\n255\n".encode()
        assert data.startswith(header)
        px = np.frombuffer(data[len(header):], dtype=np.uint8).reshape(h, w, 3)
        if g == 1:
            assert np.array_equal(px, img)
            # the reference's own output format: JPEG at quality 100 (main.rs:213-221)
            jpg = tmp_path / "out1.jpg"
            r = subprocess.run([exe, "--scene", "cornell_book3", "--width", str(w), "--height", str(h), "--spp", str(spp), "--seed", str(seed),
                                "--devices", "1", "--assets", rt.ASSET_DIR, "--out", str(jpg)], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            dj = np.abs(rt.decode_jpeg(str(jpg)).astype(int) - img.astype(int))
            assert dj.max() <= 6 and dj.mean() <= 1.0
        else:  # two float sums added in another order: at most one 8-bit level, on few pixels
            d = np.abs(px.astype(int) - img.astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 0.01
