"""Known-answer vectors K1-K12 against the CPU oracle.

The reference ships no tests; these vectors are derived by hand from its source (file:line in
SURVEY.md) and are the pins of the oracle ("parity unpinned" otherwise).
"""
import math

import numpy as np
import pytest

import rtb200 as rt
from oracle_lib import OracleScene, lib, make_rays, philox, write_color

F64_MAX = np.finfo(np.float64).max


def one_object_scene(build):
    s = rt.Scene(seed=1)
    obj = build(s)
    s.set_world(obj)
    return s, OracleScene(s)


def test_k1_k2_sphere():
    s, o = one_object_scene(lambda s: s.sphere((0, 0, 0), 1.0, s.lambertian((0.5, 0.5, 0.5))))
    h = o.trace_batch(make_rays([(0, 0, -3), (0, 0, 0)], [(0, 0, 1), (0, 0, 1)]))
    assert h["hit"].tolist() == [1, 1]
    assert h["t"][0] == 2.0 and np.array_equal(h["p"][0], [0, 0, -1]) and np.array_equal(h["n"][0], [0, 0, -1])
    assert h["front_face"][0] == 1 and h["u"][0] == 0.75 and h["v"][0] == 0.5
    assert h["t"][1] == 1.0 and h["front_face"][1] == 0 and np.array_equal(h["n"][1], [0, 0, -1])
    assert h["u"][1] == 0.25 and h["v"][1] == 0.5


def test_k3_xz_rect():
    s, o = one_object_scene(lambda s: s.xz_rect(0, 555, 0, 555, 555, s.lambertian((0.7, 0.7, 0.7))))
    h = o.trace_batch(make_rays([(278, 278, 278)], [(0, 1, 0)]))[0]
    assert h["hit"] == 1 and h["t"] == 277.0
    assert h["u"] == 0.5009009009009009 and h["v"] == 0.5009009009009009
    assert h["front_face"] == 0 and np.array_equal(h["n"], [0, -1, 0])


def test_k4_aabb():
    L = lib()
    bmin, bmax = np.zeros(3), np.ones(3)

    def hit(o, d):
        r = make_rays([o], [d])
        return L.orc_aabb_hit(bmin.ctypes.data, bmax.ctypes.data, r.ctypes.data)

    assert hit((-1, 0.5, 0.5), (1, 0, 0)) == 1
    assert hit((-1, 0.0, 0.5), (1, 0, 0)) == 1  # 0 * inf = NaN leaves the bound unchanged
    assert hit((-1, 1.5, 0.5), (1, 0, 0)) == 0


def test_k5_triangle():
    s, o = one_object_scene(lambda s: s.triangle((0, 0, 0), (1, 0, 0), (0, 1, 0), s.lambertian((0.7, 0.7, 0.7))))
    h = o.trace_batch(make_rays([(0.25, 0.25, 1)], [(0, 0, -1)]))[0]
    assert h["hit"] == 1 and h["t"] == 1.0 and np.array_equal(h["n"], [0, 0, 1]) and h["front_face"] == 1
    assert h["u"] == 0.25 and h["v"] == 0.25


def test_k6_reflectance():
    L = lib()
    assert L.orc_reflectance(1.0, 1 / 1.5) == 0.040000000000000015
    assert L.orc_reflectance(0.0, 1 / 1.5) == 1.0


def test_k7_write_color():
    assert write_color((25, 50, 100), 100).tolist() == [127, 181, 255]
    assert write_color((float("nan"), -1.0, 1e30), 1).tolist() == [0, 0, 255]


def test_k8_rect_pdf_value():
    s = rt.Scene(seed=1)
    light = s.xz_rect(213, 343, 127, 232, 554, s.diffuse_light((60, 60, 60)))
    s.set_world(light)
    s.set_lights([light])
    o = OracleScene(s)
    assert o.lights_pdf_value((278, 0, 179.5), (0, 554, 0)) == 22.484688644688646


def test_k9_sphere_pdf_value():
    s = rt.Scene(seed=1)
    sp = s.sphere((0, 10, 0), 1.0, s.diffuse_light((1, 1, 1)))
    s.set_world(sp)
    s.set_lights([sp])
    o = OracleScene(s)
    assert o.lights_pdf_value((0, 0, 0), (0, 1, 0)) == 31.75121120217509


def test_k10_rotate_y_constants():
    s = rt.Scene(seed=1)
    sp = s.sphere((0, 0, 0), 1.0, s.lambertian((0.5, 0.5, 0.5)))
    r = s.rotate_y(sp, 15.0)
    g = s.graph()
    assert g.objects[r].p[1] == 0.25881904510252074
    assert g.objects[r].p[2] == 0.9659258262890683


@pytest.mark.parametrize("n,nodes,depth,singletons", [(11, 13, 4, 3), (400, 511, 9, None), (1000, 1023, 10, 24),
                                                      (13079, 16383, 14, 3305)])
def test_k11_bvh_shape(n, nodes, depth, singletons):
    s = rt.Scene(seed=3)
    m = s.lambertian((0.5, 0.5, 0.5))
    rng = np.random.default_rng(0)
    ids = [s.sphere(tuple(rng.uniform(0, 100, 3)), 1.0, m) for _ in range(n)]
    root = s.bvh_node(ids, 0.0, 1.0)
    s.set_world(root)
    o = OracleScene(s)
    listing, boxes = o.bvh_listing(root)
    assert int((listing == -1).sum()) == nodes

    # depth and singleton count from the pre-order listing
    pos = 0
    n_single = 0

    def walk(d):
        nonlocal pos, n_single
        if listing[pos] != -1:
            pos += 1
            return d - 1, [listing[pos - 1]]
        pos += 1
        dl, ll = walk(d + 1)
        dr, lr = walk(d + 1)
        if len(ll) == 1 and len(lr) == 1 and ll[0] == lr[0]:
            n_single += 1
        return max(dl, dr), ll + lr

    import sys
    sys.setrecursionlimit(100000)
    dmax, leaves = walk(1)
    assert dmax == depth
    if singletons is not None:
        assert n_single == singletons
    assert set(leaves) == set(ids)


def test_k11_two_object_rule():
    # 2-object node: the LAST list element goes left iff its min[axis] is strictly smaller
    for seed in range(8):
        s = rt.Scene(seed=seed)
        m = s.lambertian((0.5, 0.5, 0.5))
        a = s.sphere((0, 0, 0), 1.0, m)
        b = s.sphere((5, 5, 5), 1.0, m)
        root = s.bvh_node([b, a])  # last = a has the smaller min on every axis
        g = s.graph()
        assert (g.objects[root].c0, g.objects[root].c1) == (a, b)
        root2 = s.bvh_node([a, b])  # last = b is not smaller -> (first, last)
        assert (g.objects[root2].c0, g.objects[root2].c1) == (a, b)
        c = s.sphere((0, 0, 0), 2.0, m)
        d = s.sphere((-2, -2, -2), 0.0, m)  # equal mins: not strictly smaller -> left = first popped second = ids[0]
        root3 = s.bvh_node([c, d])
        g = s.graph()
        assert (g.objects[root3].c0, g.objects[root3].c1) == (c, d)


def test_k12_constant_medium():
    s = rt.Scene(seed=1)
    boundary = s.sphere((360, 150, 145), 70.0, s.dielectric(1.5))
    med = s.constant_medium(boundary, 0.2, (0.2, 0.4, 0.9))
    s.set_world(med)
    o = OracleScene(s)
    n = 2000
    rays = make_rays([(360 - 200, 150, 145)] * n, [(1, 0, 0)] * n)
    keys = np.arange(n, dtype=np.uint64) + 1000
    h = o.trace_batch(rays, keys)
    L = lib()
    t_entry = 200.0 - 70.0
    for i in range(n):
        rnd = L.orc_stream_gen(int(keys[i]), 0, 0, 0x42544348, 0)
        dist = -5.0 * math.log(rnd)
        assert h["rng_pos"][i] == 1
        if dist <= 140.0:
            assert h["hit"][i] == 1
            assert abs(h["t"][i] - (t_entry + dist)) <= 1e-12 * (t_entry + dist)
            assert np.array_equal(h["n"][i], [1, 0, 0]) and h["front_face"][i] == 1 and h["u"][i] == 0 and h["v"][i] == 0
        else:
            assert h["hit"][i] == 0
    frac = h["hit"].mean()
    assert abs(frac - (1 - math.exp(-140 / 5))) < 1e-9 or frac > 0.99


def test_ring_known_answers():
    """Ring(r=10, t=1): annulus 9..11 in the plane y = 0 (ring.rs:22-52)"""
    s, o = one_object_scene(lambda s: s.ring(10.0, 1.0, s.lambertian((0.78, 0.78, 0.78))))
    h = o.trace_batch(make_rays([(10, 5, 0), (8.5, 5, 0), (9, 5, 0), (0, -2, 10.5), (10, 5, 0)],
                                [(0, -1, 0), (0, -1, 0), (0, -1, 0), (0, 1, 0), (1, 0, 0)]))
    assert h["hit"].tolist() == [1, 0, 1, 1, 0]  # inside, too close, on the inner edge (inclusive), from below, parallel
    assert h["t"][0] == 5.0 and np.array_equal(h["n"][0], [0, 1, 0]) and h["front_face"][0] == 1
    assert h["t"][3] == 2.0 and np.array_equal(h["n"][3], [0, -1, 0]) and h["front_face"][3] == 0
    assert np.allclose(s.bounding_box(s.graph().world), [-11, -0.0001, -11, 11, 0.0001, 11])


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32-10
    assert philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox([0xffffffff] * 4, [0xffffffff] * 2).tolist() == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]).tolist() == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_camera_derivation_matches_host():
    import ctypes as C
    s = rt.Scene(seed=1)
    s.set_camera((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16 / 9, 0.1, 10.0, 0.0, 1.0)
    g = s.graph()
    cam = rt.D.rt_camera.from_buffer_copy(g.camera)
    for f in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w"):
        getattr(cam, f)[:] = [0.0, 0.0, 0.0]
    cam.lens_radius = 0.0
    lib().orc_camera_derive(C.byref(cam))
    for f in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w"):
        assert list(getattr(cam, f)) == list(getattr(g.camera, f)), f
    assert cam.lens_radius == g.camera.lens_radius == 0.05
