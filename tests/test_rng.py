"""Pins the PRODUCT's counter-based generator (include/rt_rng.h, compiled into librthost.so and into the CUDA kernels)
against the Random123 known-answer vectors for Philox4x32-10 and against the recipes of rand 0.8.5 it restates.
(tests/test_oracle_kat.py pins the oracle's independent implementation of the same spec.)"""
import ctypes as C

import numpy as np

import rtb200 as rt
from oracle_lib import philox as oracle_philox


def product_philox(ctr, key):
    L = rt.host_lib()
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    out = (C.c_uint32 * 4)()
    L.rtg_philox4x32_10(c, k, out)
    return list(out)


def product_stream(seed, a, b, domain, start, n, index_n=7):
    L = rt.host_lib()
    L.rtg_rng_stream.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_uint32]
    slots = np.zeros(n, dtype=np.uint64)
    gen = np.zeros(n)
    rng = np.zeros(n)
    idx = np.zeros(n, dtype=np.uint32)
    L.rtg_rng_stream(seed, a, b, domain, start, n, slots.ctypes.data, gen.ctypes.data, rng.ctypes.data, idx.ctypes.data, index_n)
    return slots, gen, rng, idx


def test_product_philox_random123_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert product_philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert product_philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert product_philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_product_and_oracle_philox_agree_on_random_counters():
    rng = np.random.default_rng(7)
    for _ in range(200):
        ctr = [int(x) for x in rng.integers(0, 2**32, 4)]
        key = [int(x) for x in rng.integers(0, 2**32, 2)]
        assert product_philox(ctr, key) == oracle_philox(ctr, key).tolist()


def test_slot_layout_and_rand_recipes():
    seed, a, b, dom = 0x0123456789abcdef, 17, 4242, 0x50415448
    slots, gen, rng, idx = product_stream(seed, a, b, dom, 0, 64, index_n=7)
    # slot i lives in Philox block i/2, words (2*(i&1), 2*(i&1)+1): counter = (block, a, b, domain), key = seed
    for i in range(64):
        w = product_philox([i // 2, a, b, dom], [seed & 0xffffffff, seed >> 32])
        assert int(slots[i]) == (w[2 * (i & 1)] << 32) | w[2 * (i & 1) + 1]
    # rand 0.8.5 Standard f64: (u64 >> 11) * 2^-53, in [0, 1)
    assert np.array_equal(gen, (slots >> np.uint64(11)).astype(np.float64) * 2.0 ** -53)
    assert gen.min() >= 0.0 and gen.max() < 1.0
    # gen_range(lo, hi) = lo + (hi - lo) * gen()
    assert np.array_equal(rng, -1.0 + 2.0 * gen)
    # a stream resumed at slot k continues where it stopped (the wavefront kernels rely on it)
    s2, g2, _, _ = product_stream(seed, a, b, dom, 10, 20)
    assert np.array_equal(s2, slots[10:30]) and np.array_equal(g2, gen[10:30])


def test_gen_index_is_rands_widening_multiply_with_zone_rejection():
    seed, a, b, dom = 99, 1, 2, 0x53434e45
    for n in (1, 2, 3, 7, 11, 1000, 2**31 + 5):
        slots, _, _, _ = product_stream(seed, a, b, dom, 0, 600, index_n=n)
        _, _, _, idx = product_stream(seed, a, b, dom, 0, 100, index_n=n)
        # UniformInt<u64>::sample_single: zone = (range << lz) - 1; accept v when lo(v * range) <= zone, result = hi
        lz = 64 - int(n).bit_length()
        zone = ((n << lz) - 1) & (2**64 - 1)
        out, k = [], 0
        while len(out) < 100:
            v = int(slots[k]); k += 1
            m = v * n
            if (m & (2**64 - 1)) <= zone:
                out.append(m >> 64)
        assert idx.tolist() == out
        assert idx.max() < n


def test_next3_hands_out_the_same_slots_as_three_single_draws(tmp_path):
    """random_in_unit_sphere (vec.rs:69-76) takes its three draws per try through rt_rng::next3_u64 on the device (one converged
    Philox evaluation instead of three parity-dependent ones): same slots, same order, same generator state afterwards."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "rng_next3")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tests", "native", "rng_next3.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "next3 ok" in r.stdout, r.stdout + r.stderr
