/*
 * rt_rng.h -- counter-based Philox4x32-10 streams that replace `rand::thread_rng()`.
 *
 * The reference draws everything from an unseedable ChaCha12 thread-local generator
 * (rand 0.8.5; call sites listed in SURVEY.md 8c).  Here every consumer owns a
 * *stream*: key = 64-bit seed, counter = (block, s0, s1, domain).  A stream is a
 * sequence of 64-bit slots; slot i lives in Philox block i/2, words (2*(i&1), 2*(i&1)+1).
 * One slot feeds one `gen::<f64>()`, one `gen_range(a..b)` or one integer `gen_range`.
 *
 *   gen()            = (slot >> 11) * 2^-53               (rand's 53-bit [0,1) recipe)
 *   gen_range(a,b)   = a + (b - a) * gen()                (distribution-equal to rand's
 *                                                          52-bit [1,2)-1 recipe)
 *   gen_index(n)     = widening multiply with zone rejection (rand's UniformInt<u64>)
 *
 * Usable from host C++ and from CUDA device code (sm_100a).  The oracle carries its
 * own independent implementation of the same spec (oracle/oracle.cpp); both are
 * pinned by the Random123 known-answer vectors in tests/test_rng.py.
 */
#ifndef RT_RNG_H
#define RT_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

/* stream domains (4th counter word) */
#define RT_DOMAIN_PATH 0x50415448u  /* "PATH": s0 = pixel index, s1 = global sample index */
#define RT_DOMAIN_BATCH 0x42544348u /* "BTCH": dumped-ray batches, key = per-ray key       */
#define RT_DOMAIN_SCENE 0x53434e45u /* "SCNE": scene / BVH axis / Perlin construction      */

RT_HD uint32_t rt_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

RT_HD void rt_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                            uint32_t k1, uint32_t out[4]) {
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = rt_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = rt_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

struct rt_rng {
    uint32_t k0, k1;     /* seed */
    uint32_t s0, s1, s2; /* stream id + domain */
    uint32_t pos;        /* next slot */
    uint32_t blk;        /* cached block index, 0xffffffff = none */
    uint32_t w[4];

    RT_HD void init(uint64_t seed, uint32_t a, uint32_t b, uint32_t domain, uint32_t start = 0) {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        s0 = a;
        s1 = b;
        s2 = domain;
        pos = start;
        blk = 0xffffffffu;
    }
    RT_HD uint64_t next_u64() {
        uint32_t b = pos >> 1;
        if (b != blk) {
            rt_philox4x32_10(b, s0, s1, s2, k0, k1, w);
            blk = b;
        }
        uint32_t hi = (pos & 1u) ? w[2] : w[0];
        uint32_t lo = (pos & 1u) ? w[3] : w[1];
        ++pos;
        return ((uint64_t)hi << 32) | (uint64_t)lo;
    }
    /* The next THREE slots at once (random_in_unit_sphere draws x, y, z per try): the same values as three next_u64() calls,
     * but the Philox block that holds the later slots is computed at ONE place by every lane -- three separate calls are three
     * copies of the generator behind three parity-dependent branches, which a warp of mixed parities executes one after the other. */
    RT_HD void next3_u64(uint64_t &a, uint64_t &b, uint64_t &c) {
        const uint32_t b0 = pos >> 1;
        if (b0 != blk) rt_philox4x32_10(b0, s0, s1, s2, k0, k1, w);
        uint32_t v[4];
        rt_philox4x32_10(b0 + 1u, s0, s1, s2, k0, k1, v);
        if (pos & 1u) {
            a = ((uint64_t)w[2] << 32) | (uint64_t)w[3];
            b = ((uint64_t)v[0] << 32) | (uint64_t)v[1];
            c = ((uint64_t)v[2] << 32) | (uint64_t)v[3];
        } else {
            a = ((uint64_t)w[0] << 32) | (uint64_t)w[1];
            b = ((uint64_t)w[2] << 32) | (uint64_t)w[3];
            c = ((uint64_t)v[0] << 32) | (uint64_t)v[1];
        }
        w[0] = v[0]; w[1] = v[1]; w[2] = v[2]; w[3] = v[3];
        blk = b0 + 1u;
        pos += 3u;
    }
    static RT_HD double to_unit_f64(uint64_t u) { return (double)(u >> 11) * (1.0 / 9007199254740992.0); }
    /* rand: rng.gen::<f64>() */
    RT_HD double gen() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    /* rand: rng.gen_range(lo..hi) for f64 */
    RT_HD double gen_range(double lo, double hi) { return lo + (hi - lo) * gen(); }
    /* rand: rng.gen_range(0..n) for integers, n >= 1 */
    RT_HD uint32_t gen_index(uint32_t n) {
        uint64_t range = n;
#if defined(__CUDA_ARCH__)
        int lz = __clzll((long long)range); /* n >= 1 */
#elif defined(__GNUC__)
        int lz = __builtin_clzll(range);
#else
        int lz = 0;
        for (uint64_t t = range; !(t >> 63); t <<= 1) ++lz;
#endif
        uint64_t zone = (range << lz) - 1;
        for (;;) {
            uint64_t v = next_u64();
#if defined(__CUDA_ARCH__)
            uint64_t hi = __umul64hi(v, range);
#else
            uint64_t hi = (uint64_t)(((unsigned __int128)v * range) >> 64);
#endif
            uint64_t lo = v * range;
            if (lo <= zone) return (uint32_t)hi;
        }
    }
};

#endif /* RT_RNG_H */
