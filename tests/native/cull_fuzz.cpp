// cull_fuzz.cpp -- pins raytracer-2022_b200/csrc/rt_cull.h (the f32 pre-test of AABB::hit) against the f64 test of
// hittable/bvh/aabb.rs:15-32 on random and adversarial ray/box pairs: a "certain" answer must never disagree.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -o cull_fuzz cull_fuzz.cpp ; run: ./cull_fuzz [millions] [seed]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../raytracer-2022_b200/csrc/rt_cull.h"

// AABB::hit exactly as upstream writes it (aabb.rs:15-32): per-axis division, swap, early exit
static bool aabb_hit_ref(const double mn[3], const double mx[3], const double o[3], const double d[3], double t_min, double t_max) {
    for (int i = 0; i < 3; ++i) {
        double inv_d = 1.0 / d[i];
        double t0 = (mn[i] - o[i]) * inv_d;
        double t1 = (mx[i] - o[i]) * inv_d;
        if (inv_d < 0.0) { double t = t0; t0 = t1; t1 = t; }
        t_min = t0 > t_min ? t0 : t_min;
        t_max = t1 < t_max ? t1 : t_max;
        if (t_max <= t_min) return false;
    }
    return true;
}

int main(int argc, char **argv) {
    long long n = (argc > 1 ? atoll(argv[1]) : 20) * 1000000LL;
    std::mt19937_64 rng(argc > 2 ? atoll(argv[2]) : 12345);
    std::uniform_real_distribution<double> U(0., 1.);
    auto sym = [&](double s) { return (2. * U(rng) - 1.) * s; };
    long long cnt[3] = {0, 0, 0}, bad = 0, refhit = 0;
    for (long long it = 0; it < n; ++it) {
        double scale = std::pow(10., sym(1.) * 3. + 1.);  // 1e-2 .. 1e4
        double mn[3], mx[3], o[3], d[3];
        for (int a = 0; a < 3; ++a) {
            double c = sym(scale), h = U(rng) * scale * 0.2;
            mn[a] = c - h; mx[a] = c + h;
            if ((it & 1023) == 7 && a == (int)(it >> 10) % 3) mx[a] = mn[a];  // zero-thickness box (Q15)
        }
        int mode = (int)(it % 8);
        for (int a = 0; a < 3; ++a) o[a] = sym(scale * (mode == 0 ? 30. : 1.5));
        if (mode == 1) for (int a = 0; a < 3; ++a) o[a] = mn[a] + (mx[a] - mn[a]) * U(rng);  // origin inside
        // target: a point of the box, often on a face / edge / corner, perturbed by a tiny relative amount
        double tgt[3];
        for (int a = 0; a < 3; ++a) {
            double u = U(rng);
            if (mode >= 2) { double r = U(rng); u = r < 0.4 ? 0. : (r < 0.8 ? 1. : u); }
            tgt[a] = mn[a] + (mx[a] - mn[a]) * u;
            if (mode >= 3) tgt[a] += sym(1.) * scale * std::pow(10., -3. - 12. * U(rng));
        }
        double len = std::pow(10., sym(1.) * 2.);
        for (int a = 0; a < 3; ++a) d[a] = (tgt[a] - o[a]) * len;
        if (mode == 5) d[(int)(it >> 3) % 3] *= std::pow(10., -6. * U(rng));   // nearly axis-parallel
        if (mode == 6 && (it & 64)) d[(int)(it >> 3) % 3] = 0.;               // 0 * inf NaN cases
        if (mode == 7) for (int a = 0; a < 3; ++a) d[a] = sym(1.);
        double t_min = (it & 16) ? 0.001 : ((it & 32) ? -INFINITY : U(rng) * 2.);
        double best = 1.7976931348623157e308;
        if (it & 4) {  // a running best near the entry / exit of the box
            double te = 0.;
            for (int a = 0; a < 3; ++a) { double t = (tgt[a] - o[a]) / d[a]; if (t == t && std::fabs(t) < 1e300) te = t; }
            best = te * (1. + sym(1.) * std::pow(10., -1. - 14. * U(rng)));
        }
        double inv[3] = {1. / d[0], 1. / d[1], 1. / d[2]};
        double bmax = 0.;
        for (int a = 0; a < 3; ++a) bmax = std::fmax(bmax, std::fmax(std::fabs(mn[a]), std::fabs(mx[a])));
        float extent = (float)std::fmax((double)RT_CULL_MIN_EXTENT, bmax * (1. + U(rng)));
        extent = rt_f32_up((double)extent);
        CullRay c;
        if ((it & 3) == 3) {
            /* the same ray set up for another t_min and then re-timed (cull_ray_retime: a ConstantMedium asks its boundary from -inf,
             * then from t1 + 0.0001, while the lane keeps its ray): the answer must hold for the t_min it was re-timed to */
            double other = (it & 128) ? 0.001 : ((it & 256) ? -INFINITY : U(rng) * 3.);
            cull_ray_setup(c, o[0], o[1], o[2], inv[0], inv[1], inv[2], other, extent);
            if (it & 512) cull_ray_retime(c, -INFINITY); /* two re-timings in a row, as in the two passes over a boundary */
            cull_ray_retime(c, t_min);
        } else {
            cull_ray_setup(c, o[0], o[1], o[2], inv[0], inv[1], inv[2], t_min, extent);
        }
        int r = cull_box(c, (float)mn[0], (float)mn[1], (float)mn[2], (float)mx[0], (float)mx[1], (float)mx[2], rt_f32_down(best), rt_f32_up(best));
        bool ref = aabb_hit_ref(mn, mx, o, d, t_min, best);
        cnt[r]++;
        refhit += ref;
        if ((r == RT_CULL_HIT && !ref) || (r == RT_CULL_MISS && ref)) {
            if (bad < 10) fprintf(stderr, "MISMATCH it=%lld mode=%d cull=%d ref=%d e2=%g\n", it, mode, r, (int)ref, (double)c.e2);
            ++bad;
        }
    }
    printf("pairs %lld  ref hits %lld  cull: miss %lld hit %lld unsure %lld (%.3f %%)  mismatches %lld\n", n, refhit, cnt[0], cnt[1], cnt[2],
           100. * cnt[2] / n, bad);
    return bad ? 1 : 0;
}
