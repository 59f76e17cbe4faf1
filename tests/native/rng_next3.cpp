// rng_next3.cpp -- rt_rng::next3_u64 must hand out exactly the slots three next_u64() calls would, from any position and with any
// cache state, and leave the generator in a state from which the stream continues unchanged (include/rt_rng.h).
// Build: g++ -O2 -std=c++17 -o rng_next3 rng_next3.cpp
#include <cstdio>
#include "../../include/rt_rng.h"

int main() {
    long long checked = 0;
    for (uint32_t start = 0; start < 64; ++start)
        for (int warm = 0; warm < 4; ++warm)      /* draws before the triple: moves the cached block relative to `pos` */
            for (int triples = 1; triples <= 5; ++triples) {
                rt_rng a, b;
                a.init(0x1234567890abcdefull, 77u + start, 5u, RT_DOMAIN_PATH, start);
                b = a;
                for (int k = 0; k < warm; ++k) if (a.next_u64() != b.next_u64()) return 1;
                for (int t = 0; t < triples; ++t) {
                    uint64_t x, y, z;
                    a.next3_u64(x, y, z);
                    if (x != b.next_u64() || y != b.next_u64() || z != b.next_u64() || a.pos != b.pos) {
                        printf("mismatch at start %u warm %d triple %d\n", start, warm, t);
                        return 1;
                    }
                    ++checked;
                }
                for (int k = 0; k < 5; ++k) if (a.next_u64() != b.next_u64()) { printf("stream diverges after the triple\n"); return 1; }
            }
    printf("next3 ok: %lld triples\n", checked);
    return 0;
}
