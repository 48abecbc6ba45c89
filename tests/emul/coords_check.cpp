// exhaustive-ish check of the fp64 plane-index path of skb_coords against integer division
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "skb_platform.h"
#include "skb_items.cuh"
int main() {
    int64_t shapes[][3] = {{2048, 2048, 2048}, {70000, 300, 301}, {3, 46340, 46341}, {1 << 21, 1, 2047}, {5000, 1, 1000003}, {2147483647, 1, 3}, {4097, 1024, 1024}};
    uint64_t rng = 88172645463325252ull;
    long bad = 0, n = 0;
    for (auto& sh : shapes) {
        SkbGeom g = skb_geom_make(3, sh[0], sh[1], sh[2]);
        if (g.nvox <= 0xffffffffLL) { printf("shape too small\n"); continue; }
        for (int t = 0; t < 4000000; ++t) {
            rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
            int64_t r;
            int mode = t & 3;
            if (mode == 0) r = (int64_t)(rng % (uint64_t)g.nvox);
            else {  // plane boundaries +-2
                int64_t zz = (int64_t)(rng % (uint64_t)g.nz);
                r = zz * g.sz + (int64_t)(t % 5) - 2;
                if (r < 0) r = 0;
                if (r >= g.nvox) r = g.nvox - 1;
            }
            int32_t z, y, x;
            skb_coords(g, r, z, y, x);
            int64_t q = r / g.nx, ex = r - q * g.nx, q2 = q / g.ny, ey = q - q2 * g.ny;
            if (z != q2 || y != ey || x != ex) ++bad;
            ++n;
        }
    }
    // the invariant division of skb_udiv against the hardware's, for awkward divisors and dividends
    {
        const uint32_t ds[] = {1, 2, 3, 5, 7, 10, 255, 256, 257, 1000, 1094, 1536, 2047, 2048, 2049, 46340, 46341, 65535,
                               65536, 65537, 1000003, 0x7ffffffe, 0x7fffffff, 0x80000000u, 0x80000001u, 0xfffffffeu, 0xffffffffu};
        for (uint32_t d : ds) {
            const SkbDiv v = skb_div_make(d);
            for (int t = 0; t < 400000; ++t) {
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                uint32_t x;
                const int mode = t % 6;
                if (mode == 0) x = (uint32_t)rng;
                else if (mode == 1) x = (uint32_t)(rng % 70000);
                else {  // around multiples of d and the top of the range
                    const uint64_t k = (rng >> 8) % (0x100000000ull / d + 1);
                    const int64_t c = (int64_t)(k * d) + (int64_t)(t % 5) - 2;
                    x = (uint32_t)(c < 0 ? 0 : (c > 0xffffffffLL ? 0xffffffffLL - (t % 3) : c));
                }
                if (skb_udiv(x, v) != x / d) ++bad;
                ++n;
            }
        }
    }
    printf("%ld checked, %ld bad\n", n, bad);
    return bad != 0;
}
