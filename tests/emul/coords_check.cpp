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
        SkbGeom g;
        g.nz = sh[0]; g.ny = sh[1]; g.nx = sh[2]; g.ndim = 3; g.sz = sh[1] * sh[2]; g.nvox = sh[0] * g.sz; g.inv_sz = 1.0 / (double)g.sz;
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
    printf("%ld checked, %ld bad\n", n, bad);
    return bad != 0;
}
