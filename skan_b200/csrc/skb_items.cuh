// skan_b200 -- per-item device functions of the node/edge/path proportional stages.
//
// Each functor is "one logical thread per item" and is launched by skb_foreach (a grid-stride
// CUDA kernel on sm_100a; a serial loop in the host emulation harness used by CPU unit tests).
// Reference citations are relative to /root/reference (jni/skan).
//
// Data model (everything addressed by raster-order node id, never by a V-sized id image):
//   bits    packed foreground mask, 1 bit / voxel in C raster order, little-endian inside u64
//           words, zero pad words before word 0 and after the last word
//   pre256  exclusive popcount prefix of `bits` at every 256-bit block  -> rank(r) = node id
//   lin     raveled voxel index of node i (int64, ascending)                       csr.py:131
//   nb      27-bit raw neighbourhood mask of node i; bit k27 = (dz+1)*9+(dy+1)*3+(dx+1),
//           centre bit 13 always 0; ascending bit order == ascending neighbour id   csr.py:122-144
//   keep    nb minus the junction-junction edges dropped by the junction MST        csr.py:980-1020
#pragma once
#include "skb_platform.h"

// ---- element types of the skeleton image ------------------------------------------------
enum {
    SKB_DT_U8 = 0,  // also numpy/torch bool
    SKB_DT_I8 = 1,
    SKB_DT_U16 = 2,
    SKB_DT_I16 = 3,
    SKB_DT_U32 = 4,
    SKB_DT_I32 = 5,
    SKB_DT_U64 = 6,
    SKB_DT_I64 = 7,
    SKB_DT_F32 = 8,
    SKB_DT_F64 = 9,
    SKB_DT_COUNT = 10
};

SKB_HD int skb_dtype_size(int dt) {
    switch (dt) {
        case SKB_DT_U8: case SKB_DT_I8: return 1;
        case SKB_DT_U16: case SKB_DT_I16: return 2;
        case SKB_DT_U32: case SKB_DT_I32: case SKB_DT_F32: return 4;
        default: return 8;
    }
}

// value of voxel r as fp64 (csr.py:554-562: ints -> float64; floats are widened losslessly)
SKB_HD double skb_load_value(const void* img, int dt, int64_t r) {
    switch (dt) {
        case SKB_DT_U8: return (double)((const uint8_t*)img)[r];
        case SKB_DT_I8: return (double)((const int8_t*)img)[r];
        case SKB_DT_U16: return (double)((const uint16_t*)img)[r];
        case SKB_DT_I16: return (double)((const int16_t*)img)[r];
        case SKB_DT_U32: return (double)((const uint32_t*)img)[r];
        case SKB_DT_I32: return (double)((const int32_t*)img)[r];
        case SKB_DT_U64: return (double)((const uint64_t*)img)[r];
        case SKB_DT_I64: return (double)((const int64_t*)img)[r];
        case SKB_DT_F32: return (double)((const float*)img)[r];
        default: return ((const double*)img)[r];
    }
}

// image[i] - image[n] evaluated in the image dtype (csr.py:1024: unsigned types wrap), as fp64.
// `a`, `b` are the fp64 copies of the two voxel values (exact for every type up to 53 bits).
SKB_HD double skb_height_diff(int dt, double a, double b) {
    switch (dt) {
        case SKB_DT_U8: return (double)(uint8_t)((uint8_t)a - (uint8_t)b);
        case SKB_DT_I8: return (double)(int8_t)((int8_t)a - (int8_t)b);
        case SKB_DT_U16: return (double)(uint16_t)((uint16_t)a - (uint16_t)b);
        case SKB_DT_I16: return (double)(int16_t)((int16_t)a - (int16_t)b);
        case SKB_DT_U32: return (double)(uint32_t)((uint32_t)a - (uint32_t)b);
        case SKB_DT_I32: return (double)(int32_t)((uint32_t)(int32_t)a - (uint32_t)(int32_t)b);
        case SKB_DT_U64: return (double)(uint64_t)((uint64_t)a - (uint64_t)b);
        case SKB_DT_I64: return (double)(int64_t)((uint64_t)(int64_t)a - (uint64_t)(int64_t)b);
        case SKB_DT_F32: return (double)((float)a - (float)b);
        default: return a - b;
    }
}

// np.hypot == glibc hypot (sysdeps/ieee754/dbl-64/e_hypot.c, 2.35+, non-FMA kernel); restated so
// that height-mode edge weights (csr.py:1025) carry the same bits as the reference's.
// Built with -fmad=false / -ffp-contract=off: every operation below rounds once.
SKB_HD double skb_hypot(double x, double y) {
    x = x < 0 ? -x : x;
    y = y < 0 ? -y : y;
    if (!(x == x) || !(y == y)) return x + y;
    double ax = x < y ? y : x;
    double ay = x < y ? x : y;
    const double EPS = 5.5511151231257827e-17;  // 0x1p-54
    if (ax >= ay / EPS) return ax + ay;
    double h = skb_sqrt(ax * ax + ay * ay);
    double t1, t2;
    if (h <= 2.0 * ay) {
        double delta = h - ay;
        t1 = ax * (2.0 * delta - ax);
        t2 = (delta - 2.0 * (ax - ay)) * delta;
    } else {
        double delta = h - ax;
        t1 = 2.0 * delta * (ax - 2.0 * ay);
        t2 = (4.0 * delta - ay) * ay + delta * delta;
    }
    h -= (t1 + t2) / (2.0 * h);
    return h;
}

// ---- geometry ------------------------------------------------------------------------------
// Every image is embedded as (nz, ny, nx) with leading extents 1 for ndim < 3.
// exact unsigned 32-bit division by a divisor fixed per launch (Granlund & Montgomery 1994, figure 4.1): a multiply-high,
// a subtraction and two shifts instead of the ~20 instructions of a hardware-assisted division -- the node emission and
// the stencil turn a raveled index into coordinates once per foreground voxel
struct SkbDiv {
    uint32_t m, s1, s2, d;
};
SKB_HD SkbDiv skb_div_make(uint32_t d) {
    SkbDiv v;
    uint32_t l = 0;
    while (l < 32 && ((uint64_t)1 << l) < (uint64_t)d) ++l;  // ceil(log2 d)
    v.m = (uint32_t)(((((uint64_t)1 << l) - d) << 32) / d + 1);
    v.s1 = l < 1 ? l : 1;
    v.s2 = l < 1 ? 0 : l - 1;
    v.d = d;
    return v;
}
SKB_HD uint32_t skb_udiv(uint32_t n, const SkbDiv& v) {
    const uint32_t t1 = skb_umulhi(v.m, n);
    return (t1 + ((n - t1) >> v.s1)) >> v.s2;
}

struct SkbGeom {
    int64_t nvox;
    int64_t sz;  // ny * nx
    int32_t nz, ny, nx, ndim;
    double inv_sz;  // 1.0 / sz
    SkbDiv dx, dy;  // division by nx, by ny
};

SKB_HD SkbGeom skb_geom_make(int ndim, int64_t nz, int64_t ny, int64_t nx) {
    SkbGeom g;
    g.nz = (int32_t)nz;
    g.ny = (int32_t)ny;
    g.nx = (int32_t)nx;
    g.ndim = ndim;
    g.sz = ny * nx;
    g.inv_sz = 1.0 / (double)g.sz;
    g.nvox = nz * ny * nx;
    g.dx = skb_div_make((uint32_t)nx);
    g.dy = skb_div_make((uint32_t)ny);
    return g;
}

SKB_HD void skb_coords(const SkbGeom& g, int64_t r, int32_t& z, int32_t& y, int32_t& x) {
    if (g.nvox <= 0xffffffffLL) {
        uint32_t rr = (uint32_t)r;
        uint32_t q = skb_udiv(rr, g.dx);
        x = (int32_t)(rr - q * (uint32_t)g.nx);
        uint32_t q2 = skb_udiv(q, g.dy);
        y = (int32_t)(q - q2 * (uint32_t)g.ny);
        z = (int32_t)q2;
    } else if (g.sz <= 0x7fffffffLL && r >= 0 && r < (1LL << 52)) {
        // more than 2^32 voxels, planes below 2^31: the plane index from one fp64 product (off by at most one,
        // corrected below: r < 2^52 and a quotient below 2^31 leave the error far under 1), the rest in 32 bits
        // instead of two 64-bit integer divisions per voxel
        int64_t q2 = (int64_t)((double)r * g.inv_sz);
        int64_t rem = r - q2 * g.sz;
        if (rem < 0) {
            --q2;
            rem += g.sz;
        } else if (rem >= g.sz) {
            ++q2;
            rem -= g.sz;
        }
        uint32_t rr = (uint32_t)rem;
        uint32_t q = skb_udiv(rr, g.dx);
        x = (int32_t)(rr - q * (uint32_t)g.nx);
        y = (int32_t)q;
        z = (int32_t)q2;
    } else {
        int64_t q = r / g.nx;
        x = (int32_t)(r - q * g.nx);
        int64_t q2 = q / g.ny;
        y = (int32_t)(q - q2 * g.ny);
        z = (int32_t)q2;
    }
}

// bits p, p+1, p+2 of the mask (p >= -1; the pad words make p = -1 and p+2 >= nvox readable)
SKB_HD uint32_t skb_win3(const uint64_t* bits, int64_t p) {
    // one aligned 32-bit load (ncu, round 2: the stencil was throttled by its 18 single-byte loads per voxel); the
    // window straddles two words for 2 positions in 32 only
    const uint32_t* b32 = (const uint32_t*)bits;
    const int64_t q = p >> 5;
    const uint32_t sh = (uint32_t)(p & 31);
    uint32_t v = b32[q] >> sh;
    if (sh > 29u) v |= b32[q + 1] << (32u - sh);
    return v & 7u;
}

// number of foreground voxels with raveled index < r  (== node id of voxel r when it is set)
SKB_HD int32_t skb_rank(const uint64_t* bits, const uint32_t* pre256, int64_t r) {
    int64_t blk = r >> 8;
    uint32_t acc = pre256[blk];
    const uint64_t* w = bits + blk * 4;
    int wi = (int)((r >> 6) & 3);
    for (int j = 0; j < wi; ++j) acc += (uint32_t)skb_popcll(w[j]);
    uint64_t low = w[wi] & ((1ull << (r & 63)) - 1ull);
    return (int32_t)(acc + (uint32_t)skb_popcll(low));
}

// calls fn(k27, neighbour_id) for every set bit of mask m of node i, ascending k27.
// One iteration per set bit (typically two) keeps the lanes of a warp in step; a neighbour in
// another row costs one rank query, neighbours in the node's own row are i-1 / i+1.
// raveled index of the neighbour in direction k27 of voxel r
SKB_HD int64_t skb_neighbor_lin(const SkbGeom& g, int64_t r, int k) {
    int row = k / 3;
    int dz = row / 3 - 1, dy = row % 3 - 1, dx = k - row * 3 - 1;
    return r + dz * g.sz + dy * (int64_t)g.nx + dx;
}

// skb_rank without data-dependent control flow: the prefix and the four words of the 256-bit block are loaded
// unconditionally, so that the loads of two queries can be in flight together
struct SkbRankLoad {
    uint32_t pre;
    uint64_t w[4];
};
SKB_HD SkbRankLoad skb_rank_load(const uint64_t* bits, const uint32_t* pre256, int64_t r) {
    SkbRankLoad q;
    int64_t blk = r >> 8;
    q.pre = pre256[blk];
    const uint64_t* w = bits + blk * 4;
    q.w[0] = w[0];
    q.w[1] = w[1];
    q.w[2] = w[2];
    q.w[3] = w[3];
    return q;
}
SKB_HD int32_t skb_rank_finish(const SkbRankLoad& q, int64_t r) {
    const int wi = (int)((r >> 6) & 3);
    const uint64_t low = (1ull << (r & 63)) - 1ull;
    uint32_t acc = q.pre;
    for (int j = 0; j < 4; ++j) {
        uint64_t mask = j < wi ? ~0ull : (j == wi ? low : 0ull);
        acc += (uint32_t)skb_popcll(q.w[j] & mask);
    }
    return (int32_t)acc;
}

// one neighbour per iteration: for the junction stages, where one lane in four has work at all and the paired
// form below only costs registers (ncu: 32 -> 48 registers and 0.070 -> 0.089 ms for the junction edge emission)
template <typename Fn>
SKB_HD void skb_for_neighbors_single(const SkbGeom& g, const uint64_t* bits, const uint32_t* pre256, int64_t r,
                                     int32_t i, uint32_t m, Fn& fn) {
    while (m) {
        int k = skb_ffs(m) - 1;
        m &= m - 1u;
        int32_t id;
        if (k == 12) id = i - 1;
        else if (k == 14) id = i + 1;
        else id = skb_rank(bits, pre256, skb_neighbor_lin(g, r, k));
        fn(k, id);
    }
}

template <typename Fn>
SKB_HD void skb_for_neighbors(const SkbGeom& g, const uint64_t* bits, const uint32_t* pre256, int64_t r, int32_t i,
                              uint32_t m, Fn& fn) {
    // two neighbours per iteration (a chain voxel has two): both rank queries are issued before either result is
    // used, one memory round trip instead of two
    while (m) {
        const int k0 = skb_ffs(m) - 1;
        m &= m - 1u;
        const int k1 = m ? skb_ffs(m) - 1 : -1;
        if (k1 >= 0) m &= m - 1u;
        const bool q0 = k0 != 12 && k0 != 14, q1 = k1 >= 0 && k1 != 12 && k1 != 14;
        int64_t r0 = 0, r1 = 0;
        SkbRankLoad l0, l1;
        if (q0) {
            r0 = skb_neighbor_lin(g, r, k0);
            l0 = skb_rank_load(bits, pre256, r0);
        }
        if (q1) {
            r1 = skb_neighbor_lin(g, r, k1);
            l1 = skb_rank_load(bits, pre256, r1);
        }
        const int32_t id0 = q0 ? skb_rank_finish(l0, r0) : (k0 == 12 ? i - 1 : i + 1);
        int32_t id1 = 0;
        if (k1 >= 0) id1 = q1 ? skb_rank_finish(l1, r1) : (k1 == 12 ? i - 1 : i + 1);
        fn(k0, id0);
        if (k1 >= 0) fn(k1, id1);
    }
}

// an item whose range ends at a count that lives in device memory: launched over a capacity, does nothing beyond *n
template <typename F>
struct SkbBounded {
    F f;
    const int32_t* n;
    SKB_HD void operator()(int64_t i) const {
        if (i < (int64_t)*n) f(i);
    }
};

// ---- stage: raw 3^d neighbourhood mask (csr.py:122-144, evaluated per foreground voxel) ----
// 27-bit mask of the foreground neighbours of voxel r = (z, y, x) (local coordinates of this volume / slab)
SKB_HD uint32_t skb_raw_nb_mask(const SkbGeom& g, const uint64_t* bits, int64_t r, int32_t z, int32_t y, int32_t x,
                                int32_t planes_independent) {
    uint32_t xm = 7u;
    if (x == 0) xm &= ~1u;
    if (x == g.nx - 1) xm &= ~4u;
    uint32_t m = 0;
    for (int dz = -1; dz <= 1; ++dz) {
        if ((uint32_t)(z + dz) >= (uint32_t)g.nz || (planes_independent && dz != 0)) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            if ((uint32_t)(y + dy) >= (uint32_t)g.ny) continue;
            int64_t rr = r + dz * g.sz + dy * (int64_t)g.nx;
            uint32_t w = skb_win3(bits, rr - 1) & xm;
            m |= w << (((dz + 1) * 3 + (dy + 1)) * 3);
        }
    }
    return m & ~(1u << 13);
}

struct SkbRawNbItem {
    SkbGeom g;
    const uint64_t* bits;
    const int64_t* lin;
    uint32_t* nb;
    int32_t planes_independent;  // 1: the leading axis indexes independent 2-D images (no dz neighbours)
    SKB_HD void operator()(int64_t i) const {
        int64_t r = lin[i];
        int32_t z, y, x;
        skb_coords(g, r, z, y, x);
        nb[i] = skb_raw_nb_mask(g, bits, r, z, y, x, planes_independent);
    }
};

// pixel_graph(connectivity=c) with c < ndim (csr.py:51-158): keep the neighbour directions with at most c
// non-zero offset components (scipy.ndimage.generate_binary_structure); `allowed` has one bit per k27
struct SkbAndMaskItem {  // over nodes
    uint32_t* nb;
    uint32_t allowed;
    SKB_HD void operator()(int64_t i) const { nb[i] &= allowed; }
};

// make_degree_image (csr.py:1173-1208): the raw 3^d-neighbourhood degree of every skeleton pixel, as an image
struct SkbDegreeImageItem {  // over nodes
    const uint32_t* nb;
    const int64_t* lin;
    int64_t* image;  // V entries, zeroed by the caller
    SKB_HD void operator()(int64_t i) const { image[lin[i]] = (int64_t)skb_popc(nb[i]); }
};

// ---- stage: junction-junction edges (csr.py:1004-1016) -------------------------------------
// Edge weight of neighbour offset k27: table distance (nputil.py:278-279, host-computed) or, in
// height mode, hypot(height difference, distance) (csr.py:148-151, 1023-1025).
struct SkbWeight {
    const double* wtab;    // 27 entries
    const double* values;  // node values (fp64) or null
    int32_t height;        // 0 / 1
    int32_t dtype;
    SKB_HD double operator()(int32_t i, int32_t u, int k) const {
        double w = wtab[k];
        if (height) w = skb_hypot(skb_height_diff(dtype, values[i], values[u]), w);
        return w;
    }
};

struct SkbJuncCountFn {
    const uint32_t* nb;
    uint32_t count;
    SKB_HD void operator()(int k, int32_t u) {
        if (k > 13 && skb_popc(nb[u]) >= 3) ++count;
    }
};

struct SkbJuncCountItem {
    SkbGeom g;
    const uint64_t* bits;
    const uint32_t* pre256;
    const int64_t* lin;
    const uint32_t* nb;
    int32_t a0, a1;  // only nodes in [a0, a1) emit their forward edges (the nodes a Z-slab owns)
    uint32_t* count;
    SKB_HD void operator()(int64_t i) const {
        uint32_t m = nb[i];
        uint32_t c = 0;
        if ((int32_t)i >= a0 && (int32_t)i < a1 && skb_popc(m) >= 3 && (m >> 14)) {
            SkbJuncCountFn fn{nb, 0u};
            skb_for_neighbors_single(g, bits, pre256, lin[i], (int32_t)i, m & ~((1u << 14) - 1u), fn);
            c = fn.count;
        }
        count[i] = c;
    }
};

struct SkbJuncEmitFn {
    const uint32_t* nb;
    SkbWeight wf;
    int32_t a;
    uint32_t o;
    int32_t* ea;
    int32_t* eb;
    uint8_t* ek;
    uint64_t* ew;
    uint32_t cap;  // capacity of the edge arrays (they may be sized by a hint): nothing is written beyond it
    SKB_HD void operator()(int k, int32_t u) {
        if (k > 13 && skb_popc(nb[u]) >= 3) {
            if (o < cap) {
                ea[o] = a;
                eb[o] = u;
                ek[o] = (uint8_t)k;
                ew[o] = skb_f64_bits(wf(a, u, k));
            }
            ++o;
        }
    }
};

struct SkbJuncEmitItem {
    SkbGeom g;
    const uint64_t* bits;
    const uint32_t* pre256;
    const int64_t* lin;
    const uint32_t* nb;
    const uint32_t* offset;  // exclusive scan of SkbJuncCountItem counts
    SkbWeight wf;
    int32_t* ea;
    int32_t* eb;
    uint8_t* ek;
    uint64_t* ew;
    uint32_t cap = 0xffffffffu;
    SKB_HD void operator()(int64_t i) const {
        uint32_t m = nb[i];
        if (skb_popc(m) >= 3 && (m >> 14) && offset[i + 1] != offset[i]) {
            SkbJuncEmitFn fn{nb, wf, (int32_t)i, offset[i], ea, eb, ek, ew, cap};
            skb_for_neighbors_single(g, bits, pre256, lin[i], (int32_t)i, m & ~((1u << 14) - 1u), fn);
        }
    }
};

// ---- stage: junction MST, Boruvka under the strict order (weight, min id, max id) ----------
// The reference keeps, inside every cluster of raw-degree>=3 voxels, the edges of
// scipy.sparse.csgraph.minimum_spanning_tree (csr.py:1017): Kruskal over a stable sort by
// weight of CSR-ordered entries == the unique MST under (weight, min id, max id).  Junction
// edges are emitted in (a, k27) order with a < b, so the edge index is the tie-break.
SKB_HD int32_t skb_find(int32_t* par, int32_t v) {
    int32_t p = skb_ld_cg(par + v);
    while (p != v) {
        int32_t gp = skb_ld_cg(par + p);
        if (gp != p) skb_st_cg(par + v, gp);  // path halving; any ancestor is a valid parent
        v = p;
        p = gp;
    }
    return v;
}

struct SkbMstInitItem {  // over junction edges: only the entries of edge endpoints are ever read
    const int32_t* ea;
    const int32_t* eb;
    int32_t* par;
    uint64_t* bestw;
    uint32_t* beste;
    SKB_HD void operator()(int64_t e) const {
        int32_t a = ea[e], b = eb[e];
        if (a < 0) return;  // padding row of a gathered edge list (the caller marks it dead)
        par[a] = a;
        par[b] = b;
        bestw[a] = ~0ull;
        bestw[b] = ~0ull;
        beste[a] = ~0u;
        beste[b] = ~0u;
    }
};

struct SkbMstDeadItem {  // over junction edges: padding rows (a < 0) never take part
    const int32_t* ea;
    int32_t* era;
    SKB_HD void operator()(int64_t e) const { era[e] = ea[e] < 0 ? -2 : 0; }
};

struct SkbMstMinWItem {  // over junction edges: phase 1, lightest weight leaving each component
    const int32_t* ea;
    const int32_t* eb;
    const uint64_t* ew;
    int32_t* par;
    int32_t* era;
    int32_t* erb;
    uint64_t* bestw;
    int32_t* flag;
    SKB_HD void operator()(int64_t e) const {
        if (era[e] == -2) return;  // already inside one component
        int32_t ra = skb_find(par, ea[e]);
        int32_t rb = skb_find(par, eb[e]);
        if (ra == rb) {
            era[e] = -2;
            return;
        }
        era[e] = ra;
        erb[e] = rb;
        skb_atomic_min_u64(bestw + ra, ew[e]);
        skb_atomic_min_u64(bestw + rb, ew[e]);
        skb_raise(flag);
    }
};

struct SkbMstMinEItem {  // phase 2: among the lightest, the smallest (a, b) == smallest edge index
    const uint64_t* ew;
    const int32_t* era;
    const int32_t* erb;
    const uint64_t* bestw;
    uint32_t* beste;
    SKB_HD void operator()(int64_t e) const {
        int32_t ra = era[e];
        if (ra == -2) return;
        int32_t rb = erb[e];
        uint64_t w = ew[e];
        if (bestw[ra] == w) skb_atomic_min_u32(beste + ra, (uint32_t)e);
        if (bestw[rb] == w) skb_atomic_min_u32(beste + rb, (uint32_t)e);
    }
};

struct SkbMstHookItem {  // phase 3: chosen edges join the forest; hook one root under the other
    const int32_t* era;
    const int32_t* erb;
    const uint32_t* beste;
    int32_t* par;
    uint8_t* tree;
    SKB_HD void operator()(int64_t e) const {
        int32_t ra = era[e];
        if (ra == -2) return;
        int32_t rb = erb[e];
        bool pa = beste[ra] == (uint32_t)e;
        bool pb = beste[rb] == (uint32_t)e;
        if (!(pa || pb)) return;
        tree[e] = 1;
        // a component's root is written by exactly one edge (its chosen one); a mutual choice
        // keeps the smaller root so that the two hooks do not form a 2-cycle
        if (pa && !(pb && ra < rb)) skb_st_cg(par + ra, rb);
        if (pb && !(pa && rb < ra)) skb_st_cg(par + rb, ra);
    }
};

struct SkbMstResetItem {  // phase 4: clear the per-component minima touched this round
    const int32_t* era;
    const int32_t* erb;
    uint64_t* bestw;
    uint32_t* beste;
    SKB_HD void operator()(int64_t e) const {
        int32_t ra = era[e];
        if (ra == -2) return;
        int32_t rb = erb[e];
        bestw[ra] = ~0ull;
        bestw[rb] = ~0ull;
        beste[ra] = ~0u;
        beste[rb] = ~0u;
    }
};

struct SkbMstApplyItem {  // drop non-tree junction-junction edges in both directions (csr.py:1018-1019)
    const int32_t* ea;
    const int32_t* eb;
    const uint8_t* ek;
    const uint8_t* tree;
    int32_t id_shift;  // edge endpoints are global ids; keep[] is indexed by local id = global - id_shift
    int32_t n_local;   // endpoints that are not nodes of this slab are skipped
    uint32_t* keep;
    SKB_HD void operator()(int64_t e) const {
        if (tree[e] || ea[e] < 0) return;
        int k = ek[e];
        int32_t a = ea[e] - id_shift, b = eb[e] - id_shift;
        if (a >= 0 && a < n_local) skb_atomic_and_u32(keep + a, ~(1u << k));
        if (b >= 0 && b < n_local) skb_atomic_and_u32(keep + b, ~(1u << (26 - k)));
    }
};

// ---- stage: path tracing (csr.py:347-446) as sampled list ranking ----------------------------
// The serial walk of the reference follows degree-2 chains between terminals (degree 1 or >= 3).
// Here every half-edge leaving a terminal, and both half-edges leaving every "splitter" chain
// node (a 1/32 hash sample of the degree-2 nodes), starts a bounded local walk to the next
// terminal or splitter; the much shorter list of walk starts is then ranked by pointer jumping,
// which gives every start its path, its distance to the path's end and the minimum node id
// seen; a second local walk writes the path entries.  Junction-free cycles (csr.py:369-386)
// are the degree-2 nodes no terminal-started list covers; their minimum node becomes a
// terminal ("cycle root") and the same machinery runs again on them (phase 1).
struct alignas(16) skb_i4 {
    int32_t x, y, z, w;
};
#define SKB_REC_DEG_MASK 0xff
#define SKB_REC_CYCLE_ROOT 0x100
#define SKB_REC_SPLITTER 0x200
#define SKB_DEAD ((int32_t)0x80000000)  // a list that never reaches a terminal (cycle through splitters)

SKB_HD skb_i4 skb_ld_i4(const skb_i4* p) {
#if defined(__CUDA_ARCH__)
    int4 v = *reinterpret_cast<const int4*>(p);
    skb_i4 r;
    r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
#else
    return *p;
#endif
}
struct alignas(16) skb_d2 {
    double x, y;
};
SKB_HD skb_d2 skb_ld_d2(const skb_i4* p) {  // the second half of a 32-byte node record: two edge weights
#if defined(__CUDA_ARCH__)
    double2 v = *reinterpret_cast<const double2*>(p);
    skb_d2 r;
    r.x = v.x; r.y = v.y;
    return r;
#else
    skb_d2 r;
    memcpy(&r, p, 16);
    return r;
#endif
}
SKB_HD void skb_st_d2(skb_i4* p, double x, double y) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<double2*>(p) = make_double2(x, y);
#else
    skb_d2 r;
    r.x = x; r.y = y;
    memcpy(p, &r, 16);
#endif
}

SKB_HD void skb_st_i4(skb_i4* p, skb_i4 v) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<int4*>(p) = make_int4(v.x, v.y, v.z, v.w);
#else
    *p = v;
#endif
}

// 1/32 hash sample; the key is the node id, or the global raveled voxel index when a volume is
// split into Z-slabs (every rank must take the same decision for the same voxel)
// (shift = 32 - number of hash bits that must be zero: 27 = the default 1/32 sample)
SKB_HD bool skb_hash_splitter(int64_t key, uint32_t shift = 27u) {
    uint32_t h = (uint32_t)key ^ (uint32_t)(key >> 32);
    return ((h * 2654435761u) >> shift) == 0u;
}

// Z-slab decomposition of one volume over several GPUs (all-zero / full range on a single GPU).
// Local node ids index this rank's extended slab (owned planes + halo planes).
struct SkbSlab {
    int32_t own0, own1;          // owned nodes [own0, own1)
    int32_t cnt0, cnt1;          // nodes whose walk starts are numbered: owned + one halo plane on either side
    int32_t ga0, ga1, gb0, gb1;  // degree-2 nodes in these ranges are forced splitters (slab boundary planes)
    int32_t id_shift;            // global node id = local id + id_shift
    int32_t start_shift;         // global start id = startoff[v] + j + start_shift
    int32_t start_lo, start_hi;  // global ids of the starts leaving owned nodes
    int64_t lin_offset;          // global raveled index = lin[v] + lin_offset
};

// per-start payload carried through the ranking when results must be combined across slabs:
// running sums along the list and a description of the node the list ends at
struct alignas(16) SkbAux {
    double sw, sx, sxx;  // sum of edge weights, of node values, of squared node values
    int64_t end_lin;     // global raveled index of the end node
    int32_t end_node;    // its global node id
    int32_t end_deg;     // its degree
    int32_t end_key;     // global id of its first start (a dense key for the end node)
    int32_t pad;
    double end_value;    // its pixel value (the extra coordinate of value_is_height, csr.py:932-948)
    double pad2;         // (the records stay 16-byte aligned rows of the exchange arenas)
};

// Node record, 32 bytes = two 16-byte halves of rec[2v], rec[2v+1]:
//   (n0, n1, indptr[v], degree | flags)  and  (w0, w1) = weights of the edges to n0 and n1
// (n0, n1, w0, w1 only meaningful for degree-2 nodes).  One sector per walk step.
struct SkbRecOut {
    const int64_t* lin;  // null: splitters are sampled by node id
    SkbSlab sl;
    skb_i4* rec;
    uint32_t* count;   // phase-0 walk starts leaving v (scanned by the caller)
    int32_t* par;      // union-find parent / component minimum of the skeleton-id stage (may be null)
    int32_t* cmin;
    uint32_t* is_min;  // "is the minimum node of a component": isolated pixels are flagged here (csr.py:909)
    uint32_t split_shift = 27u;  // splitter sample = 2^-(32 - split_shift) of the chain voxels (skb_set_option "splitter_bits")
    SKB_HD void write(int64_t v, int32_t o, int32_t d, int32_t n0, int32_t n1, double w0, double w1) const {
        int32_t iv = (int32_t)v;
        skb_i4 r;
        r.x = -1;
        r.y = -1;
        r.z = o;
        r.w = d > SKB_REC_DEG_MASK ? SKB_REC_DEG_MASK : d;  // only "== 2" and row scans (bounded below) use it
        bool split = false;
        if (d == 2) {
            r.x = n0;
            r.y = n1;
            split = (iv >= sl.ga0 && iv < sl.ga1) || (iv >= sl.gb0 && iv < sl.gb1) ||
                    skb_hash_splitter(lin ? lin[v] + sl.lin_offset : (int64_t)v, split_shift);
            if (split) r.w |= SKB_REC_SPLITTER;
        }
        skb_st_i4(rec + 2 * v, r);
        skb_st_d2(rec + 2 * v + 1, w0, w1);
        uint32_t c = 0;
        if (iv >= sl.cnt0 && iv < sl.cnt1) c = (uint32_t)(d != 2 ? d : (split ? 2 : 0));
        count[v] = c;
        if (par) {
            par[v] = iv;
            cmin[v] = iv;
        }
        is_min[v] = d == 0 ? 1u : 0u;
    }
};

struct SkbPathRecItem {  // over nodes: records from an existing CSR (PathGraph.from_graph, csr.py:477-499)
    const int32_t* indptr;
    const int32_t* indices;
    const double* gdata;
    SkbRecOut out;
    SKB_HD void operator()(int64_t v) const {
        int32_t o = indptr[v], d = indptr[v + 1] - o;
        if (d == 2) out.write(v, o, d, indices[o], indices[o + 1], gdata[o], gdata[o + 1]);
        else out.write(v, o, d, -1, -1, 0.0, 0.0);
    }
};

// ---- stage: final CSR (csr.py:153-157 after 1018-1019): count -> scan -> emit ---------------
struct SkbPopcItem {
    const uint32_t* keep;
    uint32_t* deg;
    SKB_HD void operator()(int64_t i) const { deg[i] = (uint32_t)skb_popc(keep[i]); }
};

struct SkbCsrEmitFn {
    SkbWeight wf;
    int32_t a;
    int32_t o;
    int32_t* indices;
    double* data;
    int32_t first;      // position of the row's first entry
    int32_t n0, n1;     // the first two neighbours and their weights (the record of a degree-2 node)
    double w0, w1;
    int32_t cap;        // capacity of indices / data (they may be sized by a hint): nothing is written beyond it
    SKB_HD void operator()(int k, int32_t u) {
        double w = wf(a, u, k);
        if (o < cap) {
            indices[o] = u;
            data[o] = w;
        }
        if (o == first) {
            n0 = u;
            w0 = w;
        } else if (o == first + 1) {
            n1 = u;
            w1 = w;
        }
        ++o;
    }
};

struct SkbCsrEmitItem {
    SkbGeom g;
    const uint64_t* bits;
    const uint32_t* pre256;
    const int64_t* lin;
    const uint32_t* keep;
    const int32_t* indptr;
    SkbWeight wf;
    int32_t* indices;
    double* data;
    int32_t with_records;  // also write the node record the path stage walks on
    SkbRecOut out;
    int32_t cap = 0x7fffffff;
    SKB_HD void operator()(int64_t i) const {
        uint32_t m = keep[i];
        int32_t o = indptr[i];
        SkbCsrEmitFn fn{wf, (int32_t)i, o, indices, data, o, -1, -1, 0.0, 0.0, cap};
        if (m) skb_for_neighbors(g, bits, pre256, lin[i], (int32_t)i, m, fn);
        if (with_records) out.write(i, o, fn.o - o, fn.n0, fn.n1, fn.w0, fn.w1);
    }
};

// walk starts of phase 1 (junction-free cycles): both half-edges of a cycle root and of an
// uncovered splitter
struct SkbStartCountItem {  // over nodes
    const skb_i4* rec;
    const uint8_t* covered;
    SkbSlab sl;
    uint32_t* count;
    SKB_HD void operator()(int64_t v) const {
        uint32_t c = 0;
        if ((int32_t)v >= sl.own0 && (int32_t)v < sl.own1 && !covered[v]) {  // cycle roots are uncovered nodes too
            skb_i4 r = skb_ld_i4(rec + 2 * (int64_t)v);
            if (r.w & (SKB_REC_CYCLE_ROOT | SKB_REC_SPLITTER)) c = 2;
        }
        count[v] = c;
    }
};

// start s (local index) = startoff[v] - startoff[own0] + j  <->  half-edge (v -> indices[indptr[v]+j])
struct SkbStartFillItem {  // over owned nodes (item index is relative to own0)
    const int32_t* indptr;
    const uint32_t* startoff;
    int32_t own0;
    int32_t* su;
    int32_t* se;
    uint32_t cap = 0xffffffffu;  // capacity of su / se (they may be sized by a hint)
    SKB_HD void operator()(int64_t i) const {
        int64_t v = i + own0;
        uint32_t base = startoff[own0];
        uint32_t s0 = startoff[v] - base, c = startoff[v + 1] - startoff[v];
        int32_t o = indptr[v];
        for (uint32_t j = 0; j < c && s0 + j < cap; ++j) {
            su[s0 + j] = (int32_t)v;
            se[s0 + j] = o + (int32_t)j;
        }
    }
};

// ranked[s] = (x, y, z, -):  x >= 0: (global) id of the start that continues this list; x < 0:
// finished, -1-x = id of the start that is the REVERSED last half-edge (B -> m) of the list;
// y = half-edges from this start's node to (x >= 0: the next start's node | x < 0: the list end);
// z = minimum (global) node id seen so far on the list from this start on.
// Stepwise item (skb_foreach_steps): begin() reads the start, step() is one hop (one 32-byte record), finish()
// stores the result.  AUX = false on a single GPU (no per-list sums: the second walk provides them).
template <bool AUX>
struct SkbWalkItemT {  // over starts: local walk to the next terminal / cycle root / splitter
    const int32_t* indices;
    const double* gdata;   // only read when AUX
    const double* values;  // null -> 1.0
    const int64_t* lin;    // only read when AUX; may be null
    const skb_i4* rec;
    const uint32_t* startoff;
    const int32_t* su;
    const int32_t* se;
    SkbSlab sl;
    skb_i4* ranked;
    SkbAux* aux;     // null on a single GPU
    int32_t* stamp;  // null, or: stamp[v] = a start whose list passes through chain node v
    struct State {
        int32_t prev, cur, cnt, mn, out_x, start;
        double w_in, sw, sx, sxx;  // AUX only
        int64_t end_lin;
        int32_t end_node, end_deg, end_key;
        double end_value;
    };
    SKB_HD bool begin(int64_t s, State& st) const {
        int32_t u = su[s];
        int32_t arc = se[s];
        st.start = (int32_t)s;
        st.prev = u;
        st.cur = indices[arc];
        st.cnt = 1;
        st.mn = u;
        st.out_x = 0;
        if (AUX) {
            st.w_in = gdata[arc];  // weight of the edge arriving at `cur`
            double x = values ? values[u] : 1.0;
            st.sw = 0.0;
            st.sx = x;
            st.sxx = x * x;
            st.end_lin = 0;
            st.end_node = -1;
            st.end_deg = 0;
            st.end_key = -1;
            st.end_value = 0.0;
        }
        return true;
    }
    SKB_HD bool step(State& st) const {
        const int32_t cur = st.cur, prev = st.prev;
        const skb_i4* rp = rec + 2 * (int64_t)cur;
        skb_i4 r = skb_ld_i4(rp);
        if (cur < st.mn) st.mn = cur;
        if (AUX) st.sw += st.w_in;
        if (r.w != 2) {
            if ((r.w & SKB_REC_SPLITTER) && !(r.w & SKB_REC_CYCLE_ROOT)) {
                // the list continues with the splitter's other half-edge
                st.out_x = (int32_t)(startoff[cur] + (r.x == prev ? 1u : 0u)) + sl.start_shift;
                return false;
            }
            // terminal or cycle root: find the reverse half-edge (cur -> prev)
            int32_t d = r.w & SKB_REC_DEG_MASK;
            int32_t j = 0;
            if (d == SKB_REC_DEG_MASK) {  // saturated degree field: unbounded row, rely on symmetry
                while (indices[r.z + j] != prev) ++j;
            } else {
                while (j + 1 < d && indices[r.z + j] != prev) ++j;
            }
            int32_t key = (int32_t)startoff[cur] + sl.start_shift;
            st.out_x = -1 - (key + j);
            if (AUX) {
                double x = values ? values[cur] : 1.0;
                st.sx += x;
                st.sxx += x * x;
                st.end_lin = lin ? lin[cur] + sl.lin_offset : (int64_t)cur;
                st.end_node = cur + sl.id_shift;
                st.end_deg = d;
                st.end_key = key;
                st.end_value = x;
            }
            return false;
        }
        if (stamp) stamp[cur] = st.start;
        if (AUX) {
            double x = values ? values[cur] : 1.0;
            st.sx += x;
            st.sxx += x * x;
            skb_d2 wn = skb_ld_d2(rp + 1);
            st.w_in = (r.x == prev) ? wn.y : wn.x;
        }
        st.cur = (r.x == prev) ? r.y : r.x;
        st.prev = cur;
        ++st.cnt;
        return true;
    }
    SKB_HD void finish(int64_t s, const State& st) const {
        skb_i4 out;
        out.x = st.out_x;
        out.y = st.cnt;
        out.z = st.mn + sl.id_shift;
        out.w = st.cnt;  // local: half-edges to the next stop; the ranking leaves it alone (SkbEmitItem, both_ways)
        skb_st_i4(ranked + s, out);
        if (AUX) {
            SkbAux ax;
            ax.sw = st.sw;
            ax.sx = st.sx;
            ax.sxx = st.sxx;
            ax.end_lin = st.end_lin;
            ax.end_node = st.end_node;
            ax.end_deg = st.end_deg;
            ax.end_key = st.end_key;
            ax.pad = 0;
            ax.end_value = st.end_value;
            ax.pad2 = 0.0;
            aux[s] = ax;
        }
    }
    SKB_HD void operator()(int64_t s) const {
        State st;
        begin(s, st);
        while (step(st)) {
        }
        finish(s, st);
    }
};

SKB_HD void skb_aux_append(SkbAux& a, const SkbAux& b) {  // list of a followed by list of b
    a.sw += b.sw;
    a.sx += b.sx;
    a.sxx += b.sxx;
    a.end_lin = b.end_lin;
    a.end_node = b.end_node;
    a.end_deg = b.end_deg;
    a.end_key = b.end_key;
    a.end_value = b.end_value;
}

// one synchronous pointer-jumping round, src -> dst, over the starts of this rank; pointers
// outside [lo, hi) belong to other ranks and are left for the inter-slab round
struct SkbJumpItem {
    const skb_i4* src;
    skb_i4* dst;
    const SkbAux* aux_src;  // may be null
    SkbAux* aux_dst;
    int32_t lo, hi;
    int32_t* flag;  // raised when some start stops pointing into [lo, hi) in this round
    SKB_HD void operator()(int64_t s) const {
        skb_i4 a = skb_ld_i4(src + s);
        bool local = a.x >= lo && a.x < hi;
        if (local) {
            int32_t i = a.x - lo;
            skb_i4 b = skb_ld_i4(src + i);
            a.y += b.y;
            if (b.z < a.z) a.z = b.z;
            a.x = b.x;
            if (!(a.x >= lo && a.x < hi)) skb_raise(flag);
            if (aux_src) {
                SkbAux A = aux_src[s];
                skb_aux_append(A, aux_src[i]);
                aux_dst[s] = A;
            }
        } else if (aux_src) {
            aux_dst[s] = aux_src[s];
        }
        skb_st_i4(dst + s, a);
    }
};

// The lists are short (a few starts each: one per terminal half-edge plus one per sampled chain voxel), so
// before any synchronous round every start simply follows its list, up to `cap` hops, reading the
// unmodified records of src: most lists are finished by this one pass.  A start whose list closes on
// itself lies on a junction-free cycle and stays as it is (phase 1 handles it); a start that is still
// inside [lo, hi) after cap hops raises *unfinished and the synchronous rounds take over from dst.
struct SkbRankWalkItem {
    const skb_i4* src;
    skb_i4* dst;
    const SkbAux* aux_src;  // may be null
    SkbAux* aux_dst;
    int32_t lo, hi, cap;
    int32_t* unfinished;
    SKB_HD void operator()(int64_t s) const {
        skb_i4 a = skb_ld_i4(src + s);
        SkbAux A;
        if (aux_src) A = aux_src[s];
        const int32_t self = (int32_t)s + lo;
        int32_t hops = 0;
        while (a.x >= lo && a.x < hi && a.x != self) {
            if (hops == cap) {
                skb_raise(unfinished);
                break;
            }
            const int32_t i = a.x - lo;
            skb_i4 b = skb_ld_i4(src + i);
            a.y += b.y;
            if (b.z < a.z) a.z = b.z;
            a.x = b.x;
            if (aux_src) skb_aux_append(A, aux_src[i]);
            ++hops;
        }
        if (a.x == self) {  // a cycle: leave the start untouched (x >= 0 marks it as never-ending)
            a = skb_ld_i4(src + s);
            if (aux_src) A = aux_src[s];
        }
        skb_st_i4(dst + s, a);
        if (aux_src) aux_dst[s] = A;
    }
};

// ---- inter-slab level: the starts leaving slab-boundary planes, gathered from every rank -----
// table T = for rank 0, 1, ...: [starts of its first owned plane | starts of its last owned plane]
#define SKB_MAX_RANKS 16
struct SkbSlabTable {
    int32_t n_ranks;
    int32_t start_off[SKB_MAX_RANKS + 1];  // global id of every rank's first start
    int32_t n_first[SKB_MAX_RANKS];        // starts leaving the first owned plane
    int32_t last0[SKB_MAX_RANKS];          // local index of the first start leaving the last owned plane
    int32_t t_off[SKB_MAX_RANKS];          // position of the rank's block in T
};

SKB_HD int32_t skb_table_index(const SkbSlabTable& t, int32_t g) {  // global start id -> index in T, or -1
    for (int k = 0; k < t.n_ranks; ++k) {
        if (g >= t.start_off[k] && g < t.start_off[k + 1]) {
            int32_t l = g - t.start_off[k];
            if (l < t.n_first[k]) return t.t_off[k] + l;
            if (l >= t.last0[k]) return t.t_off[k] + t.n_first[k] + (l - t.last0[k]);
            return -1;
        }
    }
    return -1;
}

struct SkbTableJumpItem {  // over T: pointer jumping on the inter-slab list (every rank does the same)
    const skb_i4* src;
    skb_i4* dst;
    const SkbAux* aux_src;
    SkbAux* aux_dst;
    SkbSlabTable t;
    int32_t* flag;
    SKB_HD void operator()(int64_t s) const {
        skb_i4 a = skb_ld_i4(src + s);
        SkbAux A = aux_src[s];
        if (a.x >= 0) {
            int32_t i = skb_table_index(t, a.x);
            if (i < 0) {
                a.x = SKB_DEAD;
            } else {
                skb_i4 b = skb_ld_i4(src + i);
                a.y += b.y;
                if (b.z < a.z) a.z = b.z;
                a.x = b.x;
                skb_aux_append(A, aux_src[i]);
            }
            if (a.x < 0) skb_raise(flag);
        }
        skb_st_i4(dst + s, a);
        aux_dst[s] = A;
    }
};

struct SkbTableExportItem {  // over the rank's (padded) block of T: copy the ranked boundary-plane starts
    const skb_i4* ranked;
    const SkbAux* aux;
    int32_t n_first, last0, lo, hi;  // hi - lo = starts of this rank; pointers still inside are cycles
    int32_t n_block;                 // rows beyond the block are padding: finished, never referenced
    skb_i4* out;
    SkbAux* out_aux;
    SKB_HD void operator()(int64_t i) const {
        if (i >= n_block) {
            skb_i4 pad;
            pad.x = pad.y = pad.z = pad.w = -1;
            skb_st_i4(out + i, pad);
            return;
        }
        int64_t s = i < n_first ? i : last0 + (i - n_first);
        skb_i4 a = skb_ld_i4(ranked + s);
        if (a.x >= lo && a.x < hi) a.x = SKB_DEAD;
        skb_st_i4(out + i, a);
        out_aux[i] = aux[s];
    }
};

struct SkbTableApplyItem {  // over the starts of this rank: finish the lists that left the slab
    skb_i4* ranked;
    SkbAux* aux;
    const skb_i4* table;
    const SkbAux* table_aux;
    SkbSlabTable t;
    int32_t lo, hi;
    SKB_HD void operator()(int64_t s) const {
        skb_i4 a = skb_ld_i4(ranked + s);
        if (a.x < 0) return;
        if (a.x >= lo && a.x < hi) {  // still pointing at a start of this rank: a cycle through splitters
            a.x = SKB_DEAD;
        } else {
            int32_t i = skb_table_index(t, a.x);
            skb_i4 b;
            b.x = SKB_DEAD;
            b.y = 0;
            b.z = a.z;
            if (i >= 0) b = skb_ld_i4(table + i);
            if (b.x >= 0 || b.x == SKB_DEAD) {
                a.x = SKB_DEAD;
            } else {
                a.y += b.y;
                if (b.z < a.z) a.z = b.z;
                a.x = b.x;
                SkbAux A = aux[s];
                skb_aux_append(A, table_aux[i]);
                aux[s] = A;
            }
        }
        skb_st_i4(ranked + s, a);
    }
};

// a half-edge (A,n) starts an emitted path iff it precedes the reversed last half-edge (B,m)
// in CSR order == in start order (the serial walk of csr.py:347-409 visits start
// nodes and their neighbours ascending and marks both directions visited)
SKB_HD bool skb_start_is_head(const skb_i4& r_u, const skb_i4& a, int64_t s_global, int phase) {
    bool end_type = phase == 0 ? ((r_u.w & SKB_REC_DEG_MASK) != 2) : ((r_u.w & SKB_REC_CYCLE_ROOT) != 0);
    return end_type && a.x < 0 && a.x != SKB_DEAD && s_global < (int64_t)(-1 - a.x);
}

struct SkbSelectItem {  // over starts
    const int32_t* su;
    const skb_i4* rec;
    const skb_i4* ranked;
    int32_t phase;
    int32_t start_lo;
    uint32_t* sel;
    SKB_HD void operator()(int64_t s) const {
        sel[s] = skb_start_is_head(skb_ld_i4(rec + 2 * (int64_t)su[s]), skb_ld_i4(ranked + s), s + start_lo, phase) ? 1u : 0u;
    }
};

struct SkbHeadsItem {  // over starts: path id, start node, length, minimum node; tag the list end
    const int32_t* su;
    const skb_i4* rec;
    const skb_i4* ranked;
    const SkbAux* aux;    // may be null
    const uint32_t* sel;  // exclusive scan of SkbSelectItem flags
    int32_t phase;
    int32_t pid_base;
    int32_t start_lo;
    int32_t id_shift;
    int32_t* start_node;
    uint32_t* plen;
    int32_t* pmin;
    skb_i4* einfo;      // single GPU: pre-set to -1; einfo[reversed last start] = (path id, half-edges, -, -)
    int32_t* head_end;  // slabs: id of the reversed last start of every path
    SkbAux* head_aux;   // slabs: sums and end node of every path
    int32_t pid_cap = 0x7fffffff;  // capacity of the per-path arrays (they may be sized by a hint)
    int32_t both_ways = 0;         // single GPU: also tag the head itself, for the starts that look back at it
    SKB_HD void operator()(int64_t s) const {
        skb_i4 a = skb_ld_i4(ranked + s);
        if (!skb_start_is_head(skb_ld_i4(rec + 2 * (int64_t)su[s]), a, s + start_lo, phase)) return;
        int32_t pid = pid_base + (int32_t)sel[s];
        if (pid >= pid_cap) return;
        start_node[pid] = su[s] + id_shift;
        plen[pid] = (uint32_t)a.y + 1u;
        pmin[pid] = a.z;
        if (einfo) {
            skb_i4 info;
            info.x = pid;
            info.y = a.y;
            info.z = 0;
            info.w = 0;
            skb_st_i4(einfo + (-1 - a.x), info);
            if (both_ways) {
                // the lists of the starts that point back along this path end at the head: its slot is free
                // (the head is the reversed last start of the reversed path only, which is never emitted)
                info.w = 1;
                skb_st_i4(einfo + s, info);
            }
        }
        if (head_end) {
            head_end[pid] = -1 - a.x;
            head_aux[pid] = aux[s];
        }
    }
};

struct SkbEmitItem {  // over starts: second local walk, writes path entries (csr.py:394-409); stepwise item
    const int32_t* indices;
    const double* gdata;
    const skb_i4* rec;
    const int32_t* su;
    const int32_t* se;
    const skb_i4* ranked;
    const skb_i4* einfo;     // indexed by the id of the reversed last start: (path id, half-edges, base, looks back)
    const int32_t* pindptr;  // null: the entry offset of the path is einfo.z
    const double* values;    // null -> 1.0 (csr.py:207-215)
    uint8_t* covered;        // may be null; every node written is marked (what stays 0 is on no path headed at a terminal)
    int32_t id_shift;
    int32_t* pidx;
    double* pdata;
    double* pw;  // weight of the edge arriving at each entry (0 for the first)
    // both_ways (single GPU; needs SkbHeadsItem.both_ways and the local hop count in ranked.w): the chain between two
    // stops is written from both of its ends.  The start that looks along the path writes the first half of the
    // interior voxels, the start at the far stop that looks back writes its own voxel and the second half.  The kernel
    // lasts as long as its longest walk (dependent loads): this halves it.  0: the start that looks along the
    // path writes the whole chain and the stop.
    int32_t both_ways = 0;
    struct State {
        int64_t at;    // entry index of the node last written
        int32_t prev, cur;
        int32_t left;  // both_ways: voxels still to write; else unused (the walk ends at the stop)
        int32_t back;  // 1: walking against the path, entry indices decrease
        double w_in;   // forward: weight of the edge arriving at `cur` (the first from the graph, later ones from the records)
    };
    SKB_HD void put(int64_t at, int32_t node, double w) const {
        pidx[at] = node + id_shift;
        if (pdata) pdata[at] = values ? values[node] : 1.0;
        pw[at] = w;
        if (covered) covered[node] = 1;
    }
    SKB_HD bool begin(int64_t s, State& st) const {
        skb_i4 a = skb_ld_i4(ranked + s);
        if (a.x >= 0 || a.x == SKB_DEAD) return false;  // never reaches a terminal: a cycle that phase 1 handles
        skb_i4 info = skb_ld_i4(einfo + (-1 - a.x));
        if (info.x < 0) return false;  // on no emitted path in this direction
        const bool back = both_ways && info.w == 1;
        if (!both_ways && info.w != 0) return false;
        int64_t base = pindptr ? pindptr[info.x] : info.z;
        int32_t u = su[s];
        int32_t arc = se[s];
        st.prev = u;
        st.cur = indices[arc];
        st.back = back ? 1 : 0;
        const int32_t interior = a.w - 1;           // chain voxels strictly between this start's node and the next stop
        const int32_t ahead = (interior + 1) / 2;   // the share of the start that looks along the path
        if (back) {
            // a.y = half-edges from this node back to the path's first node = its position on the path
            st.at = base + a.y;
            put(st.at, u, gdata[arc]);  // the edge arriving here along the path is this start's own half-edge, reversed
            st.left = interior - ahead;
            st.w_in = 0.0;
            return st.left > 0;
        }
        int32_t pos = info.y - a.y;  // position of this start's node on the path
        if (pos == 0) put(base, u, 0.0);
        st.at = base + pos;
        st.w_in = gdata[arc];
        st.left = both_ways ? ahead : 0x7fffffff;
        return st.left > 0;
    }
    SKB_HD bool step(State& st) const {
        const int32_t cur = st.cur;
        const skb_i4* rp = rec + 2 * (int64_t)cur;
        skb_i4 r = skb_ld_i4(rp);
        if (st.back) {  // cur is an interior chain voxel: its entry carries the weight of the edge to the NEXT voxel back
            const int64_t at = --st.at;
            skb_d2 w = skb_ld_d2(rp + 1);
            const bool first = r.x == st.prev;
            put(at, cur, first ? w.y : w.x);
            st.cur = first ? r.y : r.x;
            st.prev = cur;
            return --st.left > 0;
        }
        const int64_t at = ++st.at;
        put(at, cur, st.w_in);
        if (r.w != 2) return false;     // the stop itself (only reached when the whole chain is this start's)
        if (--st.left == 0) return false;
        skb_d2 w = skb_ld_d2(rp + 1);
        if (r.x == st.prev) {
            st.cur = r.y;
            st.w_in = w.y;
        } else {
            st.cur = r.x;
            st.w_in = w.x;
        }
        st.prev = cur;
        return true;
    }
    SKB_HD void finish(int64_t, const State&) const {}
    SKB_HD void operator()(int64_t s) const {
        State st;
        if (!begin(s, st)) return;
        while (step(st)) {
        }
    }
};

// ---- junction-free cycles: minimum node of every uncovered degree-2 component ---------------
SKB_HD void skb_union_min(int32_t* par, int32_t a, int32_t b) {
    for (;;) {
        a = skb_find(par, a);
        b = skb_find(par, b);
        if (a == b) return;
        if (a < b) {
            int32_t t = a;
            a = b;
            b = t;
        }
        // larger root goes under the smaller one: the root stays the component minimum
        if (skb_atomic_cas_i32(par + a, a, b) == a) return;
    }
}

// Union by a pseudo-random priority instead of by id: linking roots in id order builds chains as deep as the
// component is large (find then costs one dependent load per level); random linking keeps the trees shallow.
// For callers that do not need the root to be the component minimum.
SKB_HD uint32_t skb_mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}
SKB_HD void skb_union_rand(int32_t* par, int32_t a, int32_t b) {
    for (;;) {
        a = skb_find(par, a);
        b = skb_find(par, b);
        if (a == b) return;
        const uint32_t ha = skb_mix32((uint32_t)a), hb = skb_mix32((uint32_t)b);
        if (ha < hb || (ha == hb && a < b)) {
            int32_t t = a;
            a = b;
            b = t;
        }
        if (skb_atomic_cas_i32(par + a, a, b) == a) return;
    }
}

struct SkbTilePrefixItem {  // tile_prefix[t] = nodes before tile t, read back from the rank structure
    const uint32_t* pre256;
    const int32_t* n_nodes;
    int64_t ntiles;
    uint32_t* tile_prefix;
    SKB_HD void operator()(int64_t t) const { tile_prefix[t] = t < ntiles ? pre256[t * 64] : (uint32_t)*n_nodes; }
};

struct SkbTileCountItem {  // over tiles tile0 .. : count[t] = set bits of the tile's 256 mask words (halo planes that arrived)
    const uint64_t* bits;
    uint32_t* count;
    int64_t tile0;
    SKB_HD void operator()(int64_t i) const {
        const int64_t t = tile0 + i;
        const uint64_t* w = bits + t * 256;
        uint32_t c = 0;
        for (int k = 0; k < 256; ++k) c += (uint32_t)skb_popcll(w[k]);
        count[t] = c;
    }
};

struct SkbIotaItem {
    int32_t* par;
    SKB_HD void operator()(int64_t v) const { par[v] = (int32_t)v; }
};

struct SkbCycleHookItem {  // over nodes: uncovered degree-2 nodes join their smaller neighbours
    const skb_i4* rec;
    const uint8_t* covered;
    int32_t* par;
    SKB_HD void operator()(int64_t v) const {
        if (covered[v]) return;  // one byte decides for all but the few uncovered nodes: the 32-byte record stays unread
        skb_i4 r = skb_ld_i4(rec + 2 * (int64_t)v);
        if ((r.w & SKB_REC_DEG_MASK) != 2) return;
        if (r.x < (int32_t)v) skb_union_min(par, (int32_t)v, r.x);
        if (r.y < (int32_t)v) skb_union_min(par, (int32_t)v, r.y);
    }
};

struct SkbCycleRootItem {  // over nodes: the minimum of every cycle becomes its terminal
    skb_i4* rec;
    const uint8_t* covered;
    int32_t* par;
    uint32_t* is_min;  // a junction-free cycle is a component of its own (csr.py:909)
    SKB_HD void operator()(int64_t v) const {
        if (covered[v]) return;
        skb_i4 r = skb_ld_i4(rec + 2 * (int64_t)v);
        if ((r.w & SKB_REC_DEG_MASK) != 2) return;
        if (skb_find(par, (int32_t)v) == (int32_t)v) {
            rec[2 * (int64_t)v].w |= SKB_REC_CYCLE_ROOT;
            is_min[v] = 1u;
        }
    }
};

// ---- stage: skeleton ids (csr.py:909 connected_components) on the contracted graph ----------
// Components of the pixel graph == components of the graph whose nodes are terminals and whose
// edges are paths; the reference labels components by the rank of their minimum node id, which
// may be an interior path node, hence the per-path minimum carried through the ranking above.
struct SkbCcPathUnionItem {  // over regular paths
    const int32_t* pindptr;
    const int32_t* pidx;
    int32_t* par;
    SKB_HD void operator()(int64_t p) const {
        skb_union_rand(par, pidx[pindptr[p]], pidx[pindptr[p + 1] - 1]);  // cmin, not the root, carries the minimum
    }
};

struct SkbCcPathMinItem {  // over regular paths, after all unions
    const int32_t* pindptr;
    const int32_t* pidx;
    const int32_t* pmin;
    int32_t* par;
    int32_t* cmin;
    SKB_HD void operator()(int64_t p) const {
        int32_t r = skb_find(par, pidx[pindptr[p]]);
        // thousands of paths of one large component would queue on one address: the value only decreases, so a
        // plain look first lets all but the few that still lower it skip the atomic
        const uint32_t m = (uint32_t)pmin[p];
        if ((uint32_t)skb_ld_cg(cmin + r) > m) skb_atomic_min_u32((uint32_t*)cmin + r, m);
    }
};

struct SkbCcPathMarkItem {  // over regular paths: flag the minimum node of the path's component
    const int32_t* pindptr;
    const int32_t* pidx;
    int32_t* par;
    const int32_t* cmin;
    uint32_t* is_min;
    SKB_HD void operator()(int64_t p) const { is_min[cmin[skb_find(par, pidx[pindptr[p]])]] = 1u; }
};

struct SkbCcPathIdItem {  // over all paths
    const int32_t* pindptr;
    const int32_t* pidx;
    int32_t n_regular;
    int32_t* par;
    const int32_t* cmin;
    const uint32_t* rank;  // exclusive scan of is_min
    int32_t* skel_id;
    const int32_t* n_regular_ptr = nullptr;  // the number of regular paths in device memory (overrides n_regular)
    SKB_HD void operator()(int64_t p) const {
        int32_t src = pidx[pindptr[p]];
        const int32_t nreg = n_regular_ptr ? *n_regular_ptr : n_regular;
        int32_t m = p < nreg ? cmin[skb_find(par, src)] : src;
        skel_id[p] = (int32_t)rank[m];
    }
};

// ---- Z-slab mode: pieces that only exist when a volume is split over several GPUs -------------
// copy up to SKB_MAX_SEGMENTS byte ranges (16-byte granules): regroups what an all-gather
// delivered rank by rank into one array per field
#define SKB_MAX_SEGMENTS 64
struct SkbSegCopyItem {  // over 16-byte granules of all segments
    int32_t n_seg;
    const uint8_t* src[SKB_MAX_SEGMENTS];
    uint8_t* dst[SKB_MAX_SEGMENTS];
    int64_t first[SKB_MAX_SEGMENTS + 1];  // first granule of each segment (exclusive scan of sizes)
    SKB_HD void operator()(int64_t i) const {
        int k = 0;
        while (k + 1 < n_seg && i >= first[k + 1]) ++k;
        int64_t o = (i - first[k]) * 16;
        skb_st_i4((skb_i4*)(dst[k] + o), skb_ld_i4((const skb_i4*)(src[k] + o)));
    }
};

struct SkbEdgeShiftItem {  // gathered junction edges: local ids of rank k -> global ids
    int32_t* ea;
    int32_t* eb;
    int32_t cap;  // rows per rank
    int32_t shift[SKB_MAX_RANKS];
    SKB_HD void operator()(int64_t e) const {
        if (ea[e] < 0) return;
        int32_t sh = shift[e / cap];
        ea[e] += sh;
        eb[e] += sh;
    }
};

struct SkbRankQueryItem {  // out[i] = number of foreground voxels before raveled position pos[i]
    const uint64_t* bits;
    const uint32_t* pre256;
    const int64_t* pos;
    int32_t* out;
    SKB_HD void operator()(int64_t i) const { out[i] = skb_rank(bits, pre256, pos[i]); }
};

struct SkbCoveredItem {  // over nodes: a chain node is covered when a list through it reached a terminal
    const skb_i4* rec;
    const uint32_t* startoff;
    const int32_t* stamp;  // pre-set to -1
    const skb_i4* ranked;
    int32_t own0, own1;
    uint8_t* covered;
    uint32_t* n_uncovered;  // 64 slots, pre-zeroed: owned chain nodes not covered
    SKB_HD void operator()(int64_t v) const {
        skb_i4 r = skb_ld_i4(rec + 2 * (int64_t)v);
        bool owned = (int32_t)v >= own0 && (int32_t)v < own1;
        uint8_t c = owned ? 0 : 1;  // halo nodes are another rank's business
        if ((r.w & SKB_REC_DEG_MASK) == 2 && owned) {
            int32_t s = (r.w & SKB_REC_SPLITTER) ? (int32_t)(startoff[v] - startoff[own0]) : stamp[v];
            if (s >= 0) {
                skb_i4 a = skb_ld_i4(ranked + s);
                c = (a.x < 0 && a.x != SKB_DEAD) ? 1 : 0;
            }
            if (!c) {
#if defined(__CUDA_ARCH__)
                atomicAdd(n_uncovered + (blockIdx.x & 63), 1u);
#else
                *n_uncovered += 1u;
#endif
            }
        }
        covered[v] = c;
    }
};

// per path headed on this rank: what every rank needs to place the path's entries
// (head = (reversed last start, global path id, half-edges, global entry offset)) and to label
// components (link = (key of the first node, key of the last node, minimum node id, -))
struct SkbHeadRecordItem {
    const int32_t* head_end;
    const uint32_t* plen;
    const int32_t* pindptr;  // local exclusive scan of plen
    const int32_t* pmin;
    const int32_t* start_node;  // global node ids
    const uint32_t* startoff;
    const SkbAux* head_aux;
    SkbSlab sl;
    int32_t pid_off, entry_off;
    int32_t n_paths;  // rows beyond are padding (x = -1)
    skb_i4* head;
    skb_i4* link;
    SKB_HD void operator()(int64_t p) const {
        skb_i4 h;
        if (p >= n_paths) {
            h.x = h.y = h.z = h.w = -1;
            skb_st_i4(head + p, h);
            skb_st_i4(link + p, h);
            return;
        }
        h.x = head_end[p];
        h.y = pid_off + (int32_t)p;
        h.z = (int32_t)plen[p] - 1;
        h.w = entry_off + pindptr[p];
        skb_st_i4(head + p, h);
        skb_i4 l;
        l.x = (int32_t)startoff[start_node[p] - sl.id_shift] + sl.start_shift;
        l.y = head_aux[p].end_key;
        l.z = pmin[p];
        l.w = 0;
        skb_st_i4(link + p, l);
    }
};

struct SkbEinfoScatterItem {  // over the gathered head records: einfo[reversed last start] = (id, half-edges, offset)
    const skb_i4* head;
    skb_i4* einfo;  // one slot per start of the whole volume, pre-set to -1
    SKB_HD void operator()(int64_t p) const {
        skb_i4 h = skb_ld_i4(head + p);
        if (h.x < 0) return;  // padding
        skb_i4 info;
        info.x = h.y;
        info.y = h.z;
        info.z = h.w;
        info.w = 0;
        skb_st_i4(einfo + h.x, info);
    }
};

// skeleton ids across slabs: union-find over start keys (a dense key space for terminals)
struct SkbKeyInitItem {
    int32_t* par;
    int32_t* cmin;
    SKB_HD void operator()(int64_t k) const {
        par[k] = (int32_t)k;
        cmin[k] = 0x7fffffff;
    }
};
struct SkbKeyUnionItem {  // over the gathered links
    const skb_i4* link;
    int32_t* par;
    SKB_HD void operator()(int64_t p) const {
        skb_i4 l = skb_ld_i4(link + p);
        if (l.x < 0) return;  // padding
        skb_union_rand(par, l.x, l.y);  // cmin, not the root, carries the minimum
    }
};
struct SkbKeyMinItem {
    const skb_i4* link;
    int32_t* par;
    int32_t* cmin;
    SKB_HD void operator()(int64_t p) const {
        skb_i4 l = skb_ld_i4(link + p);
        if (l.x < 0) return;
        const int32_t r = skb_find(par, l.x);
        if ((uint32_t)skb_ld_cg(cmin + r) > (uint32_t)l.z) skb_atomic_min_u32((uint32_t*)cmin + r, (uint32_t)l.z);
    }
};
struct SkbKeyMarkItem {  // component minima that are nodes of this rank
    const skb_i4* link;
    int32_t* par;
    const int32_t* cmin;
    int32_t node_lo, node_hi, id_shift;  // owned global node ids [node_lo, node_hi)
    uint32_t* is_min;                    // indexed by local node id
    SKB_HD void operator()(int64_t p) const {
        skb_i4 l = skb_ld_i4(link + p);
        if (l.x < 0) return;
        int32_t m = cmin[skb_find(par, l.x)];
        if (m >= node_lo && m < node_hi) is_min[m - id_shift] = 1u;
    }
};
struct SkbKeySkelItem {  // contribution of this rank to the skeleton id of every gathered path
    const skb_i4* link;
    int32_t* par;
    const int32_t* cmin;
    const uint32_t* rank;  // local exclusive scan of is_min
    int32_t node_lo, node_hi, id_shift, comp_off, own0;
    int32_t* skel;  // summed over ranks afterwards
    SKB_HD void operator()(int64_t p) const {
        skb_i4 l = skb_ld_i4(link + p);
        int32_t out = 0;
        if (l.x >= 0) {
            int32_t m = cmin[skb_find(par, l.x)];
            if (m >= node_lo && m < node_hi) out = comp_off + (int32_t)(rank[m - id_shift] - rank[own0]);
        }
        skel[p] = out;
    }
};

struct SkbAnyUnfinishedItem {  // over the ranked inter-slab table: a list that never ends crosses slabs in a cycle
    const skb_i4* table;
    int32_t* flag;
    SKB_HD void operator()(int64_t i) const {
        if (skb_ld_i4(table + i).x >= 0) skb_raise(flag);
    }
};

struct SkbAddOffsetItem {  // p[i] = q[i] + delta
    const int32_t* q;
    int32_t* p;
    int32_t delta;
    SKB_HD void operator()(int64_t i) const { p[i] = q[i] + delta; }
};

struct SkbSlabSkelFinalItem {  // skeleton ids of the paths headed on this rank, from the reduced table / the local ranks
    const int32_t* skel_all;  // summed over ranks; this rank's regular paths start at row0
    int64_t row0;
    int32_t n_regular;
    const int32_t* start_node;  // global ids
    int32_t id_shift, comp_off, own0;
    const uint32_t* rank;  // local exclusive scan of is_min
    int32_t* skel;
    SKB_HD void operator()(int64_t p) const {
        if (p < n_regular) skel[p] = skel_all[row0 + p];
        else skel[p] = comp_off + (int32_t)(rank[start_node[p] - id_shift] - rank[own0]);
    }
};

struct SkbCopyCycleHeadsItem {  // over the rings a rank heads: their heads move behind the regular paths' (ids pid_base + j)
    const int32_t* t_start;
    const uint32_t* t_plen;
    const int32_t* t_pmin;
    const SkbAux* t_aux;
    int32_t pid_base;
    int32_t* start_node;
    uint32_t* plen;
    int32_t* pmin;
    SkbAux* head_aux;
    SKB_HD void operator()(int64_t j) const {
        const int64_t p = pid_base + j;
        start_node[p] = t_start[j];
        plen[p] = t_plen[j];
        pmin[p] = t_pmin[j];
        head_aux[p] = t_aux[j];
    }
};

struct SkbFillF64Item {  // out[i] = v (paths.data of a boolean image is all ones: one coalesced fill)
    double* out;
    double v;
    SKB_HD void operator()(int64_t i) const { out[i] = v; }
};

struct SkbSetI64Item {  // out[i] = v[i]: a few host integers as a kernel argument
    int64_t v[8];
    int64_t* out;
    SKB_HD void operator()(int64_t i) const { out[i] = v[i]; }
};

// ---- Z-slab mode: junction-free cycles that touch a slab boundary ------------------------------------
// Rare, and not worth a second two-level ranking: every rank exports its uncovered owned chain voxels
// (48 bytes each), all ranks gather them in rank order (= ascending global id) and solve the same small
// problem: union-find over the records, one ring per root; the rank that owns a ring's minimum node
// emits the whole ring by walking it once (csr.py:369-386: [s, smaller neighbour of s, ..., s]).
struct alignas(16) SkbRingRec {
    int32_t gid, gn0, gn1, pad;  // global ids of the voxel and of its two neighbours
    double w0, w1, value;        // weights of the edges to gn0 / gn1, node value (1.0 for boolean images)
};

struct SkbRingFlagItem {  // over nodes: flag[v] = owned, degree 2, not covered by a terminal-started list
    const skb_i4* rec;
    const uint8_t* covered;
    int32_t own0, own1;
    uint32_t* flag;
    SKB_HD void operator()(int64_t v) const {
        uint32_t f = 0;
        if ((int32_t)v >= own0 && (int32_t)v < own1 && !covered[v])
            f = (skb_ld_i4(rec + 2 * v).w & SKB_REC_DEG_MASK) == 2 ? 1u : 0u;
        flag[v] = f;
    }
};

struct SkbRingExportItem {  // over nodes, after the scan of the flags (pos[v+1] > pos[v] <=> flagged)
    const skb_i4* rec;
    const uint32_t* pos;
    const double* values;  // may be null
    int32_t id_shift;
    int64_t cap;           // rows of the send buffer; rows beyond the flagged count get gid = -1
    SkbRingRec* out;
    SKB_HD void operator()(int64_t v) const {
        if (pos[v + 1] == pos[v] || (int64_t)pos[v] >= cap) return;
        skb_i4 r = skb_ld_i4(rec + 2 * v);
        skb_d2 w = skb_ld_d2(rec + 2 * v + 1);
        SkbRingRec o;
        o.gid = (int32_t)v + id_shift;
        o.gn0 = r.x + id_shift;
        o.gn1 = r.y + id_shift;
        o.pad = 0;
        o.w0 = w.x;
        o.w1 = w.y;
        o.value = values ? values[v] : 1.0;
        out[pos[v]] = o;
    }
};

SKB_HD int32_t skb_ring_find(const SkbRingRec* R, int32_t m, int32_t gid) {  // index of the record of gid, or -1
    int32_t lo = 0, hi = m - 1;
    while (lo <= hi) {
        int32_t mid = (lo + hi) >> 1;
        int32_t g = R[mid].gid;
        if (g == gid) return mid;
        if (g < gid) lo = mid + 1;
        else hi = mid - 1;
    }
    return -1;
}

struct SkbRingHookItem {  // over records: join the smaller neighbours (par initialised to the identity)
    const SkbRingRec* R;
    int32_t m;
    int32_t* par;
    int32_t* bad;  // raised when a neighbour is missing: the exported set is not a union of rings
    SKB_HD void operator()(int64_t i) const {
        const int32_t n[2] = {R[i].gn0, R[i].gn1};
        for (int k = 0; k < 2; ++k) {
            int32_t j = skb_ring_find(R, m, n[k]);
            if (j < 0) skb_raise(bad);
            else if (j < (int32_t)i) skb_union_min(par, (int32_t)i, j);
        }
    }
};

struct SkbRingSizeItem {  // over records: size[root] += 1; owned[i] = record i is the root of a ring headed on this rank
    const SkbRingRec* R;
    int32_t* par;
    int32_t node_lo, node_hi;
    uint32_t* size;   // zeroed
    uint32_t* owned;  // m + 1 entries, scanned afterwards
    SKB_HD void operator()(int64_t i) const {
        int32_t r = skb_find(par, (int32_t)i);
#if defined(__CUDA_ARCH__)
        atomicAdd(size + r, 1u);
#else
        size[r] += 1u;
#endif
        owned[i] = (r == (int32_t)i && R[i].gid >= node_lo && R[i].gid < node_hi) ? 1u : 0u;
    }
};

struct SkbRingHeadItem {  // over records: the rings headed here become paths pid_base + rank
    const SkbRingRec* R;
    const uint32_t* size;
    const uint32_t* oidx;  // scanned `owned`
    int32_t pid_base, id_shift;
    int32_t* start_node;
    uint32_t* plen;
    int32_t* pmin;
    uint32_t* is_min;  // a junction-free cycle is a component of its own (csr.py:909)
    SKB_HD void operator()(int64_t i) const {
        if (oidx[i + 1] == oidx[i]) return;
        const int32_t pid = pid_base + (int32_t)oidx[i];
        start_node[pid] = R[i].gid;
        plen[pid] = size[i] + 1u;
        pmin[pid] = R[i].gid;
        is_min[R[i].gid - id_shift] = 1u;
    }
};

struct SkbRingWalkItem {  // over records: the owner of a ring writes all its entries and its sums
    const SkbRingRec* R;
    int32_t m;
    const uint32_t* oidx;
    int32_t pid_base, id_shift;
    const int32_t* gptr;  // global entry offset of every path headed here
    const int64_t* lin;   // local raveled indices
    int64_t lin_offset;
    int32_t* pidx;
    double* pdata;
    double* pw;
    SkbAux* head_aux;
    SKB_HD void operator()(int64_t i) const {
        if (oidx[i + 1] == oidx[i]) return;
        const int32_t pid = pid_base + (int32_t)oidx[i];
        int64_t at = gptr[pid];
        const SkbRingRec s = R[i];
        SkbAux a;
        a.sw = 0.0;
        a.sx = s.value;
        a.sxx = s.value * s.value;
        pidx[at] = s.gid;
        pdata[at] = s.value;
        pw[at] = 0.0;
        ++at;
        // towards the smaller neighbour first
        int32_t prev = s.gid;
        int32_t cur = s.gn0 < s.gn1 ? s.gn0 : s.gn1;
        double w_in = s.gn0 < s.gn1 ? s.w0 : s.w1;
        for (int32_t steps = 0; steps <= m; ++steps) {
            const int32_t j = cur == s.gid ? (int32_t)i : skb_ring_find(R, m, cur);
            const SkbRingRec c = R[j];
            pidx[at] = c.gid;
            pdata[at] = c.value;
            pw[at] = w_in;
            ++at;
            a.sw += w_in;
            a.sx += c.value;
            a.sxx += c.value * c.value;
            if (c.gid == s.gid) break;
            const bool via0 = c.gn0 != prev;
            prev = c.gid;
            cur = via0 ? c.gn0 : c.gn1;
            w_in = via0 ? c.w0 : c.w1;
        }
        a.end_lin = lin[s.gid - id_shift] + lin_offset;
        a.end_node = s.gid;
        a.end_deg = 2;
        a.end_key = 0;
        a.pad = 0;
        a.end_value = s.value;
        a.pad2 = 0.0;
        head_aux[pid] = a;
    }
};

#define SKB_MAX_FETCH 16
struct SkbGatherI32Item {  // out[i] = *ptr[i]: the scalars of one host read-back, made contiguous
    const int32_t* ptr[SKB_MAX_FETCH];
    int32_t* out;
    SKB_HD void operator()(int64_t i) const { out[i] = *ptr[i]; }
};

struct SkbSumSlotsItem {  // one item: out = sum of n spread counters
    const uint32_t* in;
    int32_t n;
    int32_t* out;
    SKB_HD void operator()(int64_t) const {
        uint32_t t = 0;
        for (int i = 0; i < n; ++i) t += in[i];
        *out = (int32_t)t;
    }
};

struct SkbPeerSumItem {  // dst[i] = sum over ranks of src[k][i] (every rank's contribution, read in place over NVLink)
    const int32_t* src[SKB_MAX_RANKS];
    int32_t n_src;
    int32_t* dst;
    SKB_HD void operator()(int64_t i) const {
        int32_t t = 0;
        for (int k = 0; k < n_src; ++k) t += src[k][i];
        dst[i] = t;
    }
};

struct SkbSlabStatsItem {  // over the paths headed on this rank
    const uint32_t* plen;
    const SkbAux* head_aux;
    double* dist;
    double* mean;
    double* stdev;
    SKB_HD void operator()(int64_t p) const {
        double n = (double)plen[p];
        SkbAux a = head_aux[p];
        dist[p] = a.sw;
        double m = a.sx / n;
        mean[p] = m;
        double var = a.sxx / n - m * m;
        if (!(var > 0.0)) var = (var == var) ? 0.0 : var;
        stdev[p] = skb_sqrt(var);
    }
};

// ---- stage: per-path statistics (csr.py:962-977, 792-816) and branch table (csr.py:909-952) --
// np.add.reduceat (csr.py:800,813) = first element + NumPy's pairwise sum of the rest
// (numpy/_core/src/umath/loops_utils.h.src pairwise_sum: 8 accumulators, blocks of 128).
struct SkbSqLoad {
    const double* p;
    SKB_HD double operator()(int64_t j) const { return p[j] * p[j]; }
};
struct SkbIdLoad {
    const double* p;
    SKB_HD double operator()(int64_t j) const { return p[j]; }
};

template <typename Load>
SKB_HD double skb_np_block_sum(const Load& ld, int64_t s, int64_t n) {  // n <= 128
    if (n < 8) {
        double r = 0.0;
        for (int64_t i = 0; i < n; ++i) r += ld(s + i);
        return r;
    }
    double r0 = ld(s), r1 = ld(s + 1), r2 = ld(s + 2), r3 = ld(s + 3);
    double r4 = ld(s + 4), r5 = ld(s + 5), r6 = ld(s + 6), r7 = ld(s + 7);
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
        r0 += ld(s + i);
        r1 += ld(s + i + 1);
        r2 += ld(s + i + 2);
        r3 += ld(s + i + 3);
        r4 += ld(s + i + 4);
        r5 += ld(s + i + 5);
        r6 += ld(s + i + 6);
        r7 += ld(s + i + 7);
    }
    double res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; ++i) res += ld(s + i);
    return res;
}

// iterative form of the recursion  f(s,n) = n<=128 ? block : f(s,n2) + f(s+n2,n-n2),
// n2 = (n/2) rounded down to a multiple of 8; explicit stack (depth <= 40)
template <typename Load>
SKB_HD double skb_np_pairwise(const Load& ld, int64_t s0, int64_t n0) {
    int64_t st_s[48], st_n[48];
    double st_v[48];
    int8_t st_state[48];
    int top = 0;
    st_s[0] = s0;
    st_n[0] = n0;
    st_state[0] = 0;
    double ret = 0.0;
    while (top >= 0) {
        int64_t s = st_s[top], n = st_n[top];
        if (st_state[top] == 0) {
            if (n <= 128) {
                ret = skb_np_block_sum(ld, s, n);
                --top;
            } else {
                int64_t n2 = n / 2;
                n2 -= n2 % 8;
                st_state[top] = 1;
                ++top;
                st_s[top] = s;
                st_n[top] = n2;
                st_state[top] = 0;
            }
        } else if (st_state[top] == 1) {
            int64_t n2 = n / 2;
            n2 -= n2 % 8;
            st_v[top] = ret;
            st_state[top] = 2;
            ++top;
            st_s[top] = s + n2;
            st_n[top] = n - n2;
            st_state[top] = 0;
        } else {
            ret = st_v[top] + ret;
            --top;
        }
    }
    return ret;
}

template <typename Load>
SKB_HD double skb_np_reduceat(const Load& ld, int64_t s, int64_t n) {
    double first = ld(s);
    if (n <= 1) return first;
    return first + skb_np_pairwise(ld, s + 1, n - 1);
}

// Paths of more than SKB_LONG_PATH entries are left to skb_long_path_stats_kernel (one thread block per path, sums in a
// fixed tree order) in the CUDA build, so that a single 10^7-voxel spiral is not one thread's serial loop.
#define SKB_LONG_PATH 32768
struct SkbPathStatsItem {  // over paths
    const int32_t* pindptr;
    const double* pdata;
    const double* pw;
    double* dist;   // may be null
    double* mean;   // may be null
    double* stdev;  // may be null
    int32_t max_len = 0x7fffffff;  // longer paths are skipped here
    SKB_HD void operator()(int64_t p) const {
        int32_t s = pindptr[p], e = pindptr[p + 1];
        if (e - s > max_len) return;
        if (dist) {
            double acc = 0.0;  // left to right over the path's edges, csr.py:971-977
            for (int32_t j = s + 1; j < e; ++j) acc += pw[j];
            dist[p] = acc;
        }
        if (!pdata) {  // boolean image: every entry of paths.data is 1.0, so the mean is 1 and the variance 0, exactly
            if (mean) mean[p] = 1.0;
            if (stdev) stdev[p] = 0.0;
        } else if (mean || stdev) {
            double n = (double)(e - s);
            double m = skb_np_reduceat(SkbIdLoad{pdata}, s, e - s) / n;
            if (mean) mean[p] = m;
            if (stdev) {
                double sq = skb_np_reduceat(SkbSqLoad{pdata}, s, e - s);
                double var = sq / n - m * m;
                if (!(var > 0.0)) var = (var == var) ? 0.0 : var;  // np.clip(., 0, None); NaN stays NaN
                stdev[p] = skb_sqrt(var);
            }
        }
    }
};

// Branch table columns (csr.py:909-952).  Output is three SoA buffers:
//   i32 [3][P]            skeleton_id, node_id_src, node_id_dst
//   i64 [1 + 2d][P]       branch_type, image_coord_src_0..d-1, image_coord_dst_0..d-1
//   f64 [2(d+h) + 1][P]   coord_src_0..d-1(+h), coord_dst_0..d-1(+h), euclidean_distance
//                         (coord_* hold int64 bit patterns when spacing is integer-typed)
struct SkbBranchItem {
    int64_t P;
    int32_t d;
    int32_t height;       // adds the pixel value as an extra coordinate (csr.py:932-948)
    int32_t spacing_int;  // coord columns stay int64 (csr.py:929 with an integer spacing array)
    const int32_t* pindptr;
    const int32_t* pidx;
    const int32_t* gindptr;
    const int32_t* skel_id;     // per path (SkbCcPathIdItem)
    const SkbAux* head_aux;     // slabs: the last node of a path may live on another rank
    const int32_t* start_node;  // slabs: global id of the first node
    int32_t id_shift;           // slabs: local node id = global id - id_shift
    int64_t gny, gnx;           // slabs: extents of the whole volume (for the last node's coordinates)
    const int64_t* coords;      // (N, d)
    const double* values;       // node values or null
    double sp_f[3];
    int64_t sp_i[3];
    int32_t* out_i32;
    int64_t* out_i64;
    double* out_f64;
    const int32_t* P_ptr = nullptr;  // the number of paths (= the stride of the column arrays) in device memory (overrides P)
    SKB_HD void operator()(int64_t p) const {
        const int64_t P = P_ptr ? (int64_t)*P_ptr : this->P;
        int32_t src, dst, ds, dd;
        int64_t cdst[3] = {0, 0, 0};
        double vdst = 0.0;
        if (head_aux) {
            SkbAux a = head_aux[p];
            src = start_node[p] - id_shift;
            dst = a.end_node;
            dd = a.end_deg;
            vdst = a.end_value;
            int64_t q = a.end_lin / gnx;
            cdst[2] = a.end_lin - q * gnx;
            int64_t q2 = q / gny;
            cdst[1] = q - q2 * gny;
            cdst[0] = q2;
            if (d < 3) {  // (y, x) or (x)
                cdst[0] = cdst[3 - d];
                if (d == 2) cdst[1] = cdst[2];
            }
        } else {
            src = pidx[pindptr[p]];
            dst = pidx[pindptr[p + 1] - 1];
            dd = gindptr[dst + 1] - gindptr[dst];
            for (int i = 0; i < d; ++i) cdst[i] = coords[(int64_t)dst * d + i];
            if (height) vdst = values[dst];
        }
        ds = gindptr[src + 1] - gindptr[src];
        out_i32[p] = skel_id[p];
        out_i32[P + p] = src + id_shift;
        out_i32[2 * P + p] = dst;
        int64_t kind = 2;  // csr.py:916-922, applied in this order
        if (ds == 1 || dd == 1) kind = 1;
        if (ds == 1 && dd == 1) kind = 0;
        if (src + id_shift == dst) kind = 3;
        out_i64[p] = kind;
        int dh = d + (height ? 1 : 0);
        double acc = 0.0;
        for (int i = 0; i < d; ++i) {
            int64_t cs = coords[(int64_t)src * d + i];
            int64_t cd = cdst[i];
            out_i64[(1 + i) * P + p] = cs;
            out_i64[(1 + d + i) * P + p] = cd;
            double diff;
            if (spacing_int) {
                int64_t rs = cs * sp_i[i], rd = cd * sp_i[i];
                memcpy(&out_f64[i * P + p], &rs, 8);
                memcpy(&out_f64[(dh + i) * P + p], &rd, 8);
                diff = (double)(rd - rs);
            } else {
                double rs = (double)cs * sp_f[i], rd = (double)cd * sp_f[i];
                out_f64[i * P + p] = rs;
                out_f64[(dh + i) * P + p] = rd;
                diff = rd - rs;
            }
            acc += diff * diff;
        }
        if (height) {
            double vs = values[src], vd = vdst;
            out_f64[d * P + p] = vs;
            out_f64[(dh + d) * P + p] = vd;
            double diff = vd - vs;
            acc += diff * diff;
        }
        out_f64[2 * dh * P + p] = skb_sqrt(acc);
    }
};

// path_label_image (csr.py:776-790): every path writes id+1 at its voxels; the reference loops
// over paths in order, so where paths share a voxel (junctions) the highest path id wins.
struct SkbLabelImageItem {  // over (path, lane): 32 items share a path, so that one very long path is not one thread's loop
    const int32_t* pindptr;
    const int32_t* pidx;
    const int64_t* lin;
    int64_t* image;  // pre-zeroed, V entries
    SKB_HD void operator()(int64_t i) const {
        const int64_t p = i >> 5;
        for (int32_t j = pindptr[p] + (int32_t)(i & 31); j < pindptr[p + 1]; j += 32) {
            int64_t* dst = image + lin[pidx[j]];
#if defined(__CUDA_ARCH__)
            atomicMax((long long*)dst, (long long)(p + 1));
#else
            if (*dst < p + 1) *dst = p + 1;
#endif
        }
    }
};

#include "skb_sholl.cuh"
#include "skb_mesh.cuh"
