// skan_b200 -- per-item device functions of the node/edge/path proportional stages.
//
// Each functor is "one logical thread per item" and is launched by skb_foreach (a grid-stride
// CUDA kernel on sm_100a; a serial loop in the host emulation harness used by CPU unit tests).
// Reference citations are relative to /root/reference (jni/skan).
//
// Data model (everything addressed by raster-order node id, never by a V-sized id image):
//   bits    packed foreground mask, 1 bit / voxel in C raster order, little-endian inside u64
//           words, one zero pad word before word 0 and after the last word
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
struct SkbGeom {
    int64_t nvox;
    int64_t sz;  // ny * nx
    int32_t nz, ny, nx, ndim;
};

SKB_HD void skb_coords(const SkbGeom& g, int64_t r, int32_t& z, int32_t& y, int32_t& x) {
    if (g.nvox <= 0xffffffffLL) {
        uint32_t rr = (uint32_t)r;
        uint32_t q = rr / (uint32_t)g.nx;
        x = (int32_t)(rr - q * (uint32_t)g.nx);
        uint32_t q2 = q / (uint32_t)g.ny;
        y = (int32_t)(q - q2 * (uint32_t)g.ny);
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
    const uint8_t* b8 = (const uint8_t*)bits;
    int64_t q = p >> 3;
    uint32_t v = (uint32_t)b8[q] | ((uint32_t)b8[q + 1] << 8);
    return (v >> (uint32_t)(p & 7)) & 7u;
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

// calls fn(k27, neighbour_id) for every set bit of mask m of node i, ascending k27
template <typename Fn>
SKB_HD void skb_for_neighbors(const SkbGeom& g, const uint64_t* bits, const uint32_t* pre256, int64_t r, int32_t i,
                              uint32_t m, Fn& fn) {
    for (int row = 0; row < 9; ++row) {
        uint32_t s = (m >> (row * 3)) & 7u;
        if (!s) continue;
        if (row == 4) {
            if (s & 1u) fn(12, i - 1);
            if (s & 4u) fn(14, i + 1);
        } else {
            int dz = row / 3 - 1, dy = row % 3 - 1;
            int64_t rr = r + dz * g.sz + dy * (int64_t)g.nx;
            // m may be a subset of the foreground voxels of this row (edges dropped by the MST):
            // ids advance with the voxels actually set in the mask, not with the bits of m
            int b = skb_ffs(s) - 1;
            int32_t id0 = skb_rank(bits, pre256, rr - 1 + b);
            uint32_t win = skb_win3(bits, rr - 1) >> b;
            for (int t = b; t < 3; ++t)
                if ((s >> t) & 1u) fn(row * 3 + t, id0 + skb_popc(win & ((1u << (t - b)) - 1u)));
        }
    }
}

// ---- stage: raw 3^d neighbourhood mask (csr.py:122-144, evaluated per foreground voxel) ----
struct SkbRawNbItem {
    SkbGeom g;
    const uint64_t* bits;
    const int64_t* lin;
    uint32_t* nb;
    SKB_HD void operator()(int64_t i) const {
        int64_t r = lin[i];
        int32_t z, y, x;
        skb_coords(g, r, z, y, x);
        uint32_t xm = 7u;
        if (x == 0) xm &= ~1u;
        if (x == g.nx - 1) xm &= ~4u;
        uint32_t m = 0;
        for (int dz = -1; dz <= 1; ++dz) {
            if ((uint32_t)(z + dz) >= (uint32_t)g.nz) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                if ((uint32_t)(y + dy) >= (uint32_t)g.ny) continue;
                int64_t rr = r + dz * g.sz + dy * (int64_t)g.nx;
                uint32_t w = skb_win3(bits, rr - 1) & xm;
                m |= w << (((dz + 1) * 3 + (dy + 1)) * 3);
            }
        }
        nb[i] = m & ~(1u << 13);
    }
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
    uint32_t* count;
    SKB_HD void operator()(int64_t i) const {
        uint32_t m = nb[i];
        uint32_t c = 0;
        if (skb_popc(m) >= 3 && (m >> 14)) {
            SkbJuncCountFn fn{nb, 0u};
            skb_for_neighbors(g, bits, pre256, lin[i], (int32_t)i, m & ~((1u << 14) - 1u), fn);
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
    SKB_HD void operator()(int k, int32_t u) {
        if (k > 13 && skb_popc(nb[u]) >= 3) {
            ea[o] = a;
            eb[o] = u;
            ek[o] = (uint8_t)k;
            ew[o] = skb_f64_bits(wf(a, u, k));
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
    SKB_HD void operator()(int64_t i) const {
        uint32_t m = nb[i];
        if (skb_popc(m) >= 3 && (m >> 14) && offset[i + 1] != offset[i]) {
            SkbJuncEmitFn fn{nb, wf, (int32_t)i, offset[i], ea, eb, ek, ew};
            skb_for_neighbors(g, bits, pre256, lin[i], (int32_t)i, m & ~((1u << 14) - 1u), fn);
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

struct SkbMstInitItem {  // over nodes
    int32_t* par;
    uint64_t* bestw;
    uint32_t* beste;
    SKB_HD void operator()(int64_t v) const {
        par[v] = (int32_t)v;
        bestw[v] = ~0ull;
        beste[v] = ~0u;
    }
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
    uint32_t* keep;
    SKB_HD void operator()(int64_t e) const {
        if (tree[e]) return;
        int k = ek[e];
        skb_atomic_and_u32(keep + ea[e], ~(1u << k));
        skb_atomic_and_u32(keep + eb[e], ~(1u << (26 - k)));
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
    SKB_HD void operator()(int k, int32_t u) {
        indices[o] = u;
        data[o] = wf(a, u, k);
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
    SKB_HD void operator()(int64_t i) const {
        uint32_t m = keep[i];
        if (!m) return;
        SkbCsrEmitFn fn{wf, (int32_t)i, indptr[i], indices, data};
        skb_for_neighbors(g, bits, pre256, lin[i], (int32_t)i, m, fn);
    }
};

// ---- stage: connected components (csr.py:909), lock-free union-find, root = minimum id ------
struct SkbIotaItem {
    int32_t* par;
    SKB_HD void operator()(int64_t v) const { par[v] = (int32_t)v; }
};

struct SkbCcHookItem {  // over nodes: union with every smaller neighbour
    const int32_t* indptr;
    const int32_t* indices;
    int32_t* par;
    SKB_HD void operator()(int64_t v) const {
        for (int32_t e = indptr[v]; e < indptr[v + 1]; ++e) {
            int32_t u = indices[e];
            if (u >= (int32_t)v) break;  // sorted row
            int32_t a = (int32_t)v, b = u;
            for (;;) {
                a = skb_find(par, a);
                b = skb_find(par, b);
                if (a == b) break;
                if (a < b) {
                    int32_t t = a;
                    a = b;
                    b = t;
                }
                // larger root goes under the smaller one: the root stays the component minimum
                if (skb_atomic_cas_i32(par + a, a, b) == a) break;
            }
        }
    }
};

// node classes used by the path stages
enum { SKB_NT_ISOLATED = 0, SKB_NT_CHAIN = 1, SKB_NT_TERMINAL = 2, SKB_NT_CYCLE_START = 3 };

struct SkbCcFlattenItem {  // label = root; roots that own a node of degree != 2 are marked
    const int32_t* indptr;
    int32_t* par;
    int32_t* label;
    uint8_t* has_terminal;  // per root, pre-zeroed
    SKB_HD void operator()(int64_t v) const {
        int32_t r = skb_find(par, (int32_t)v);
        label[v] = r;
        if (indptr[v + 1] - indptr[v] != 2) has_terminal[r] = 1;
    }
};

struct SkbNodeTypeItem {
    const int32_t* indptr;
    const int32_t* label;
    const uint8_t* has_terminal;
    uint8_t* ntype;
    uint32_t* is_root;
    SKB_HD void operator()(int64_t v) const {
        int32_t d = indptr[v + 1] - indptr[v];
        int32_t r = label[v];
        uint8_t t;
        if (d == 0) t = SKB_NT_ISOLATED;
        else if (d != 2) t = SKB_NT_TERMINAL;
        else t = (r == (int32_t)v && !has_terminal[r]) ? SKB_NT_CYCLE_START : SKB_NT_CHAIN;
        ntype[v] = t;
        is_root[v] = (r == (int32_t)v) ? 1u : 0u;
    }
};

// ---- stage: path tracing (csr.py:347-446) as list ranking over half-edges --------------------
// arcs[e] for CSR entry e = (u -> v):
//   x >= 0 : successor arc still to be jumped over, y = arcs between e and arcs[e].x
//   x <  0 : finished; -1-x = CSR position of the REVERSED last arc (B -> m) of the list that
//            e lies on, y = number of arcs after e on that list
// A list starts at a terminal (degree 1 or >= 3) or at the minimum node of a junction-free
// cycle (csr.py:369-386) and runs through degree-2 nodes to the next such node.
struct SkbArcInitItem {  // over nodes u
    const int32_t* indptr;
    const int32_t* indices;
    const uint8_t* ntype;
    skb_i2* arcs;
    SKB_HD void operator()(int64_t u) const {
        for (int32_t e = indptr[u]; e < indptr[u + 1]; ++e) {
            int32_t v = indices[e];
            skb_i2 a;
            if (ntype[v] == SKB_NT_CHAIN) {
                int32_t o = indptr[v];
                a.x = (indices[o] == (int32_t)u) ? o + 1 : o;
                a.y = 1;
            } else {
                int32_t j = indptr[v];
                while (indices[j] != (int32_t)u) ++j;  // reverse arc (v -> u)
                a.x = -1 - j;
                a.y = 0;
            }
            arcs[e] = a;
        }
    }
};

struct SkbArcJumpItem {  // over arcs; in-place asynchronous pointer jumping
    skb_i2* arcs;
    int32_t* flag;
    int32_t jumps;
    SKB_HD void operator()(int64_t e) const {
        skb_i2 a = skb_ld_cg(arcs + e);
        if (a.x < 0) return;
        for (int j = 0; j < jumps && a.x >= 0; ++j) {
            skb_i2 s = skb_ld_cg(arcs + a.x);  // 8-byte atomic read: always a consistent pair
            a.x = s.x;
            a.y += s.y;
        }
        skb_st_cg(arcs + e, a);
        if (a.x >= 0) skb_raise(flag);
    }
};

// a half-edge (A,n) starts an emitted path iff it precedes the reversed last half-edge (B,m)
// in CSR order (the serial walk of csr.py:347-409 visits start nodes and their
// neighbours ascending and marks both directions visited)
struct SkbPathCountItem {  // over nodes: lo32 = regular paths starting here, hi32 = cycle paths
    const int32_t* indptr;
    const uint8_t* ntype;
    const skb_i2* arcs;
    uint64_t* count;
    SKB_HD void operator()(int64_t u) const {
        uint8_t t = ntype[u];
        uint64_t c = 0;
        if (t >= SKB_NT_TERMINAL) {
            for (int32_t e = indptr[u]; e < indptr[u + 1]; ++e) {
                int32_t y = -1 - arcs[e].x;
                if (e < y) c += (t == SKB_NT_TERMINAL) ? 1ull : (1ull << 32);
            }
        }
        count[u] = c;
    }
};

struct SkbPathStartItem {  // over nodes: path id, start node, length; tag the list by its end arc
    const int32_t* indptr;
    const uint8_t* ntype;
    const skb_i2* arcs;
    const uint64_t* prefix;  // exclusive scan of SkbPathCountItem counts
    uint32_t n_regular;
    int32_t* start_node;
    uint32_t* plen;
    skb_i2* einfo;  // pre-set to -1; einfo[y] = (path id, arcs on the path)
    SKB_HD void operator()(int64_t u) const {
        uint8_t t = ntype[u];
        if (t < SKB_NT_TERMINAL) return;
        uint64_t p = prefix[u];
        uint32_t pid = (t == SKB_NT_TERMINAL) ? (uint32_t)(p & 0xffffffffull) : n_regular + (uint32_t)(p >> 32);
        for (int32_t e = indptr[u]; e < indptr[u + 1]; ++e) {
            skb_i2 a = arcs[e];
            int32_t y = -1 - a.x;
            if (e < y) {
                start_node[pid] = (int32_t)u;
                plen[pid] = (uint32_t)a.y + 2u;
                skb_i2 info;
                info.x = (int32_t)pid;
                info.y = a.y + 1;
                einfo[y] = info;
                ++pid;
            }
        }
    }
};

struct SkbPathEmitItem {  // over arcs: every arc of an emitted list writes its head node
    const int32_t* indices;
    const skb_i2* arcs;
    const skb_i2* einfo;
    const int32_t* pindptr;
    const int32_t* start_node;
    const double* values;  // null -> 1.0 (csr.py:207-215)
    int32_t* pidx;
    double* pdata;
    SKB_HD void operator()(int64_t e) const {
        skb_i2 a = arcs[e];
        skb_i2 info = einfo[-1 - a.x];
        if (info.x < 0) return;
        int32_t base = pindptr[info.x];
        int32_t pos = info.y - a.y;
        int32_t v = indices[e];
        pidx[base + pos] = v;
        pdata[base + pos] = values ? values[v] : 1.0;
        if (pos == 1) {
            int32_t s = start_node[info.x];
            pidx[base] = s;
            pdata[base] = values ? values[s] : 1.0;
        }
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

struct SkbPathStatsItem {  // over paths
    const int32_t* gindptr;
    const int32_t* gindices;
    const double* gdata;
    const int32_t* pindptr;
    const int32_t* pidx;
    const double* pdata;
    double* dist;   // may be null
    double* mean;   // may be null
    double* stdev;  // may be null
    SKB_HD void operator()(int64_t p) const {
        int32_t s = pindptr[p], e = pindptr[p + 1];
        if (dist) {
            double acc = 0.0;  // left-to-right, csr.py:971-977
            for (int32_t j = s; j + 1 < e; ++j) {
                int32_t u = pidx[j], v = pidx[j + 1];
                double w = 0.0;
                for (int32_t r = gindptr[u]; r < gindptr[u + 1]; ++r)
                    if (gindices[r] == v) {
                        w = gdata[r];
                        break;
                    }
                acc += w;
            }
            dist[p] = acc;
        }
        if (mean || stdev) {
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
    const int32_t* label;
    const uint32_t* root_rank;  // exclusive scan of is_root
    const int64_t* coords;      // (N, d)
    const double* values;       // node values or null
    double sp_f[3];
    int64_t sp_i[3];
    int32_t* out_i32;
    int64_t* out_i64;
    double* out_f64;
    SKB_HD void operator()(int64_t p) const {
        int32_t src = pidx[pindptr[p]];
        int32_t dst = pidx[pindptr[p + 1] - 1];
        out_i32[p] = (int32_t)root_rank[label[src]];
        out_i32[P + p] = src;
        out_i32[2 * P + p] = dst;
        int32_t ds = gindptr[src + 1] - gindptr[src];
        int32_t dd = gindptr[dst + 1] - gindptr[dst];
        int64_t kind = 2;  // csr.py:916-922, applied in this order
        if (ds == 1 || dd == 1) kind = 1;
        if (ds == 1 && dd == 1) kind = 0;
        if (src == dst) kind = 3;
        out_i64[p] = kind;
        int dh = d + (height ? 1 : 0);
        double acc = 0.0;
        for (int i = 0; i < d; ++i) {
            int64_t cs = coords[(int64_t)src * d + i];
            int64_t cd = coords[(int64_t)dst * d + i];
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
            double vs = values[src], vd = values[dst];
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
struct SkbLabelImageItem {
    const int32_t* pindptr;
    const int32_t* pidx;
    const int64_t* lin;
    int64_t* image;  // pre-zeroed, V entries
    SKB_HD void operator()(int64_t p) const {
        for (int32_t j = pindptr[p]; j < pindptr[p + 1]; ++j) {
            int64_t* dst = image + lin[pidx[j]];
#if defined(__CUDA_ARCH__)
            atomicMax((long long*)dst, (long long)(p + 1));
#else
            if (*dst < p + 1) *dst = p + 1;
#endif
        }
    }
};
