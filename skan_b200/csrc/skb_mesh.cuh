// skan_b200 -- image_stats.mesh_sizes (reference image_stats.py:8-46): areas of the background regions a 2-D
// skeleton encloses, and the number of connected foreground components (image_stats.py:86-87).
// Included by skb_items.cuh; entry points in skb_api.inl.
//
// scipy.ndimage.label(~skeleton) with its default structure labels the 4-connected background components in
// raster order of their first pixel.  Here the background is handled as horizontal RUNS (maximal sequences of
// zero bits within a row of the packed mask): with a density of a few per cent there are about as many runs as
// skeleton pixels, so nothing pixel-sized beyond the 1-bit mask is ever built.
//   1. run starts: a second bit mask `rs` (bit p set <=> p is background and starts a run) + popcount prefix per
//      word -> run id = rank of the start bit = raster order of the run's first pixel;
//   2. every run unites with the runs of the row above that it overlaps (their ids are consecutive: two rank
//      queries), union by minimum id, so a component's root is its first run in raster order;
//   3. sizes are summed at the roots, components with a run on the image border are dropped, the rest is
//      compacted in root order == scipy's label order.
#pragma once

struct SkbMeshGeom {
    int64_t nvox;
    int32_t ny, nx;
};

// bit p of the mask (p in [-1, nvox]: the pad words in front of and behind the mask are zero)
SKB_HD uint32_t skb_mask_bit(const uint64_t* bits, int64_t p) { return (uint32_t)((bits[p >> 6] >> (p & 63)) & 1ull); }

// run-start word w of a 2-D image: background pixels whose left neighbour is foreground or that sit in column 0
struct SkbRunStartItem {  // over mask words
    SkbMeshGeom g;
    const uint64_t* bits;
    uint64_t* rs;
    uint32_t* count;  // run starts per word (scanned by the caller)
    SKB_HD void operator()(int64_t w) const {
        const int64_t p0 = w * 64;
        const uint64_t fg = bits[w];
        const uint64_t prev_fg = (fg << 1) | (w > 0 ? bits[w - 1] >> 63 : 0ull);  // bit b: pixel p0 + b - 1 is foreground
        uint64_t valid = ~0ull;
        if (p0 + 64 > g.nvox) valid = g.nvox > p0 ? ((1ull << (g.nvox - p0)) - 1ull) : 0ull;
        // bits at column 0: x = (p0 + b) mod nx == 0
        uint64_t col0 = 0ull;
        int64_t first = ((p0 + g.nx - 1) / g.nx) * (int64_t)g.nx;  // first multiple of nx >= p0
        for (int64_t q = first; q < p0 + 64; q += g.nx) col0 |= 1ull << (q - p0);
        const uint64_t start = ~fg & (prev_fg | col0) & valid;
        rs[w] = start;
        count[w] = (uint32_t)skb_popcll(start);
    }
};

// number of run starts at positions <= q
SKB_HD uint32_t skb_runs_upto(const uint64_t* rs, const uint32_t* rsp, int64_t q) {
    const int64_t w = q >> 6;
    const int b = (int)(q & 63);
    const uint64_t m = b == 63 ? ~0ull : ((1ull << (b + 1)) - 1ull);
    return rsp[w] + (uint32_t)skb_popcll(rs[w] & m);
}

// per run: extent, length, border flag; union with the overlapping runs of the row above
struct SkbRunLinkItem {  // over mask words (every set bit of rs[w] is one run)
    SkbMeshGeom g;
    const uint64_t* bits;
    const uint64_t* rs;
    const uint32_t* rsp;  // exclusive scan of the per-word counts
    int32_t* par;         // one entry per run, initialised to its own id
    uint32_t* len;        // pixels of every run
    uint8_t* border;      // 1: the run touches the image border
    SKB_HD void operator()(int64_t w) const {
        uint64_t starts = rs[w];
        uint32_t id = rsp[w];
        while (starts) {
            const int b = skb_ffsll(starts) - 1;
            starts &= starts - 1ull;
            const int64_t p = w * 64 + b;
            const int64_t y = p / g.nx;
            const int64_t row_end = (y + 1) * (int64_t)g.nx - 1;
            // end of the run: the pixel before the next foreground pixel of the row, or the row end
            int64_t e = row_end;
            {
                int64_t ww = p >> 6;
                uint64_t word = bits[ww] & ~((1ull << (p & 63)) - 1ull);  // foreground at positions >= p (p itself is background)
                for (;;) {
                    if (word) {
                        int64_t f = ww * 64 + (skb_ffsll(word) - 1);
                        if (f <= row_end) e = f - 1;
                        break;
                    }
                    ++ww;
                    if (ww * 64 > row_end) break;
                    word = bits[ww];
                }
            }
            len[id] = (uint32_t)(e - p + 1);
            const int64_t x0 = p - y * g.nx, x1 = e - y * g.nx;
            border[id] = (uint8_t)(y == 0 || y == g.ny - 1 || x0 == 0 || x1 == g.nx - 1);
            if (y > 0) {
                const int64_t q0 = p - g.nx, q1 = e - g.nx;
                const uint32_t upto0 = skb_runs_upto(rs, rsp, q0);
                // first candidate: the run that contains q0 when q0 is background, else the next one to start
                const int64_t a = skb_mask_bit(bits, q0) ? (int64_t)upto0 : (int64_t)upto0 - 1;
                const int64_t bb = (int64_t)skb_runs_upto(rs, rsp, q1) - 1;
                for (int64_t j = a; j <= bb; ++j) skb_union_min(par, (int32_t)id, (int32_t)j);
            }
            ++id;
        }
    }
};

struct SkbRunBorderItem {  // over runs, after all unions: a component with a run on the border is dropped
    int32_t* par;
    const uint8_t* border;
    uint32_t* root_border;  // per run, zeroed; meaningful at roots
    SKB_HD void operator()(int64_t r) const {
        if (border[r]) root_border[skb_find(par, (int32_t)r)] = 1u;
    }
};

struct SkbRunSizeItem {  // over runs: sizes gathered at the roots of the components that are kept (the space around
                         // the skeleton, by far the largest component, touches the border: no atomics for its runs)
    int32_t* par;
    const uint32_t* len;
    const uint32_t* root_border;
    unsigned long long* size;  // per run, zeroed; meaningful at roots
    SKB_HD void operator()(int64_t r) const {
        const int32_t root = skb_find(par, (int32_t)r);
        if (!root_border[root]) skb_atomic_add_u64(size + root, (unsigned long long)len[r]);
    }
};

struct SkbRunKeepItem {  // over runs: 1 for the root of a component that does not touch the border
    const int32_t* par;
    const uint32_t* root_border;
    uint32_t* keep;
    SKB_HD void operator()(int64_t r) const { keep[r] = (par[r] == (int32_t)r && !root_border[r]) ? 1u : 0u; }
};

struct SkbRunEmitItem {  // over runs: compacted sizes in root order (== scipy.ndimage.label's label order)
    const int32_t* par;
    const uint32_t* root_border;
    const uint32_t* keepoff;  // exclusive scan of keep
    const unsigned long long* size;
    int64_t* out;
    SKB_HD void operator()(int64_t r) const {
        if (par[r] == (int32_t)r && !root_border[r]) out[keepoff[r]] = (int64_t)size[r];
    }
};

// foreground pixel on the image border?  (decides whether the count of foreground pixels stays in the result:
// image_stats.py:41-44 zeroes sizes[0] only when label 0 is among the border labels)
struct SkbBorderFgItem {  // over 2 * (nx + ny) border positions
    SkbMeshGeom g;
    const uint64_t* bits;
    int32_t* flag;
    SKB_HD void operator()(int64_t i) const {
        int64_t p;
        if (i < g.nx) p = i;                                                      // first row
        else if (i < 2 * (int64_t)g.nx) p = (int64_t)(g.ny - 1) * g.nx + (i - g.nx);  // last row
        else if (i < 2 * (int64_t)g.nx + g.ny) p = (i - 2 * (int64_t)g.nx) * g.nx;    // first column
        else p = (i - 2 * (int64_t)g.nx - g.ny) * g.nx + (g.nx - 1);                 // last column
        if (skb_mask_bit(bits, p)) skb_raise(flag);
    }
};

// ---- connected components of the whole pixel graph (image_stats.py:86-87: ndi.label with the full structure) ----
struct SkbCcEdgeUnionItem {  // over nodes: unite with the smaller neighbours
    const int32_t* indptr;
    const int32_t* indices;
    int32_t* par;  // initialised to iota
    SKB_HD void operator()(int64_t i) const {
        for (int32_t e = indptr[i]; e < indptr[i + 1]; ++e) {
            const int32_t j = indices[e];
            if (j < (int32_t)i) skb_union_min(par, (int32_t)i, j);
        }
    }
};

struct SkbCcRootCountItem {  // over nodes
    const int32_t* par;
    unsigned long long* count;
    SKB_HD void operator()(int64_t i) const {
        if (par[i] == (int32_t)i) skb_atomic_add_u64(count, 1ull);
    }
};
