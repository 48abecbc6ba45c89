// skan_b200 -- native sequencing of the whole hot path for one image (or one stack of 2-D images):
// mask sweep -> nodes -> raw stencil -> junction MST -> CSR (+ node records) -> paths -> statistics
// -> branch table, in ONE call.  It is the C++ twin of skan_b200/_pipeline.py: the same entry points
// in the same order, but allocation is a bump pointer into two caller-provided device arenas and the
// handful of counts that size the next allocation are read back here, so a call costs ~8 stream
// synchronisations and no interpreter time.  Included after skb_api.inl by skb_cuda.cu and by
// tests/emul/skb_emul.cpp (the including file provides skb_copy_bytes and skb_fetch_i32).
//
// Arenas: `work` holds scratch (mask, rank structure, neighbourhood masks, junction edges, walk
// starts ...), `out` holds everything the caller may read afterwards.  When an arena is too small
// the call returns SKB_RUN_NEED_MORE with the sizes needed so far in result->work_need / out_need;
// the caller grows the arena and calls again.

#define SKB_RUN_NEED_MORE 1
#define SKB_MASK_PAD 16  // zero words in front of the mask (one 128-byte line)
#define SKB_SMALL_CYCLE_LIMIT 4096  // uncovered chain voxels one thread block handles (skb_small_cycles_kernel)

// indices into skb_run_result.off (byte offsets into the `out` arena; -1 = absent)
enum {
    SKB_O_LIN = 0, SKB_O_COORDS, SKB_O_VALUES, SKB_O_DEGREES, SKB_O_INDPTR, SKB_O_INDICES, SKB_O_DATA,
    SKB_O_PINDPTR, SKB_O_PIDX, SKB_O_PDATA, SKB_O_PW, SKB_O_SKEL, SKB_O_DIST, SKB_O_MEAN, SKB_O_STDEV,
    SKB_O_T32, SKB_O_T64, SKB_O_TF, SKB_O_COMP_RANK, SKB_O_COUNT
};
// indices into skb_run_result.counts
enum {
    SKB_C_NODES = 0, SKB_C_EDGES, SKB_C_PATHS, SKB_C_REGULAR, SKB_C_CYCLES, SKB_C_ENTRIES, SKB_C_JUNCTION_EDGES,
    SKB_C_STARTS, SKB_C_SYNCS, SKB_C_SWEEP_NS, SKB_C_OVERLAPPED, SKB_C_RETAKEN, SKB_C_JUNCTION_NODES, SKB_C_COUNT
};
// indices into skb_run_params.hint: capacities for the arrays sized by a count that is still on its way to the
// host when their first consumer is launched (0 = no hint: wait for the count, then launch)
enum { SKB_H_JUNCTION_EDGES = 0, SKB_H_EDGES, SKB_H_STARTS, SKB_H_REGULAR, SKB_H_JUNCTION_NODES, SKB_H_COUNT };
// stage boundaries of skb_run_result.stage_ns (time_sweep == 2): stage_ns[i] = device time between
// the previous boundary and boundary i, host synchronisation bubbles included
enum {
    SKB_M_START = 0, SKB_M_SWEEP, SKB_M_NODES, SKB_M_RAW, SKB_M_JCOUNT, SKB_M_MST, SKB_M_DEGREES, SKB_M_CSR,
    SKB_M_WALK, SKB_M_RANK, SKB_M_SELECT, SKB_M_HEADS, SKB_M_EMIT, SKB_M_CYCLES, SKB_M_SKEL, SKB_M_STATS,
    SKB_M_TABLE, SKB_M_COUNT
};

struct skb_run_params {
    const void* image;
    int32_t dtype, ndim;
    int64_t shape[3];
    int64_t z_origin;
    int32_t image_stack;  // leading axis = independent 2-D images
    int32_t height;       // value_is_height
    int32_t pack_variant;
    int32_t stop_after_graph;  // 1: skeleton_to_csgraph only
    int32_t want_values;       // fp64 node values (non-boolean images)
    int32_t time_sweep;        // 1: counts[SKB_C_SWEEP_NS] = device time of the mask sweep kernel (events);
                               // 2: also stage_ns[] (one event per stage boundary)
    const double* wtab;        // device, 27 weights
    double spacing_f[3];       // for the branch table (table_ndim entries)
    int64_t spacing_i[3];
    int32_t spacing_is_int, table_ndim;
    void* work;
    int64_t work_bytes;
    void* out;
    int64_t out_bytes;
    void* stream;
    int64_t node_cap_hint;  // capacity hint for the node arrays (fused sweep: 0 = V/256 + 4096; else 0 = no hint)
    int64_t hint[8];        // SKB_H_*: capacity hints, typically the counts of the previous image of this kind plus
                            // slack.  With a hint the read-back of the count overlaps the first kernel that fills
                            // the arrays it sizes; a hint that turns out too small costs that kernel a second launch.
};

struct skb_run_result {
    int64_t counts[16];
    int64_t off[32];
    int64_t work_need, out_need;
    int64_t stage_ns[32];
};

// Debugging aid (compute-sanitizer is closed on the pool this was built on): with skb_set_option("arena_guard", 1)
// every allocation of a runner is followed by a 256-byte guard gap whose offset is recorded; the caller fills the arenas
// with a byte pattern before the call and checks afterwards that every gap still holds it (an out-of-bounds write of a
// kernel lands in a gap), and runs twice with different patterns (a result that depends on the pattern was computed
// from uninitialised memory).  tests/test_gpu_guards.py does both.
#define SKB_GUARD_BYTES 256
#define SKB_MAX_GUARDS 512
// g_skb_arena_guard is defined by the including file (skb_set_option "arena_guard")
struct SkbGuardLog {
    int64_t off[2][SKB_MAX_GUARDS];
    int n[2];
};
static thread_local SkbGuardLog g_skb_guards = {};

struct SkbArena {
    uint8_t* base;
    int64_t cap, off, need;
    int64_t high;  // high-water mark of `off`: a gap recorded below it lies in space an earlier (rewound) allocation used
    bool ok;
    int which;  // 0 = work, 1 = out (index into the guard log)
    void init(void* p, int64_t bytes, int which_ = 0) {
        base = (uint8_t*)p;
        cap = bytes;
        off = 0;
        need = 0;
        high = 0;
        ok = true;
        which = which_;
        g_skb_guards.n[which] = 0;
    }
    template <typename T>
    T* take(int64_t count) {
        int64_t bytes = (count > 0 ? count : 1) * (int64_t)sizeof(T);
        off = (off + 255) & ~(int64_t)255;
        const int64_t guard = g_skb_arena_guard ? SKB_GUARD_BYTES : 0;
        need = off + bytes + guard;
        if (need > cap || !base) {
            ok = false;
            off = need;  // keep counting so that the caller learns how much this stage wanted
            return nullptr;
        }
        T* p = (T*)(base + off);
        if (guard) {
            // (a runner that rewinds the bump pointer allocates over earlier gaps: drop the records at or beyond `off`)
            int& n = g_skb_guards.n[which];
            while (n > 0 && (g_skb_guards.off[which][n - 1] >= off || -g_skb_guards.off[which][n - 1] - 1 >= off)) --n;
            const int64_t gap = (off + bytes + 255) & ~(int64_t)255;
            // a gap in space that was handed out before the runner rewound its bump pointer may have been written to
            // legitimately: recorded as -(offset) - 1, which the checker skips
            if (n < SKB_MAX_GUARDS) g_skb_guards.off[which][n++] = gap < high ? -gap - 1 : gap;
            need = ((off + bytes + 255) & ~(int64_t)255) + guard;
            if (need > cap) {
                ok = false;
                off = need;
                return nullptr;
            }
        }
        off = need;
        if (off > high) high = off;
        return p;
    }
};

#define SKB_NEED(arena_ok)                                                     \
    do {                                                                       \
        if (!(arena_ok)) {                                                     \
            res->work_need = work.need + work.need / 2 + (1 << 20);            \
            res->out_need = out.need + out.need / 2 + (1 << 20);               \
            return SKB_RUN_NEED_MORE;                                          \
        }                                                                      \
    } while (0)

#define SKB_FETCH(dst, ptr)                                   \
    do {                                                      \
        int32_t _v = 0;                                       \
        SKB_TRY(skb_fetch_i32((const int32_t*)(ptr), &_v, s)); \
        (dst) = _v;                                           \
        ++nsync;                                              \
        if (_v < 0) return skb_fail("a count exceeds int32: the reference's int32 index arrays cannot hold this graph"); \
    } while (0)

// posted read-back of one count (see skb_fetch_post): the kernels launched before SKB_WAIT overlap its round trip
#define SKB_POST(ptr)                                       \
    do {                                                    \
        const int32_t* _p[1] = {(const int32_t*)(ptr)};     \
        SKB_TRY(skb_fetch_post(_p, 1, s));                  \
    } while (0)
#define SKB_WAIT(dst)                         \
    do {                                      \
        int64_t _v[1] = {0};                  \
        SKB_TRY(skb_fetch_wait(1, _v, s));    \
        (dst) = _v[0];                        \
        ++nsync;                              \
        if (_v[0] < 0) return skb_fail("a count exceeds int32: the reference's int32 index arrays cannot hold this graph"); \
    } while (0)

extern "C" int skb_run_image(const skb_run_params* prm, skb_run_result* res) {
    skb_stream_t s = (skb_stream_t)prm->stream;
    const int ndim = prm->ndim;
    SKB_TRY(skb_check_geom(ndim, prm->shape));
    if (prm->image_stack && ndim != 3) return skb_fail("a stack of 2-D images is a 3-D array");
    SkbGeom g = skb_make_geom(ndim, prm->shape);
    const int64_t V = g.nvox;
    const int64_t ntiles = (V + SKB_TILE_VOX - 1) / SKB_TILE_VOX;
    int64_t nsync = 0, noverlap = 0, nretaken = 0;
    const int both_ways = skb_option_both_ways();  // option: second walk written from both ends of every chain
    for (int i = 0; i < 16; ++i) res->counts[i] = 0;
    for (int i = 0; i < 32; ++i) res->off[i] = -1;
    res->work_need = res->out_need = 0;
    for (int i = 0; i < 32; ++i) res->stage_ns[i] = 0;
    const bool stages = prm->time_sweep == 2;
    uint8_t marked[32] = {};
#define SKB_MARK(i)                           \
    do {                                      \
        if (stages) {                         \
            SKB_TRY(skb_stage_mark((i), s));  \
            marked[(i)] = 1;                  \
        }                                     \
    } while (0)
#define SKB_MARKS_DONE()                                                          \
    do {                                                                          \
        if (stages) SKB_TRY(skb_stage_read(marked, SKB_M_COUNT, res->stage_ns));  \
    } while (0)
    SkbArena work, out;
    work.init(prm->work, prm->work_bytes, 0);
    out.init(prm->out, prm->out_bytes, 1);
    // every successful return reports what the image needed: the caller sizes the next image's arenas by it
#define SKB_RUN_DONE()                            \
    do {                                          \
        res->counts[SKB_C_SYNCS] = nsync;         \
        res->counts[SKB_C_OVERLAPPED] = noverlap; \
        res->counts[SKB_C_RETAKEN] = nretaken;    \
        res->work_need = work.need;               \
        res->out_need = out.need;                 \
    } while (0)
    auto out_off = [&](const void* p) -> int64_t { return p ? (int64_t)((const uint8_t*)p - out.base) : -1; };

    // ---- mask sweep, scan, nodes (one fused pass, or sweep -> tile scan -> node emission) ---------------
    const bool fused = prm->pack_variant == 2;
    const bool want_values = prm->want_values != 0;
    uint64_t* store = work.take<uint64_t>(ntiles * SKB_TILE_WORDS + SKB_MASK_PAD + 2);
    uint32_t* tile_prefix = fused ? nullptr : work.take<uint32_t>(ntiles + 1);
    void* scratch = fused ? work.take<uint8_t>(skb_sweep_scratch_bytes(V))
                          : work.take<uint8_t>(skb_scan_scratch_bytes(ntiles + 1));
    int32_t* flag = work.take<int32_t>(4);
    int32_t* n_dev = work.take<int32_t>(4);
    uint32_t* pre256 = work.take<uint32_t>(ntiles * 64);
    SKB_NEED(work.ok);
    // the mask starts one 128-byte line into its (256-byte aligned) allocation: every 256-bit block of the rank
    // structure is one 32-byte sector, every tile a whole number of lines; the words before and after it are zero
    uint64_t* bits = store + SKB_MASK_PAD;
    SKB_TRY(skb_fill_bytes(store, 0, SKB_MASK_PAD * 8, s));
    SKB_TRY(skb_fill_bytes(store + SKB_MASK_PAD + ntiles * SKB_TILE_WORDS, 0, 16, s));
    void* ev = nullptr;
    int64_t N;
    int64_t *lin = nullptr, *coords = nullptr;
    double* values = nullptr;
    if (fused) {
        int64_t node_cap = prm->node_cap_hint > 0 ? prm->node_cap_hint : V / 256 + 4096;
        if (node_cap > V) node_cap = V;
        const int64_t out_mark = out.off;
        lin = out.take<int64_t>(node_cap);
        coords = out.take<int64_t>(node_cap * ndim);
        values = want_values ? out.take<double>(node_cap) : nullptr;
        SKB_NEED(out.ok);
        SKB_MARK(SKB_M_START);
        if (prm->time_sweep) SKB_TRY(skb_timer_start(&ev, s));
        SKB_TRY(skb_sweep_nodes(prm->image, prm->dtype, ndim, prm->shape, prm->z_origin, bits, ntiles, pre256, node_cap,
                                lin, coords, values, scratch, n_dev, s));
        if (prm->time_sweep) SKB_TRY(skb_timer_stop(ev, s));
        SKB_MARK(SKB_M_SWEEP);
        {
            const int32_t* ptrs[2] = {n_dev, n_dev + 1};
            int64_t got[2];
            SKB_TRY(skb_fetch_i32s(ptrs, 2, got, s));
            ++nsync;
            N = got[0];
            if (got[1] != 0) return skb_fail("fused sweep: a wait inside the kernel timed out (chunk hand-off protocol)");
        }
        if (N > node_cap) {  // the capacity guess was too small: emit again from the finished mask
            out.off = out_mark;
            lin = out.take<int64_t>(N);
            coords = out.take<int64_t>(N * ndim);
            values = want_values ? out.take<double>(N) : nullptr;
            tile_prefix = work.take<uint32_t>(ntiles + 1);
            SKB_NEED(work.ok && out.ok);
            SKB_TRY(skb_tile_prefix(pre256, n_dev, ntiles, tile_prefix, s));
            SKB_TRY(skb_emit_nodes(bits, tile_prefix, ntiles, ndim, prm->shape, prm->image, prm->dtype, prm->z_origin,
                                   lin, coords, values, pre256, s));
        }
    } else {
        SKB_TRY(skb_fill_bytes(scratch, 0, (size_t)skb_scan_scratch_bytes(ntiles + 1), s));
        SKB_MARK(SKB_M_START);
        if (prm->time_sweep) SKB_TRY(skb_timer_start(&ev, s));
        SKB_TRY(skb_pack_count(prm->image, prm->dtype, V, bits, tile_prefix, ntiles, prm->pack_variant, s));
        if (prm->time_sweep) SKB_TRY(skb_timer_stop(ev, s));
        SKB_TRY(skb_scan_u32_impl(tile_prefix, tile_prefix, ntiles, scratch, s));
        SKB_MARK(SKB_M_SWEEP);
        int64_t cap = prm->node_cap_hint > 0 ? prm->node_cap_hint : 0;
        if (cap > V) cap = V;
        const int64_t out_mark = out.off;
        bool emitted = false;
        SKB_POST(tile_prefix + ntiles);
        if (cap > 0) {  // node emission sized by the hint runs while the count travels
            lin = out.take<int64_t>(cap);
            coords = out.take<int64_t>(cap * ndim);
            values = want_values ? out.take<double>(cap) : nullptr;
            if (out.ok) {
                SKB_TRY(skb_emit_nodes_cap(bits, tile_prefix, ntiles, ndim, prm->shape, prm->image, prm->dtype,
                                           prm->z_origin, lin, coords, values, pre256, cap, s));
                emitted = true;
                ++noverlap;
            }
        }
        SKB_WAIT(N);
        if (!emitted || N > cap) {
            if (emitted) ++nretaken;
            out.off = out_mark;
            out.ok = true;
            lin = out.take<int64_t>(N);
            coords = out.take<int64_t>(N * ndim);
            values = want_values ? out.take<double>(N) : nullptr;
            SKB_NEED(out.ok);
            SKB_TRY(skb_emit_nodes(bits, tile_prefix, ntiles, ndim, prm->shape, prm->image, prm->dtype, prm->z_origin,
                                   lin, coords, values, pre256, s));
        }
    }
    if (prm->time_sweep) SKB_TRY(skb_timer_read_ns(ev, &res->counts[SKB_C_SWEEP_NS]));
    res->counts[SKB_C_NODES] = N;
    SKB_MARK(SKB_M_NODES);
    {   // every later scan runs over at most 27 N items (nodes, half-edges, walk starts, paths)
        const int64_t bytes = skb_scan_scratch_bytes(27 * N + ntiles + 1);
        scratch = work.take<uint8_t>(bytes);
        SKB_NEED(work.ok);
        SKB_TRY(skb_fill_bytes(scratch, 0, (size_t)bytes, s));
    }
    // ---- raw neighbourhood -----------------------------------------------------------------------------
    uint32_t* nb = work.take<uint32_t>(N);
    uint32_t* jcount = work.take<uint32_t>(N + 1);
    SKB_NEED(work.ok);
    SKB_TRY(skb_raw_neighbors(bits, ndim, prm->shape, lin, N, prm->image_stack, nb, s));
    SKB_MARK(SKB_M_RAW);
    res->off[SKB_O_LIN] = out_off(lin);
    res->off[SKB_O_COORDS] = out_off(coords);
    res->off[SKB_O_VALUES] = out_off(values);
    // ---- junction MST ----------------------------------------------------------------------------------
    const int height = prm->height == 1 ? 1 : 0;
    int64_t nej = 0;
    uint32_t* keep = nb;
    int32_t *ea = nullptr, *eb = nullptr;
    uint8_t *ek = nullptr, *tree = nullptr;
    uint64_t* ew = nullptr;
    int32_t *mpar = nullptr, *era = nullptr, *erb = nullptr;
    uint64_t* bestw = nullptr;
    int64_t n_junction_nodes = 0;
#ifdef SKB_HAVE_FUSED_LOOPS
    SkbDense dense{};  // dense ids of the junction voxels for the Boruvka arrays (capacity hint: their number last time)
#endif
    auto take_junction = [&](int64_t cap, int64_t jcap) {
        ea = work.take<int32_t>(cap);
        eb = work.take<int32_t>(cap);
        ek = work.take<uint8_t>(cap);
        ew = work.take<uint64_t>(cap);
        mpar = work.take<int32_t>(jcap > 0 ? jcap : N);
        bestw = work.take<uint64_t>(2 * (jcap > 0 ? jcap : N));
        era = work.take<int32_t>(cap);
        erb = work.take<int32_t>(cap);
        tree = work.take<uint8_t>(cap);
#ifdef SKB_HAVE_FUSED_LOOPS
        dense = SkbDense{};
        // the junction voxels are counted (the next image's capacity hint) only where dense ids can be used: the count
        // costs every thread block of the append pass an atomic it otherwise needs only when it has edges
        dense.jcounter = N > g_skb_dense_mst_min_nodes ? n_dev + 2 : nullptr;
        if (jcap > 0) {
            dense.jid = work.take<int32_t>(N);
            dense.eda = work.take<int32_t>(cap);
            dense.edb = work.take<int32_t>(cap);
            dense.j_cap = jcap;
        }
#endif
    };
    {
        const int64_t cap = prm->hint[SKB_H_JUNCTION_EDGES] > 0 ? prm->hint[SKB_H_JUNCTION_EDGES] : 0;
        const int64_t work_mark = work.off;
        bool emitted = false, appended = false;
        if (cap > 0 && !height) {
            // one pass, no order (skb_junction_append_kernel): the edges are appended with an atomic counter, and the
            // Boruvka launch reads their number from that counter while it travels to the host
#ifdef SKB_HAVE_FUSED_LOOPS
            // dense ids for the Boruvka arrays pay when the node-indexed arrays (20 bytes per node) are far larger than L2:
            // measured (round 2) -0.16 ms at 30 M nodes / 2.65 M junction edges, +-0.01 ms at 0.8-7.5 M nodes
            const int64_t jcap = (prm->hint[SKB_H_JUNCTION_NODES] > 0 && N > g_skb_dense_mst_min_nodes) ? prm->hint[SKB_H_JUNCTION_NODES] : 0;
            take_junction(cap, jcap);
            if (work.ok) {
                int r = skb_junction_append(bits, pre256, ndim, prm->shape, lin, nb, N, 0, N, ea, eb, ek, cap, flag + 3, s, nullptr,
                                            &dense);
                if (r < 0) return r;
                appended = r == 0;
            }
            if (appended) {
                const int32_t* ptrs[2] = {flag + 3, dense.jcounter ? dense.jcounter : flag + 3};
                SKB_TRY(skb_fetch_post(ptrs, 2, s));
                SKB_TRY(skb_mst_keys_run(ea, eb, ek, prm->wtab, cap, flag + 3, mpar, (unsigned long long*)bestw, N, era, erb,
                                         tree, flag, keep, 0, (int32_t)N, s, &dense));
                ++noverlap;
                int64_t got[2] = {0, 0};
                SKB_TRY(skb_fetch_wait(2, got, s));
                ++nsync;
                if (got[0] < 0 || got[1] < 0) return skb_fail("a count exceeds int32: the reference's int32 index arrays cannot hold this graph");
                nej = got[0];
                n_junction_nodes = dense.jcounter ? got[1] : 0;
                if (nej > cap) {  // the capacity was too small: the masks are half-edited, so the stencil runs again
                    appended = false;
                    ++nretaken;
                    SKB_TRY(skb_raw_neighbors(bits, ndim, prm->shape, lin, N, prm->image_stack, nb, s));
                } else if (jcap > 0 && n_junction_nodes > jcap) {
                    // more junction voxels than the dense arrays hold: the launch did nothing (masks and edges are intact);
                    // once more with arrays indexed by node id
                    ++nretaken;
                    mpar = work.take<int32_t>(N);
                    bestw = work.take<uint64_t>(2 * N);
                    SKB_NEED(work.ok);
                    SKB_TRY(skb_mst_keys_run(ea, eb, ek, prm->wtab, nej, nullptr, mpar, (unsigned long long*)bestw, N, era, erb,
                                             tree, flag, keep, 0, (int32_t)N, s, nullptr));
                }
            }
#endif
            if (!appended) {
                work.off = work_mark;
                work.ok = true;
            }
        }
        if (!appended) {
            SKB_TRY(skb_junction_edge_count(bits, pre256, ndim, prm->shape, lin, nb, N, 0, N, jcount, s));
            SKB_TRY(skb_scan_u32_impl(jcount, jcount, N, scratch, s));
            SKB_MARK(SKB_M_JCOUNT);
            SKB_POST(jcount + N);
            if (cap > 0) {  // edge emission sized by the hint runs while the count travels
                take_junction(cap, 0);
                if (work.ok) {
                    SKB_TRY(skb_junction_edge_emit_cap(bits, pre256, ndim, prm->shape, lin, nb, jcount, N, prm->wtab, values,
                                                       height, prm->dtype, ea, eb, ek, ew, cap, s));
                    emitted = true;
                    ++noverlap;
                }
            }
            SKB_WAIT(nej);
            if (nej && (!emitted || nej > cap)) {
                if (emitted) ++nretaken;
                work.off = work_mark;
                work.ok = true;
                take_junction(nej, 0);
                SKB_NEED(work.ok);
                SKB_TRY(skb_junction_edge_emit(bits, pre256, ndim, prm->shape, lin, nb, jcount, N, prm->wtab, values, height,
                                               prm->dtype, ea, eb, ek, ew, s));
            } else if (!nej) {
                work.off = work_mark;
                work.ok = true;
            }
            if (nej)  // the dropped junction-junction edges are cleared from the raw masks in place: keep == nb
                SKB_TRY(skb_mst_junctions_apply(ea, eb, ek, ew, prm->wtab, height, nej, N, mpar, bestw, era, erb, tree, flag,
                                                keep, 0, N, s));
        }
    }
    res->counts[SKB_C_JUNCTION_EDGES] = nej;
    res->counts[SKB_C_JUNCTION_NODES] = n_junction_nodes;
    SKB_MARK(SKB_M_MST);
    // ---- CSR + node records ------------------------------------------------------------------------------
    int32_t* deg = out.take<int32_t>(N);
    int32_t* indptr = out.take<int32_t>(N + 1);
    SKB_NEED(out.ok);
    SKB_TRY(skb_degrees_indptr(keep, N, deg, indptr, scratch, s));
    SKB_MARK(SKB_M_DEGREES);
    int64_t E;
    const bool paths_wanted = !prm->stop_after_graph;
    int32_t* indices = nullptr;
    double* data = nullptr;
    skb_i4* rec = nullptr;
    uint32_t* startoff = nullptr;
    int32_t *cc_par = nullptr, *cc_min = nullptr;
    uint32_t* is_min = nullptr;
    auto take_csr = [&](int64_t cap) {
        indices = out.take<int32_t>(cap);
        data = out.take<double>(cap);
        rec = paths_wanted ? work.take<skb_i4>(2 * N) : nullptr;
        startoff = paths_wanted ? work.take<uint32_t>(N + 1) : nullptr;
        cc_par = paths_wanted ? work.take<int32_t>(N) : nullptr;
        cc_min = paths_wanted ? work.take<int32_t>(N) : nullptr;
        is_min = paths_wanted ? out.take<uint32_t>(N + 1) : nullptr;
    };
    {
        const int64_t cap = prm->hint[SKB_H_EDGES] > 0 ? prm->hint[SKB_H_EDGES] : 0;
        const int64_t work_mark = work.off, out_mark = out.off;
        bool emitted = false;
        SKB_POST(indptr + N);
        if (cap > 0) {  // CSR emission sized by the hint runs while the count travels
            take_csr(cap);
            if (work.ok && out.ok) {
                SKB_TRY(skb_csr_emit_cap(bits, pre256, ndim, prm->shape, lin, keep, indptr, N, prm->wtab, values, height,
                                         prm->dtype, indices, data, rec, nullptr, 0, startoff, cc_par, cc_min, is_min,
                                         scratch, cap, s, N));
                emitted = true;
                ++noverlap;
            }
        }
        SKB_WAIT(E);
        if (!emitted || E > cap) {
            if (emitted) ++nretaken;
            work.off = work_mark;
            out.off = out_mark;
            work.ok = out.ok = true;
            take_csr(E);
            SKB_NEED(work.ok && out.ok);
            SKB_TRY(skb_csr_emit_cap(bits, pre256, ndim, prm->shape, lin, keep, indptr, N, prm->wtab, values, height,
                                     prm->dtype, indices, data, rec, nullptr, 0, startoff, cc_par, cc_min, is_min, scratch,
                                     0x7fffffffLL, s, N));
        }
    }
    res->counts[SKB_C_EDGES] = E;
    SKB_MARK(SKB_M_CSR);
    res->off[SKB_O_DEGREES] = out_off(deg);
    res->off[SKB_O_INDPTR] = out_off(indptr);
    res->off[SKB_O_INDICES] = out_off(indices);
    res->off[SKB_O_DATA] = out_off(data);
    res->off[SKB_O_COMP_RANK] = out_off(is_min);
    if (!paths_wanted || N == 0 || E == 0) {
        SKB_MARKS_DONE();
        SKB_RUN_DONE();
        return 0;
    }
    // ---- paths: phase 0 (between terminals), phase 1 (junction-free cycles) ------------------------------------
    struct Phase {
        int64_t S;
        int32_t *su, *se;
        skb_i4 *ranked, *einfo;
        uint32_t* sel;
    };
    auto run_phase = [&](int ph, const uint32_t* so, Phase& P, int64_t* n_heads) -> int {
        skb_i4* buf1 = nullptr;
        auto take_starts = [&](int64_t cap) {
            P.su = work.take<int32_t>(cap);
            P.se = work.take<int32_t>(cap);
            P.ranked = work.take<skb_i4>(cap);
            buf1 = work.take<skb_i4>(cap);
            P.sel = work.take<uint32_t>(cap + 1);
            P.einfo = work.take<skb_i4>(cap);
        };
        const int64_t cap = (ph == 0 && prm->hint[SKB_H_STARTS] > 0) ? prm->hint[SKB_H_STARTS] : 0;
        const int64_t work_mark = work.off;
        bool filled = false;
        SKB_POST(so + N);
        if (cap > 0) {  // the start lists sized by the hint are filled while the count travels
            take_starts(cap);
            if (work.ok) {
                SKB_TRY(skb_path_start_fill_cap(indptr, so, N, P.su, P.se, cap, s));
                filled = true;
                ++noverlap;
            }
        }
        SKB_WAIT(P.S);
        *n_heads = 0;
        if (P.S == 0) {
            work.off = work_mark;
            work.ok = true;
            return 0;
        }
        if (!filled || P.S > cap) {
            if (filled) ++nretaken;
            work.off = work_mark;
            work.ok = true;
            take_starts(P.S);
            if (!work.ok) return SKB_RUN_NEED_MORE;
            filled = false;
        }
        // indptr NULL: su / se are filled already
        SKB_TRY(skb_path_walk(filled ? nullptr : indptr, indices, data, values, nullptr, rec, so, nullptr, N, P.S, P.su,
                              P.se, P.ranked, nullptr, nullptr, s));
        if (ph == 0) SKB_MARK(SKB_M_WALK);
        int r = skb_path_rank(P.ranked, buf1, nullptr, nullptr, P.S, 0, P.S, flag, s);
        if (r < 0) return r;
        if (r & 1) P.ranked = buf1;
        if (ph == 0) SKB_MARK(SKB_M_RANK);
        SKB_TRY(skb_path_select(P.su, rec, P.ranked, P.S, ph, 0, P.sel, scratch, s));
        if (ph == 0) SKB_MARK(SKB_M_SELECT);
        if (ph == 0 && prm->hint[SKB_H_REGULAR] > 0) {
            SKB_POST(P.sel + P.S);  // the caller launches the heads pass sized by the hint, then waits
            *n_heads = -1;
            return 0;
        }
        SKB_FETCH(*n_heads, P.sel + P.S);
        return 0;
    };
    Phase A{}, B{};
    int64_t n_regular = 0, n_cycles = 0;
    {
        int r = run_phase(0, startoff, A, &n_regular);
        if (r == SKB_RUN_NEED_MORE) SKB_NEED(false);
        if (r) return r;
    }
    res->counts[SKB_C_STARTS] = A.S;
    int32_t *start_node = nullptr, *pmin = nullptr, *pindptr = nullptr;
    uint32_t* plen = nullptr;
    auto take_paths = [&](int64_t regular_cap) {
        const int64_t P_cap = regular_cap + E / 6 + 1;  // a junction-free cycle has at least 3 edges
        start_node = work.take<int32_t>(P_cap);
        plen = work.take<uint32_t>(P_cap + 1);
        pmin = work.take<int32_t>(P_cap);
        pindptr = out.take<int32_t>(P_cap + 1);
    };
    bool heads_done = false;
    if (n_regular < 0) {  // the count of regular paths is on its way: the heads pass sized by the hint runs meanwhile
        const int64_t cap = prm->hint[SKB_H_REGULAR];
        const int64_t work_mark = work.off, out_mark = out.off;
        take_paths(cap);
        if (work.ok && out.ok) {
            SKB_TRY(skb_path_heads_cap(A.su, rec, A.ranked, nullptr, A.sel, A.S, 0, 0, 0, 0, start_node, plen, pmin,
                                       A.einfo, nullptr, nullptr, cap, both_ways, s));
            heads_done = true;
            ++noverlap;
        }
        SKB_WAIT(n_regular);
        if (!heads_done || n_regular > cap) {
            if (heads_done) ++nretaken;
            work.off = work_mark;
            out.off = out_mark;
            work.ok = out.ok = true;
            heads_done = false;
        }
    }
    if (!heads_done) {
        take_paths(n_regular);
        SKB_NEED(work.ok && out.ok);
        if (n_regular)
            SKB_TRY(skb_path_heads_cap(A.su, rec, A.ranked, nullptr, A.sel, A.S, 0, 0, 0, 0, start_node, plen, pmin,
                                       A.einfo, nullptr, nullptr, 0x7fffffffLL, both_ways, s));
    }
    // Entries: every undirected edge is on exactly one path, so L = E/2 + P; with at most E/6 cycles the arrays can
    // be sized, and the second walk launched, before the number of entries on regular paths (which tells how
    // many chain voxels no regular path covers) has arrived.
    const int64_t L_cap = E / 2 + n_regular + E / 6 + 1;
    int32_t* pidx = out.take<int32_t>(L_cap);
    double* pdata = out.take<double>(L_cap);
    double* pw = out.take<double>(L_cap);
    uint8_t* covered = work.take<uint8_t>(N);
    SKB_NEED(work.ok && out.ok);
    int64_t L_reg = 0;
    if (n_regular) {
        SKB_TRY(skb_scan_u32_impl(plen, (uint32_t*)pindptr, n_regular, scratch, s));
        SKB_MARK(SKB_M_HEADS);
        SKB_POST(pindptr + n_regular);
    }
    SKB_TRY(skb_fill_bytes(covered, 0, (size_t)N, s));
    // boolean image: paths.data is all ones (csr.py:207-215) -- one coalesced fill instead of a scattered 8-byte
    // store per path entry in the walks below
    double* pdata_w = values ? pdata : nullptr;
    if (!values) SKB_TRY(skb_foreach(L_cap, s, SkbFillF64Item{pdata, 1.0}));
    if (n_regular) {
        SKB_TRY(skb_path_emit_ways(indices, data, rec, A.su, A.se, A.ranked, A.einfo, pindptr, values, covered, 0, A.S,
                                   pidx, pdata_w, pw, both_ways, s));
        ++noverlap;
        SKB_WAIT(L_reg);
    }
    int64_t n_uncovered = E / 2 - (L_reg - n_regular);
    if (n_uncovered < 0) n_uncovered = 0;
    SKB_MARK(SKB_M_EMIT);
    bool cycles_done = false;
    if (n_uncovered > 0) {
        // few uncovered chain voxels (the usual case: short rings): one thread block does the whole phase
        int32_t* list = work.take<int32_t>(n_uncovered);
        SKB_NEED(work.ok);
        if (!n_regular) SKB_TRY(skb_fill_bytes(pindptr, 0, 4, s));
        int r = skb_cycles_small(rec, covered, N, n_uncovered, values, n_regular, start_node, plen, pmin, pindptr, pidx,
                                 pdata_w, pw, is_min, list, flag, s);
        if (r < 0) return r;
        if (r == 0) {
            const int32_t* ptrs[2] = {flag, flag + 2};
            int64_t got[2];
            SKB_TRY(skb_fetch_i32s(ptrs, 2, got, s));
            ++nsync;
            if (got[1] != 0) return skb_fail("junction-free cycles: the uncovered chain voxels do not form closed rings");
            n_cycles = got[0];
            cycles_done = true;
        }
    }
    if (n_uncovered > 0 && !cycles_done) {
        int32_t* cpar = work.take<int32_t>(N);
        uint32_t* startoff_b = work.take<uint32_t>(N + 1);
        SKB_NEED(work.ok);
        SKB_TRY(skb_cycle_roots(rec, covered, N, cpar, is_min, s));
        SKB_TRY(skb_path_start_count(rec, covered, nullptr, N, startoff_b, scratch, s));
        int r = run_phase(1, startoff_b, B, &n_cycles);
        if (r == SKB_RUN_NEED_MORE) SKB_NEED(false);
        if (r) return r;
        if (n_cycles) {
            SKB_TRY(skb_path_heads_cap(B.su, rec, B.ranked, nullptr, B.sel, B.S, 1, n_regular, 0, 0, start_node, plen,
                                       pmin, B.einfo, nullptr, nullptr, 0x7fffffffLL, both_ways, s));
            SKB_TRY(skb_scan_u32_impl(plen, (uint32_t*)pindptr, n_regular + n_cycles, scratch, s));
            SKB_TRY(skb_path_emit_ways(indices, data, rec, B.su, B.se, B.ranked, B.einfo, pindptr, values, nullptr, 0,
                                       B.S, pidx, pdata_w, pw, both_ways, s));
        }
    }
    SKB_MARK(SKB_M_CYCLES);
    const int64_t P = n_regular + n_cycles;
    const int64_t L = L_reg + n_uncovered + n_cycles;
    res->counts[SKB_C_PATHS] = P;
    res->counts[SKB_C_REGULAR] = n_regular;
    res->counts[SKB_C_CYCLES] = n_cycles;
    res->counts[SKB_C_ENTRIES] = L;
    res->off[SKB_O_PINDPTR] = out_off(pindptr);
    res->off[SKB_O_PIDX] = out_off(pidx);
    res->off[SKB_O_PDATA] = out_off(pdata);
    res->off[SKB_O_PW] = out_off(pw);
    if (P == 0) {
        SKB_MARKS_DONE();
        SKB_RUN_DONE();
        return 0;
    }
    // ---- skeleton ids, statistics, branch table ----------------------------------------------------------------------
    int32_t* skel = out.take<int32_t>(P);
    double* dist = out.take<double>(P);
    double* mean = out.take<double>(P);
    double* stdev = out.take<double>(P);
    const int td = prm->table_ndim;
    const int dh = td + height;
    int32_t* t32 = out.take<int32_t>(3 * P);
    int64_t* t64 = out.take<int64_t>((1 + 2 * td) * P);
    double* tf = out.take<double>((2 * dh + 1) * P);
    SKB_NEED(out.ok);
    SKB_TRY(skb_skeleton_ids(pindptr, pidx, pmin, N, n_regular, P, cc_par, cc_min, is_min, skel, scratch, s));
    SKB_MARK(SKB_M_SKEL);
    SKB_TRY(skb_path_stats(pindptr, values ? pdata : nullptr, pw, P, dist, mean, stdev, s));
    SKB_MARK(SKB_M_STATS);
    SKB_TRY(skb_branch_table(P, td, height, prm->spacing_is_int, prm->spacing_f, prm->spacing_i, pindptr, pidx, indptr,
                             skel, coords, values, nullptr, nullptr, 0, nullptr, t32, t64, tf, s));
    SKB_MARK(SKB_M_TABLE);
    SKB_MARKS_DONE();
    res->off[SKB_O_SKEL] = out_off(skel);
    res->off[SKB_O_DIST] = out_off(dist);
    res->off[SKB_O_MEAN] = out_off(mean);
    res->off[SKB_O_STDEV] = out_off(stdev);
    res->off[SKB_O_T32] = out_off(t32);
    res->off[SKB_O_T64] = out_off(t64);
    res->off[SKB_O_TF] = out_off(tf);
    SKB_RUN_DONE();
    return 0;
}
