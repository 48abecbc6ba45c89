// skan_b200 -- bodies of the C-ABI entry points declared in include/skan_b200.h.
//
// Included by skb_cuda.cu (sm_100a build: pointers are device pointers, `stream` is a
// cudaStream_t) and by tests/emul/skb_emul.cpp (host emulation harness for CPU unit tests).
// The including file provides:
//   int  skb_foreach(int64_t n, skb_stream_t s, Functor f)
//   int  skb_fill_bytes(void* p, int byte, size_t nbytes, skb_stream_t s)
//   int  skb_fetch_i32(const int32_t* p, int32_t* host_out, skb_stream_t s)      (synchronises)
//   int  skb_scan_u32_impl (exclusive scan, n+1 outputs; scratch zeroed once by the caller)
//   int  skb_fail(const char* what)                                               (records the message)
// and the three block-cooperative stages (pack, node emit, scans) itself.

#ifndef SKB_TRY
#define SKB_TRY(expr)            \
    do {                         \
        int _e = (expr);         \
        if (_e != 0) return _e;  \
    } while (0)
#endif

#ifndef SKB_RANK_WALK_CAP
#define SKB_RANK_WALK_CAP 48  // hops of the list-following pass before the synchronous pointer-jumping rounds take over
#endif

static SkbGeom skb_make_geom(int ndim, const int64_t* shape) {
    int64_t s[3] = {1, 1, 1};
    for (int i = 0; i < ndim; ++i) s[3 - ndim + i] = shape[i];
    return skb_geom_make(ndim, s[0], s[1], s[2]);
}

static int skb_check_geom(int ndim, const int64_t* shape) {
    if (ndim < 1 || ndim > 3) return skb_fail("ndim must be 1, 2 or 3");
    for (int i = 0; i < ndim; ++i)
        if (shape[i] < 1 || shape[i] > 0x7fffffffLL) return skb_fail("bad extent");
    return 0;
}

// `slab` (host, 13 x i64, or NULL on a single GPU) = own0, own1, cnt0, cnt1, ga0, ga1, gb0, gb1,
// id_shift, start_shift, start_lo, start_hi, lin_offset -- see struct SkbSlab.
static SkbSlab skb_make_slab(const int64_t* v, int64_t n) {
    SkbSlab sl;
    if (!v) {
        sl.own0 = 0; sl.own1 = (int32_t)n; sl.cnt0 = 0; sl.cnt1 = (int32_t)n;
        sl.ga0 = sl.ga1 = sl.gb0 = sl.gb1 = 0;
        sl.id_shift = 0; sl.start_shift = 0; sl.start_lo = 0; sl.start_hi = 0x7fffffff; sl.lin_offset = 0;
        return sl;
    }
    sl.own0 = (int32_t)v[0]; sl.own1 = (int32_t)v[1]; sl.cnt0 = (int32_t)v[2]; sl.cnt1 = (int32_t)v[3];
    sl.ga0 = (int32_t)v[4]; sl.ga1 = (int32_t)v[5]; sl.gb0 = (int32_t)v[6]; sl.gb1 = (int32_t)v[7];
    sl.id_shift = (int32_t)v[8]; sl.start_shift = (int32_t)v[9];
    sl.start_lo = (int32_t)v[10]; sl.start_hi = (int32_t)v[11]; sl.lin_offset = v[12];
    return sl;
}

// one logical thread per item i < *n_ptr (a count in device memory), launched over `cap` items
template <typename F>
static int skb_foreach_n(int64_t cap, const int32_t* n_ptr, skb_stream_t s, F f) {
    return skb_foreach(cap, s, SkbBounded<F>{f, n_ptr});
}

extern "C" {

int skb_abi_version(void) { return 1; }

// tile prefix (ntiles + 1 entries) of a mask whose rank structure exists already (after skb_sweep_nodes)
int skb_tile_prefix(const uint32_t* pre256, const int32_t* n_nodes, int64_t ntiles, uint32_t* tile_prefix,
                    void* stream) {
    SkbTilePrefixItem it{pre256, n_nodes, ntiles, tile_prefix};
    return skb_foreach(ntiles + 1, (skb_stream_t)stream, it);
}

// raw 3^d neighbourhood masks of the N foreground voxels            (replaces csr.py:122-144)
int skb_raw_neighbors(const uint64_t* bits, int ndim, const int64_t* shape, const int64_t* lin, int64_t n,
                      int planes_independent, uint32_t* nb, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    SkbRawNbItem it{skb_make_geom(ndim, shape), bits, lin, nb, planes_independent};
    return skb_foreach(n, (skb_stream_t)stream, it);
}

// restrict raw neighbourhood masks to the directions in `allowed` (bit k27), for pixel_graph(connectivity < ndim)
int skb_mask_neighbors(uint32_t* nb, int64_t n, uint32_t allowed, void* stream) {
    return skb_foreach(n, (skb_stream_t)stream, SkbAndMaskItem{nb, allowed});
}

// make_degree_image (csr.py:1173-1208): image i64[V] (zeroed by the caller) receives the raw degree at every node
int skb_degree_image(const uint32_t* nb, const int64_t* lin, int64_t n, int64_t* image, void* stream) {
    return skb_foreach(n, (skb_stream_t)stream, SkbDegreeImageItem{nb, lin, image});
}

// junction-junction edges: per-node forward counts, then emission       (csr.py:1004-1016)
int skb_junction_edge_count(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape,
                            const int64_t* lin, const uint32_t* nb, int64_t n, int64_t a0, int64_t a1,
                            uint32_t* count, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    SkbJuncCountItem it{skb_make_geom(ndim, shape), bits, pre256, lin, nb, (int32_t)a0, (int32_t)a1, count};
    return skb_foreach(n, (skb_stream_t)stream, it);
}

// (the _cap variants take the capacity of the arrays they fill: the native runner sizes them by a hint while
// the exact count is still on its way to the host; entries beyond the capacity are counted, not written)
static int skb_junction_edge_emit_cap(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape,
                                      const int64_t* lin, const uint32_t* nb, const uint32_t* offset, int64_t n,
                                      const double* wtab, const double* values, int height, int dtype, int32_t* ea,
                                      int32_t* eb, uint8_t* ek, uint64_t* ew, int64_t cap, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    if (height && !values) return skb_fail("height mode needs node values");
    SkbWeight wf{wtab, values, height, dtype};
    SkbJuncEmitItem it{skb_make_geom(ndim, shape), bits, pre256, lin, nb, offset, wf, ea, eb, ek, ew,
                       cap > 0xffffffffLL ? 0xffffffffu : (uint32_t)cap};
    return skb_foreach(n, (skb_stream_t)stream, it);
}
int skb_junction_edge_emit(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape,
                           const int64_t* lin, const uint32_t* nb, const uint32_t* offset, int64_t n,
                           const double* wtab, const double* values, int height, int dtype, int32_t* ea,
                           int32_t* eb, uint8_t* ek, uint64_t* ew, void* stream) {
    return skb_junction_edge_emit_cap(bits, pre256, ndim, shape, lin, nb, offset, n, wtab, values, height, dtype, ea,
                                      eb, ek, ew, 0xffffffffLL, stream);
}

// junction MST (csr.py:980-1020).  Edge endpoints are node ids < n (global ids when the edges of
// several Z-slabs were gathered).  Scratch: par i32[n], bestw u64[n], beste u32[n] (only the
// entries of edge endpoints are touched), era/erb i32[nej], flag i32[1].
// Out: tree u8[nej] (1 = edge stays).  Returns rounds >= 0.
int skb_mst_junctions(const int32_t* ea, const int32_t* eb, const uint64_t* ew, int64_t nej, int64_t n,
                      int32_t* par, uint64_t* bestw, uint32_t* beste, int32_t* era, int32_t* erb, uint8_t* tree,
                      int32_t* flag, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (nej <= 0) return 0;
    (void)n;
#ifdef SKB_HAVE_FUSED_LOOPS
    return skb_mst_run(ea, eb, ew, nej, par, bestw, beste, era, erb, tree, flag, s);
#else
    SKB_TRY(skb_foreach(nej, s, SkbMstInitItem{ea, eb, par, bestw, beste}));
    SKB_TRY(skb_foreach(nej, s, SkbMstDeadItem{ea, era}));
    SKB_TRY(skb_fill_bytes(tree, 0, (size_t)nej, s));
    int rounds = 0;
    for (;; ++rounds) {
        if (rounds > 64) return skb_fail("junction MST did not converge");
        SKB_TRY(skb_fill_bytes(flag, 0, 4, s));
        SKB_TRY(skb_foreach(nej, s, SkbMstMinWItem{ea, eb, ew, par, era, erb, bestw, flag}));
        int32_t any = 0;
        SKB_TRY(skb_fetch_i32(flag, &any, s));
        if (!any) break;
        SKB_TRY(skb_foreach(nej, s, SkbMstMinEItem{ew, era, erb, bestw, beste}));
        SKB_TRY(skb_foreach(nej, s, SkbMstHookItem{era, erb, beste, par, tree}));
        SKB_TRY(skb_foreach(nej, s, SkbMstResetItem{era, erb, bestw, beste}));
    }
    return rounds;
#endif
}

// Junction MST + removal of the dropped edges in one call, for the runners: keep u32[n_local] holds the raw masks of the
// nodes with ids id_shift .. id_shift + n_local - 1 on entry and the final masks on return.  Table weights (height == 0):
// packed-key rounds, one cooperative launch, see skb_mst_keys_coop_kernel; scratch best u64[2 * n_ids].  Height mode or
// the emulation harness: skb_mst_junctions + skb_mst_apply with the same scratch (bestw = best, beste = best + n_ids).
int skb_mst_apply(const int32_t* ea, const int32_t* eb, const uint8_t* ek, const uint8_t* tree, int64_t nej,
                  int64_t id_shift, int64_t n_local, uint32_t* keep, void* stream);
static int skb_mst_junctions_apply(const int32_t* ea, const int32_t* eb, const uint8_t* ek, const uint64_t* ew,
                                   const double* wtab, int height, int64_t nej, int64_t n_ids, int32_t* par,
                                   uint64_t* best, int32_t* era, int32_t* erb, uint8_t* tree, int32_t* flag,
                                   uint32_t* keep, int64_t id_shift, int64_t n_local, void* stream) {
    if (nej <= 0) return 0;
#ifdef SKB_HAVE_FUSED_LOOPS
    if (!height && g_skb_mst_keys)
        return skb_mst_keys_run(ea, eb, ek, wtab, nej, nullptr, par, (unsigned long long*)best, n_ids, era, erb, tree, flag,
                                keep, (int32_t)id_shift, (int32_t)n_local, (skb_stream_t)stream);
#endif
    (void)wtab;
    int r = skb_mst_junctions(ea, eb, ew, nej, n_ids, par, best, (uint32_t*)(best + n_ids), era, erb, tree, flag, stream);
    if (r < 0) return r;
    return skb_mst_apply(ea, eb, ek, tree, nej, id_shift, n_local, keep, stream);
}

// keep[] starts as a copy of nb[]; clear both directions of every non-tree junction edge.
// keep has n_local entries for the nodes with (global) ids id_shift .. id_shift + n_local - 1.
int skb_mst_apply(const int32_t* ea, const int32_t* eb, const uint8_t* ek, const uint8_t* tree, int64_t nej,
                  int64_t id_shift, int64_t n_local, uint32_t* keep, void* stream) {
    return skb_foreach(nej, (skb_stream_t)stream,
                       SkbMstApplyItem{ea, eb, ek, tree, (int32_t)id_shift, (int32_t)n_local, keep});
}

// degrees (post-MST, csr.py:660) and graph.indptr: deg i32[n], indptr i32[n+1]
int skb_degrees_indptr(const uint32_t* keep, int64_t n, int32_t* deg, int32_t* indptr, void* scratch,
                       void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n, s, SkbPopcItem{keep, (uint32_t*)deg}));
    return skb_scan_u32_impl((const uint32_t*)deg, (uint32_t*)indptr, n, scratch, s);
}

// graph.indices / graph.data in canonical CSR order (csr.py:153-157).  With rec != NULL the same
// pass also writes what skb_path_records would: node records rec (32 B per node), scanned phase-0
// start offsets startoff u32[n+1] and the initial state of the skeleton-id stage.
static int skb_csr_emit_cap(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape,
                            const int64_t* lin, const uint32_t* keep, const int32_t* indptr, int64_t n,
                            const double* wtab, const double* values, int height, int dtype, int32_t* indices,
                            double* data, skb_i4* rec, const int64_t* slab, int lin_splitters, uint32_t* startoff,
                            int32_t* par, int32_t* cmin, uint32_t* is_min, void* scratch, int64_t edge_cap,
                            void* stream, int64_t split_population = 0) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_check_geom(ndim, shape));
    if (height && !values) return skb_fail("height mode needs node values");
    SkbWeight wf{wtab, values, height, dtype};
    // split_population: the node count the splitter density is chosen by (the runners pass it; every rank of a Z-slab run
    // passes the same number); 0 = the fixed default
    SkbRecOut out{lin_splitters ? lin : nullptr, skb_make_slab(slab, n), rec, startoff, par, cmin, is_min,
                  skb_option_split_shift(split_population)};
    SkbCsrEmitItem it{skb_make_geom(ndim, shape), bits, pre256, lin, keep, indptr, wf, indices, data, rec ? 1 : 0, out,
                      edge_cap > 0x7fffffffLL ? 0x7fffffff : (int32_t)edge_cap};
    SKB_TRY(skb_foreach(n, s, it));
    if (rec) return skb_scan_u32_impl(startoff, startoff, n, scratch, s);
    return 0;
}
int skb_csr_emit(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape, const int64_t* lin,
                 const uint32_t* keep, const int32_t* indptr, int64_t n, const double* wtab, const double* values,
                 int height, int dtype, int32_t* indices, double* data, skb_i4* rec, const int64_t* slab,
                 int lin_splitters, uint32_t* startoff, int32_t* par, int32_t* cmin, uint32_t* is_min, void* scratch,
                 void* stream) {
    return skb_csr_emit_cap(bits, pre256, ndim, shape, lin, keep, indptr, n, wtab, values, height, dtype, indices, data,
                            rec, slab, lin_splitters, startoff, par, cmin, is_min, scratch, 0x7fffffffLL, stream);
}

// ---- path tracing (csr.py:347-446): sampled list ranking, see skb_items.cuh ------------------
// node records from an existing CSR: rec (32 B per node: (n0, n1, indptr, degree|flags), (w0, w1)),
// the phase-0 start offsets (startoff u32[n+1], scanned) and the initial state of the skeleton-id
// stage (par, cmin i32[n], may be NULL; is_min u32[n+1]).  lin (may be NULL) makes the splitter
// sample a function of the voxel position instead of the node id.
int skb_path_records(const int32_t* indptr, const int32_t* indices, const double* gdata, const int64_t* lin,
                     const int64_t* slab, int64_t n, skb_i4* rec, uint32_t* startoff, int32_t* par, int32_t* cmin,
                     uint32_t* is_min, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SkbRecOut out{lin, skb_make_slab(slab, n), rec, startoff, par, cmin, is_min, skb_option_split_shift()};
    SKB_TRY(skb_foreach(n, s, SkbPathRecItem{indptr, indices, gdata, out}));
    return skb_scan_u32_impl(startoff, startoff, n, scratch, s);
}

// walk starts of phase 1 (cycle roots + uncovered splitters), scanned in place: startoff u32[n+1]
int skb_path_start_count(const skb_i4* rec, const uint8_t* covered, const int64_t* slab, int64_t n,
                         uint32_t* startoff, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (!covered) return skb_fail("phase 1 needs the covered marks");
    // covered nodes start nothing: zero the counts in one fill, visit only the uncovered nodes
    SKB_TRY(skb_fill_bytes(startoff, 0, (size_t)n * 4, s));
    SKB_TRY(skb_foreach_unmarked(n, covered, s, SkbStartCountItem{rec, covered, skb_make_slab(slab, n), startoff}));
    return skb_scan_u32_impl(startoff, startoff, n, scratch, s);
}

// the first half of skb_path_walk on its own, with a capacity for su / se
static int skb_path_start_fill_cap(const int32_t* indptr, const uint32_t* startoff, int64_t n, int32_t* su, int32_t* se,
                                   int64_t cap, void* stream) {
    return skb_foreach(n, (skb_stream_t)stream,
                       SkbStartFillItem{indptr, startoff, 0, su, se, cap > 0xffffffffLL ? 0xffffffffu : (uint32_t)cap});
}

// su/se i32[S] (node and half-edge of every start leaving an owned node) and the first local walk:
// ranked i4[S]; aux (48 B per start; NULL on a single GPU) carries sums and the end node; stamp
// i32[n] (may be NULL, pre-set to -1) remembers a start per chain node passed.
int skb_path_walk(const int32_t* indptr, const int32_t* indices, const double* gdata, const double* values,
                  const int64_t* lin, const skb_i4* rec, const uint32_t* startoff, const int64_t* slab, int64_t n,
                  int64_t n_starts, int32_t* su, int32_t* se, skb_i4* ranked, void* aux, int32_t* stamp,
                  void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SkbSlab sl = skb_make_slab(slab, n);
    if (indptr)  // NULL: su / se were filled by skb_path_start_fill_cap already
        SKB_TRY(skb_foreach((int64_t)sl.own1 - sl.own0, s, SkbStartFillItem{indptr, startoff, sl.own0, su, se}));
    if (aux)
        return skb_foreach_steps(n_starts, s,
                                 SkbWalkItemT<true>{indices, gdata, values, lin, rec, startoff, su, se, sl, ranked,
                                                    (SkbAux*)aux, stamp});
    return skb_foreach_steps(
        n_starts, s, SkbWalkItemT<false>{indices, gdata, values, lin, rec, startoff, su, se, sl, ranked, nullptr, stamp});
}

// pointer jumping over the starts of this rank (synchronous rounds, ping-pong between buf0/aux0
// and buf1/aux1; aux may be NULL).  Starts with global ids [lo, hi) are this rank's.  Stops after
// the first round in which no start stops pointing into [lo, hi): every list that ends at a
// terminal or leaves the slab is then resolved, what is left cycles.
// Returns 2*rounds + (index of the result buffer).
int skb_path_rank(skb_i4* buf0, skb_i4* buf1, void* aux0, void* aux1, int64_t n_starts, int64_t lo, int64_t hi,
                  int32_t* flag, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (n_starts <= 0) return 0;
    // one pass in which every start follows its (short) list: buf0 -> buf1; flag[3] != 0 afterwards means
    // that some list is longer than the hop limit and the synchronous rounds have to finish the job
    SKB_TRY(skb_fill_bytes(flag, 0, 16, s));
    SKB_TRY(skb_foreach(n_starts, s,
                        SkbRankWalkItem{buf0, buf1, (const SkbAux*)aux0, (SkbAux*)aux1, (int32_t)lo, (int32_t)hi,
                                        SKB_RANK_WALK_CAP, flag + 3}));
#ifdef SKB_HAVE_FUSED_LOOPS
    // the cooperative kernel returns at once when flag[3] is 0, else it runs the rounds with buf1 as its buffer 0
    SKB_TRY(skb_rank_run(buf1, buf0, (SkbAux*)aux1, (SkbAux*)aux0, n_starts, (int32_t)lo, (int32_t)hi, flag, s));
    return 1;
#else
    {
        int32_t unfinished = 0;
        SKB_TRY(skb_fetch_i32(flag + 3, &unfinished, s));
        if (!unfinished) return 1;
    }
    skb_i4* b[2] = {buf0, buf1};
    SkbAux* a[2] = {(SkbAux*)aux0, (SkbAux*)aux1};
    int cur = 1, rounds = 0;
    const int blind = 3;  // rounds run before the first host check (lists of up to 8 starts)
    for (;; ++rounds) {
        if (rounds > 40) return skb_fail("list ranking did not converge");
        if (rounds >= blind || rounds == 0) SKB_TRY(skb_fill_bytes(flag, 0, 4, s));
        SKB_TRY(skb_foreach(n_starts, s,
                            SkbJumpItem{b[cur], b[cur ^ 1], a[cur], a[cur ^ 1], (int32_t)lo, (int32_t)hi, flag}));
        cur ^= 1;
        if (rounds + 1 >= blind) {
            int32_t any = 0;
            SKB_TRY(skb_fetch_i32(flag, &any, s));
            if (!any) {
                ++rounds;
                break;
            }
        }
    }
    return 2 * rounds + cur;
#endif
}

// heads of emitted paths among the starts: sel u32[S+1] = exclusive scan of the head flags
int skb_path_select(const int32_t* su, const skb_i4* rec, const skb_i4* ranked, int64_t n_starts, int phase,
                    int64_t start_lo, uint32_t* sel, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n_starts, s, SkbSelectItem{su, rec, ranked, phase, (int32_t)start_lo, sel}));
    return skb_scan_u32_impl(sel, sel, n_starts, scratch, s);
}

// per path (ids pid_base + ...): start_node i32 (global id), plen u32 (entries), pmin i32.
// Single GPU: einfo i4[S] is filled (and cleared first).  Slabs: einfo NULL; head_end i32 and
// head_aux (48 B each) receive the reversed last start and the sums / end node of every path.
static int skb_path_heads_cap(const int32_t* su, const skb_i4* rec, const skb_i4* ranked, const void* aux,
                              const uint32_t* sel, int64_t n_starts, int phase, int64_t pid_base, int64_t start_lo,
                              int64_t id_shift, int32_t* start_node, uint32_t* plen, int32_t* pmin, skb_i4* einfo,
                              int32_t* head_end, void* head_aux, int64_t pid_cap, int both_ways, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (einfo) SKB_TRY(skb_fill_bytes(einfo, 0xff, (size_t)n_starts * 16, s));
    return skb_foreach(n_starts, s,
                       SkbHeadsItem{su, rec, ranked, (const SkbAux*)aux, sel, phase, (int32_t)pid_base,
                                    (int32_t)start_lo, (int32_t)id_shift, start_node, plen, pmin, einfo, head_end,
                                    (SkbAux*)head_aux, pid_cap > 0x7fffffffLL ? 0x7fffffff : (int32_t)pid_cap,
                                    both_ways ? 1 : 0});
}
int skb_path_heads(const int32_t* su, const skb_i4* rec, const skb_i4* ranked, const void* aux,
                   const uint32_t* sel, int64_t n_starts, int phase, int64_t pid_base, int64_t start_lo,
                   int64_t id_shift, int32_t* start_node, uint32_t* plen, int32_t* pmin, skb_i4* einfo,
                   int32_t* head_end, void* head_aux, void* stream) {
    return skb_path_heads_cap(su, rec, ranked, aux, sel, n_starts, phase, pid_base, start_lo, id_shift, start_node, plen,
                              pmin, einfo, head_end, head_aux, 0x7fffffffLL, 0, stream);
}

// second local walk: paths.indices i32[L], paths.data f64[L] (csr.py:396,402; values NULL -> 1.0;
// pdata NULL -> not written: the caller fills the ones of a boolean image in one coalesced pass),
// pw f64[L] = weight of the edge arriving at each entry; chain nodes written are marked in
// covered u8[n] when it is given.  einfo is indexed by the id of a list's reversed last start;
// pindptr NULL means the entry offset of a path is einfo.z (slabs).
// both_ways: the chain between two stops is written from both ends (needs einfo from skb_path_heads_cap with
// both_ways and ranked from skb_path_walk + skb_path_rank of this library: the local hop count rides in ranked.w)
static int skb_path_emit_ways(const int32_t* indices, const double* gdata, const skb_i4* rec, const int32_t* su,
                              const int32_t* se, const skb_i4* ranked, const skb_i4* einfo, const int32_t* pindptr,
                              const double* values, uint8_t* covered, int64_t id_shift, int64_t n_starts, int32_t* pidx,
                              double* pdata, double* pw, int both_ways, void* stream) {
    return skb_foreach_steps(n_starts, (skb_stream_t)stream,
                             SkbEmitItem{indices, gdata, rec, su, se, ranked, einfo, pindptr, values, covered,
                                         (int32_t)id_shift, pidx, pdata, pw, both_ways ? 1 : 0});
}
int skb_path_emit(const int32_t* indices, const double* gdata, const skb_i4* rec, const int32_t* su,
                  const int32_t* se, const skb_i4* ranked, const skb_i4* einfo, const int32_t* pindptr,
                  const double* values, uint8_t* covered, int64_t id_shift, int64_t n_starts, int32_t* pidx,
                  double* pdata, double* pw, void* stream) {
    return skb_path_emit_ways(indices, gdata, rec, su, se, ranked, einfo, pindptr, values, covered, id_shift, n_starts,
                              pidx, pdata, pw, 0, stream);
}

// junction-free cycles (csr.py:369-386): the minimum node of every component of uncovered
// degree-2 nodes is flagged in rec as a cycle root and in is_min.  Scratch: par i32[n].
int skb_cycle_roots(skb_i4* rec, const uint8_t* covered, int64_t n, int32_t* par, uint32_t* is_min, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    // only uncovered nodes act: covered ones are skipped 16 marks at a time.  Every parent is initialised, though:
    // in a Z-slab an uncovered halo node may have covered neighbours
    SKB_TRY(skb_foreach(n, s, SkbIotaItem{par}));
    SKB_TRY(skb_foreach_unmarked(n, covered, s, SkbCycleHookItem{rec, covered, par}));
    return skb_foreach_unmarked(n, covered, s, SkbCycleRootItem{rec, covered, par, is_min});
}

// skeleton ids (csr.py:909): connected components of the contracted graph, labelled by the rank
// of each component's minimum node id.  par, cmin i32[n] and is_min u32[n+1] come initialised from
// skb_path_records / skb_cycle_roots; is_min is scanned in place into the rank table.
int skb_skeleton_ids(const int32_t* pindptr, const int32_t* pidx, const int32_t* pmin, int64_t n,
                     int64_t n_regular, int64_t n_paths, int32_t* par, int32_t* cmin, uint32_t* is_min,
                     int32_t* skel_id, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n_regular, s, SkbCcPathUnionItem{pindptr, pidx, par}));
    SKB_TRY(skb_foreach(n_regular, s, SkbCcPathMinItem{pindptr, pidx, pmin, par, cmin}));
    SKB_TRY(skb_foreach(n_regular, s, SkbCcPathMarkItem{pindptr, pidx, par, cmin, is_min}));
    SKB_TRY(skb_scan_u32_impl(is_min, is_min, n, scratch, s));
    return skb_foreach(n_paths, s, SkbCcPathIdItem{pindptr, pidx, (int32_t)n_regular, par, cmin, is_min, skel_id});
}

// branch_distance / mean / stdev per path (csr.py:962-977, 792-816); null outputs are skipped
// pdata NULL: the image is boolean and every entry of paths.data is 1.0 (mean 1, stdev 0 without reading them)
int skb_path_stats(const int32_t* pindptr, const double* pdata, const double* pw, int64_t n_paths, double* dist,
                   double* mean, double* stdev, void* stream) {
#ifdef SKB_HAVE_FUSED_LOOPS
    SKB_TRY(skb_foreach(n_paths, (skb_stream_t)stream, SkbPathStatsItem{pindptr, pdata, pw, dist, mean, stdev, g_skb_long_path}));
    return skb_long_path_stats(pindptr, pdata, pw, n_paths, dist, mean, stdev, (skb_stream_t)stream);
#else
    return skb_foreach(n_paths, (skb_stream_t)stream, SkbPathStatsItem{pindptr, pdata, pw, dist, mean, stdev});
#endif
}

// the remaining columns of summarize() (csr.py:909-952)
int skb_branch_table(int64_t n_paths, int ndim, int height, int spacing_is_int, const double* spacing_f,
                     const int64_t* spacing_i, const int32_t* pindptr, const int32_t* pidx,
                     const int32_t* gindptr, const int32_t* skel_id, const int64_t* coords, const double* values,
                     const void* head_aux, const int32_t* start_node, int64_t id_shift, const int64_t* global_shape,
                     int32_t* out_i32, int64_t* out_i64, double* out_f64, void* stream) {
    if (ndim < 1 || ndim > 3) return skb_fail("ndim must be 1, 2 or 3");
    if (height && !values) return skb_fail("height mode needs node values");
    if (head_aux && (!start_node || !global_shape)) return skb_fail("bad slab-mode branch table call");
    SkbBranchItem it;
    it.P = n_paths;
    it.d = ndim;
    it.height = height;
    it.spacing_int = spacing_is_int;
    it.pindptr = pindptr;
    it.pidx = pidx;
    it.gindptr = gindptr;
    it.skel_id = skel_id;
    it.head_aux = (const SkbAux*)head_aux;
    it.start_node = start_node;
    it.id_shift = (int32_t)id_shift;
    it.gny = 1;
    it.gnx = 1;
    if (global_shape) {
        it.gnx = global_shape[ndim - 1];
        it.gny = ndim >= 2 ? global_shape[ndim - 2] : 1;
    }
    it.coords = coords;
    it.values = values;
    for (int i = 0; i < 3; ++i) {
        it.sp_f[i] = i < ndim ? spacing_f[i] : 1.0;
        it.sp_i[i] = i < ndim ? spacing_i[i] : 1;
    }
    it.out_i32 = out_i32;
    it.out_i64 = out_i64;
    it.out_f64 = out_f64;
    return skb_foreach(n_paths, (skb_stream_t)stream, it);
}

// ---- Z-slab mode (one volume over several GPUs) ----------------------------------------------
// out[i] = number of foreground voxels before raveled position pos[i] (pos, out on the device)
int skb_rank_query(const uint64_t* bits, const uint32_t* pre256, const int64_t* pos, int64_t n, int32_t* out,
                   void* stream) {
    return skb_foreach(n, (skb_stream_t)stream, SkbRankQueryItem{bits, pre256, pos, out});
}

// covered u8[n] from the stamps of the first walk; n_uncovered u32[64] (summed = owned chain nodes
// that no terminal-started list passes: they lie on junction-free cycles)
int skb_slab_covered(const skb_i4* rec, const uint32_t* startoff, const int32_t* stamp, const skb_i4* ranked,
                     int64_t n, int64_t own0, int64_t own1, uint8_t* covered, uint32_t* n_uncovered,
                     void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_fill_bytes(n_uncovered, 0, 64 * 4, s));
    return skb_foreach(
        n, s, SkbCoveredItem{rec, startoff, stamp, ranked, (int32_t)own0, (int32_t)own1, covered, n_uncovered});
}

static SkbSlabTable skb_make_table(const int64_t* v) {
    // host array: n_ranks, then start_off[n_ranks+1], n_first[n_ranks], last0[n_ranks], t_off[n_ranks]
    SkbSlabTable t;
    int g = (int)v[0];
    t.n_ranks = g;
    const int64_t* p = v + 1;
    for (int k = 0; k <= g; ++k) t.start_off[k] = (int32_t)p[k];
    p += g + 1;
    for (int k = 0; k < g; ++k) t.n_first[k] = (int32_t)p[k];
    p += g;
    for (int k = 0; k < g; ++k) t.last0[k] = (int32_t)p[k];
    p += g;
    for (int k = 0; k < g; ++k) t.t_off[k] = (int32_t)p[k];
    return t;
}

// this rank's block of the inter-slab table: the ranked starts of its two boundary planes
int skb_slab_export(const skb_i4* ranked, const void* aux, int64_t n_first, int64_t last0, int64_t n_block,
                    int64_t n_rows, int64_t lo, int64_t hi, skb_i4* out, void* out_aux, void* stream) {
    return skb_foreach(n_rows, (skb_stream_t)stream,
                       SkbTableExportItem{ranked, (const SkbAux*)aux, (int32_t)n_first, (int32_t)last0, (int32_t)lo,
                                          (int32_t)hi, (int32_t)n_block, out, (SkbAux*)out_aux});
}

// pointer jumping over the gathered table (same on every rank); returns 2*rounds + result buffer
int skb_slab_table_rank(skb_i4* buf0, skb_i4* buf1, void* aux0, void* aux1, int64_t n_table,
                        const int64_t* table_desc, int32_t* flag, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (n_table <= 0) return 0;
    if (table_desc[0] > SKB_MAX_RANKS) return skb_fail("too many ranks");
    SkbSlabTable t = skb_make_table(table_desc);
#ifdef SKB_HAVE_FUSED_LOOPS
    return skb_table_rank_run(buf0, buf1, (SkbAux*)aux0, (SkbAux*)aux1, n_table, t, flag, s);
#else
    skb_i4* b[2] = {buf0, buf1};
    SkbAux* a[2] = {(SkbAux*)aux0, (SkbAux*)aux1};
    int cur = 0, rounds = 0;
    for (;; ++rounds) {
        if (rounds > 40) return skb_fail("inter-slab list ranking did not converge");
        SKB_TRY(skb_fill_bytes(flag, 0, 4, s));
        SKB_TRY(skb_foreach(n_table, s, SkbTableJumpItem{b[cur], b[cur ^ 1], a[cur], a[cur ^ 1], t, flag}));
        cur ^= 1;
        int32_t any = 0;
        SKB_TRY(skb_fetch_i32(flag, &any, s));
        if (!any) {
            ++rounds;
            break;
        }
    }
    return 2 * rounds + cur;
#endif
}

// finish the lists of this rank that left the slab with the ranked table
int skb_slab_apply(skb_i4* ranked, void* aux, const skb_i4* table, const void* table_aux, int64_t n_starts,
                   const int64_t* table_desc, int64_t lo, int64_t hi, void* stream) {
    if (table_desc[0] > SKB_MAX_RANKS) return skb_fail("too many ranks");
    return skb_foreach(n_starts, (skb_stream_t)stream,
                       SkbTableApplyItem{ranked, (SkbAux*)aux, table, (const SkbAux*)table_aux,
                                         skb_make_table(table_desc), (int32_t)lo, (int32_t)hi});
}

// head / link records (i4 each) of the paths headed on this rank, for the all-gather
int skb_slab_head_records(const int32_t* head_end, const uint32_t* plen, const int32_t* pindptr,
                          const int32_t* pmin, const int32_t* start_node, const uint32_t* startoff,
                          const void* head_aux, const int64_t* slab, int64_t n, int64_t n_paths, int64_t n_rows,
                          int64_t pid_off, int64_t entry_off, skb_i4* head, skb_i4* link, void* stream) {
    return skb_foreach(n_rows, (skb_stream_t)stream,
                       SkbHeadRecordItem{head_end, plen, pindptr, pmin, start_node, startoff,
                                         (const SkbAux*)head_aux, skb_make_slab(slab, n), (int32_t)pid_off,
                                         (int32_t)entry_off, (int32_t)n_paths, head, link});
}

// regroup gathered data: desc (host i64) = n_seg, then n_seg x (src address, dst address, bytes);
// every address and size a multiple of 16
int skb_copy_segments(const int64_t* desc, void* stream) {
    int n = (int)desc[0];
    if (n < 0 || n > SKB_MAX_SEGMENTS) return skb_fail("too many segments");
    SkbSegCopyItem it;
    it.n_seg = n;
    int64_t total = 0;
    for (int k = 0; k < n; ++k) {
        it.src[k] = (const uint8_t*)(uintptr_t)desc[1 + 3 * k];
        it.dst[k] = (uint8_t*)(uintptr_t)desc[2 + 3 * k];
        int64_t bytes = desc[3 + 3 * k];
        if ((desc[1 + 3 * k] | desc[2 + 3 * k] | bytes) & 15) return skb_fail("segments must be 16-byte aligned");
        it.first[k] = total;
        total += bytes / 16;
    }
    it.first[n] = total;
    return skb_foreach(total, (skb_stream_t)stream, it);
}

// gathered junction edges (cap rows per rank, a < 0 = padding): add every rank's id shift
int skb_slab_edge_shift(int32_t* ea, int32_t* eb, int64_t n_rows, int64_t cap, const int64_t* shifts, int64_t n_ranks,
                        void* stream) {
    if (n_ranks > SKB_MAX_RANKS) return skb_fail("too many ranks");
    SkbEdgeShiftItem it;
    it.ea = ea;
    it.eb = eb;
    it.cap = (int32_t)cap;
    for (int k = 0; k < SKB_MAX_RANKS; ++k) it.shift[k] = k < n_ranks ? (int32_t)shifts[k] : 0;
    return skb_foreach(n_rows, (skb_stream_t)stream, it);
}

// einfo i4[total starts] (cleared here) from all gathered head records (x < 0 = padding)
int skb_slab_einfo(const skb_i4* head, int64_t n_records, int64_t n_total_starts, skb_i4* einfo, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_fill_bytes(einfo, 0xff, (size_t)n_total_starts * 16, s));
    return skb_foreach(n_records, s, SkbEinfoScatterItem{head, einfo});
}

// skeleton ids across slabs, step 1: components of the contracted graph over start keys; marks
// the component minima that are owned nodes in is_min (local ids) and scans it in place.
// Scratch: par, cmin i32[n_keys].  link = all gathered link records (x < 0 = padding).
int skb_slab_components(const skb_i4* link, int64_t n_records, int64_t n_keys, int32_t* par, int32_t* cmin,
                        int64_t node_lo, int64_t node_hi, int64_t id_shift, uint32_t* is_min, int64_t n,
                        void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n_keys, s, SkbKeyInitItem{par, cmin}));
    SKB_TRY(skb_foreach(n_records, s, SkbKeyUnionItem{link, par}));
    SKB_TRY(skb_foreach(n_records, s, SkbKeyMinItem{link, par, cmin}));
    SKB_TRY(skb_foreach(n_records, s,
                        SkbKeyMarkItem{link, par, cmin, (int32_t)node_lo, (int32_t)node_hi, (int32_t)id_shift, is_min}));
    return skb_scan_u32_impl(is_min, is_min, n, scratch, s);
}

// step 2: this rank's contribution to the skeleton id of every gathered path (summed over ranks
// by the caller): comp_off + rank of the component minimum among the owned component minima
int skb_slab_skeleton_ids(const skb_i4* link, int64_t n_records, int32_t* par, const int32_t* cmin,
                          const uint32_t* rank, int64_t node_lo, int64_t node_hi, int64_t id_shift,
                          int64_t comp_off, int64_t own0, int32_t* skel, void* stream) {
    return skb_foreach(n_records, (skb_stream_t)stream,
                       SkbKeySkelItem{link, par, cmin, rank, (int32_t)node_lo, (int32_t)node_hi, (int32_t)id_shift,
                                      (int32_t)comp_off, (int32_t)own0, skel});
}

// branch_distance / mean / stdev of the paths headed on this rank from the sums carried through
// the ranking (tree-order fp64 sums: within 1e-12 of the reference's left-to-right sums)
int skb_slab_path_stats(const uint32_t* plen, const void* head_aux, int64_t n_paths, double* dist, double* mean,
                        double* stdev, void* stream) {
    return skb_foreach(n_paths, (skb_stream_t)stream,
                       SkbSlabStatsItem{plen, (const SkbAux*)head_aux, dist, mean, stdev});
}

// Skeleton.path_label_image (csr.py:776-790): image i64[V] must be zeroed by the caller
int skb_path_label_image(const int32_t* pindptr, const int32_t* pidx, const int64_t* lin, int64_t n_paths,
                         int64_t* image, void* stream) {
    return skb_foreach(n_paths * 32, (skb_stream_t)stream, SkbLabelImageItem{pindptr, pidx, lin, image});
}

// ---- Sholl analysis (csr.py:1332-1390), see skb_sholl.cuh ------------------------------------------------
// Per node: distance from `center` of coordinates * spacing and its shell index np.digitize(d, radii).
// center / spacing_f / spacing_i: host, 3 entries (ndim used).  radii: device f64[n_radii], monotonic
// (decreasing != 0: decreasing).  Any of bins i32[n], dist f64[n], max_bits u64[1] (pre-zeroed; receives the
// bit pattern of the largest distance) may be NULL; bins needs radii.
int skb_sholl_bins(const int64_t* coords, int64_t n, int ndim, int spacing_is_int, const double* spacing_f,
                   const int64_t* spacing_i, const double* center, const double* radii, int64_t n_radii,
                   int decreasing, int32_t* bins, double* dist, uint64_t* max_bits, void* stream) {
    if (ndim < 1 || ndim > 3) return skb_fail("ndim must be 1, 2 or 3");
    if (bins && (!radii || n_radii < 1 || n_radii > 0x7fffffffLL)) return skb_fail("shell indices need radii");
    SkbShollGeom g;
    for (int k = 0; k < 3; ++k) {
        g.center[k] = k < ndim ? center[k] : 0.0;
        g.spacing_f[k] = k < ndim ? spacing_f[k] : 1.0;
        g.spacing_i[k] = k < ndim ? spacing_i[k] : 1;
    }
    g.ndim = ndim;
    g.spacing_is_int = spacing_is_int ? 1 : 0;
    return skb_foreach(n, (skb_stream_t)stream,
                       SkbShollBinItem{g, coords, radii, (int32_t)n_radii, decreasing ? 1 : 0, bins, dist, max_bits});
}

// hist u64[n_radii + 1], zeroed by the caller: hist[s] += 1 for every DIRECTED edge whose endpoints lie in
// different shells, s = the smaller shell index (np.bincount(np.minimum(bins0, bins1)[crossings]))
int skb_sholl_count(const int32_t* indptr, const int32_t* indices, const int32_t* bins, int64_t n,
                    unsigned long long* hist, void* stream) {
    return skb_foreach(n, (skb_stream_t)stream, SkbShollCountItem{indptr, indices, bins, hist});
}

// ---- image_stats.mesh_sizes (image_stats.py:8-46), see skb_mesh.cuh ---------------------------------------
// Run starts of the background of a 2-D mask: rs u64[nwords] (nwords = ceil(ny*nx / 64)) and rsp u32[nwords+1],
// the exclusive scan of the starts per word; rsp[nwords] = number of runs.
int skb_mesh_runs(const uint64_t* bits, int64_t ny, int64_t nx, uint64_t* rs, uint32_t* rsp, void* scratch,
                  void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (ny < 1 || nx < 1 || ny > 0x7fffffffLL || nx > 0x7fffffffLL) return skb_fail("bad extent");
    SkbMeshGeom g{ny * nx, (int32_t)ny, (int32_t)nx};
    const int64_t nwords = (g.nvox + 63) / 64;
    SKB_TRY(skb_foreach(nwords, s, SkbRunStartItem{g, bits, rs, rsp}));
    return skb_scan_u32_impl(rsp, rsp, nwords, scratch, s);
}

// Components of the runs: par i32[R], len u32[R], border u8[R], size u64[R], root_border u32[R] (size and
// root_border zeroed by the caller), keepoff u32[R+1] = exclusive scan of "root of a component that does not touch
// the border"; keepoff[R] = number of such components.
int skb_mesh_link(const uint64_t* bits, int64_t ny, int64_t nx, const uint64_t* rs, const uint32_t* rsp,
                  int64_t n_runs, int32_t* par, uint32_t* len, uint8_t* border, unsigned long long* size,
                  uint32_t* root_border, uint32_t* keepoff, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (ny < 1 || nx < 1 || ny > 0x7fffffffLL || nx > 0x7fffffffLL) return skb_fail("bad extent");
    if (n_runs > 0x7fffffffLL) return skb_fail("too many background runs");
    SkbMeshGeom g{ny * nx, (int32_t)ny, (int32_t)nx};
    const int64_t nwords = (g.nvox + 63) / 64;
    SKB_TRY(skb_foreach(n_runs, s, SkbIotaItem{par}));
    SKB_TRY(skb_foreach(nwords, s, SkbRunLinkItem{g, bits, rs, rsp, par, len, border}));
    SKB_TRY(skb_foreach(n_runs, s, SkbRunBorderItem{par, border, root_border}));
    SKB_TRY(skb_foreach(n_runs, s, SkbRunSizeItem{par, len, root_border, size}));
    SKB_TRY(skb_foreach(n_runs, s, SkbRunKeepItem{par, root_border, keepoff}));
    return skb_scan_u32_impl(keepoff, keepoff, n_runs, scratch, s);
}

// sizes i64[keepoff[R]] of the kept components, in the order of their first pixel (scipy.ndimage.label's order)
int skb_mesh_emit(const int32_t* par, const uint32_t* root_border, const uint32_t* keepoff,
                  const unsigned long long* size, int64_t n_runs, int64_t* out, void* stream) {
    return skb_foreach(n_runs, (skb_stream_t)stream, SkbRunEmitItem{par, root_border, keepoff, size, out});
}

// flag i32[1] (zeroed by the caller) is raised when a foreground pixel lies on the border of the 2-D image
int skb_mesh_border_foreground(const uint64_t* bits, int64_t ny, int64_t nx, int32_t* flag, void* stream) {
    if (ny < 1 || nx < 1 || ny > 0x7fffffffLL || nx > 0x7fffffffLL) return skb_fail("bad extent");
    SkbMeshGeom g{ny * nx, (int32_t)ny, (int32_t)nx};
    return skb_foreach(2 * (ny + nx), (skb_stream_t)stream, SkbBorderFgItem{g, bits, flag});
}

// number of connected components of a symmetric CSR graph, isolated nodes included (image_stats.py:86-87):
// par i32[n] scratch, count u64[1] zeroed by the caller
int skb_cc_count(const int32_t* indptr, const int32_t* indices, int64_t n, int32_t* par, unsigned long long* count,
                 void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n, s, SkbIotaItem{par}));
    SKB_TRY(skb_foreach(n, s, SkbCcEdgeUnionItem{indptr, indices, par}));
    return skb_foreach(n, s, SkbCcRootCountItem{par, count});
}

int skb_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* scratch, void* stream) {
    return skb_scan_u32_impl(in, out, n, scratch, (skb_stream_t)stream);
}


}  // extern "C"
