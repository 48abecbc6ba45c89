// skan_b200 -- bodies of the C-ABI entry points declared in include/skan_b200.h.
//
// Included by skb_cuda.cu (sm_100a build: pointers are device pointers, `stream` is a
// cudaStream_t) and by tests/emul/skb_emul.cpp (host emulation harness for CPU unit tests).
// The including file provides:
//   int  skb_foreach(int64_t n, skb_stream_t s, Functor f)
//   int  skb_fill_bytes(void* p, int byte, size_t nbytes, skb_stream_t s)
//   int  skb_fetch_i32(const int32_t* p, int32_t* host_out, skb_stream_t s)      (synchronises)
//   int  skb_scan_u32_impl / skb_scan_u64_impl (exclusive scans, n+1 outputs)
//   int  skb_fail(const char* what)                                               (records the message)
// and the three block-cooperative stages (pack, node emit, scans) itself.

#define SKB_TRY(expr)            \
    do {                         \
        int _e = (expr);         \
        if (_e != 0) return _e;  \
    } while (0)

static SkbGeom skb_make_geom(int ndim, const int64_t* shape) {
    SkbGeom g;
    int64_t s[3] = {1, 1, 1};
    for (int i = 0; i < ndim; ++i) s[3 - ndim + i] = shape[i];
    g.nz = (int32_t)s[0];
    g.ny = (int32_t)s[1];
    g.nx = (int32_t)s[2];
    g.ndim = ndim;
    g.sz = s[1] * s[2];
    g.nvox = s[0] * s[1] * s[2];
    return g;
}

static int skb_check_geom(int ndim, const int64_t* shape) {
    if (ndim < 1 || ndim > 3) return skb_fail("ndim must be 1, 2 or 3");
    for (int i = 0; i < ndim; ++i)
        if (shape[i] < 1 || shape[i] > 0x7fffffffLL) return skb_fail("bad extent");
    return 0;
}

extern "C" {

int skb_abi_version(void) { return 1; }

// raw 3^d neighbourhood masks of the N foreground voxels            (replaces csr.py:122-144)
int skb_raw_neighbors(const uint64_t* bits, int ndim, const int64_t* shape, const int64_t* lin, int64_t n,
                      uint32_t* nb, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    SkbRawNbItem it{skb_make_geom(ndim, shape), bits, lin, nb};
    return skb_foreach(n, (skb_stream_t)stream, it);
}

// junction-junction edges: per-node forward counts, then emission       (csr.py:1004-1016)
int skb_junction_edge_count(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape,
                            const int64_t* lin, const uint32_t* nb, int64_t n, uint32_t* count, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    SkbJuncCountItem it{skb_make_geom(ndim, shape), bits, pre256, lin, nb, count};
    return skb_foreach(n, (skb_stream_t)stream, it);
}

int skb_junction_edge_emit(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape,
                           const int64_t* lin, const uint32_t* nb, const uint32_t* offset, int64_t n,
                           const double* wtab, const double* values, int height, int dtype, int32_t* ea,
                           int32_t* eb, uint8_t* ek, uint64_t* ew, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    if (height && !values) return skb_fail("height mode needs node values");
    SkbWeight wf{wtab, values, height, dtype};
    SkbJuncEmitItem it{skb_make_geom(ndim, shape), bits, pre256, lin, nb, offset, wf, ea, eb, ek, ew};
    return skb_foreach(n, (skb_stream_t)stream, it);
}

// junction MST (csr.py:980-1020).  Scratch: par i32[n], bestw u64[n], beste u32[n],
// era/erb i32[nej], flag i32[1].  Out: tree u8[nej] (1 = edge stays).  Returns rounds >= 0.
int skb_mst_junctions(const int32_t* ea, const int32_t* eb, const uint64_t* ew, int64_t nej, int64_t n,
                      int32_t* par, uint64_t* bestw, uint32_t* beste, int32_t* era, int32_t* erb, uint8_t* tree,
                      int32_t* flag, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (nej <= 0) return 0;
    SKB_TRY(skb_foreach(n, s, SkbMstInitItem{par, bestw, beste}));
    SKB_TRY(skb_fill_bytes(era, 0, (size_t)nej * 4, s));
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
}

// keep[] starts as a copy of nb[]; clear both directions of every non-tree junction edge
int skb_mst_apply(const int32_t* ea, const int32_t* eb, const uint8_t* ek, const uint8_t* tree, int64_t nej,
                  uint32_t* keep, void* stream) {
    return skb_foreach(nej, (skb_stream_t)stream, SkbMstApplyItem{ea, eb, ek, tree, keep});
}

// degrees (post-MST, csr.py:660) and graph.indptr: deg i32[n], indptr i32[n+1]
int skb_degrees_indptr(const uint32_t* keep, int64_t n, int32_t* deg, int32_t* indptr, void* scratch,
                       void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n, s, SkbPopcItem{keep, (uint32_t*)deg}));
    return skb_scan_u32_impl((const uint32_t*)deg, (uint32_t*)indptr, n, scratch, s);
}

// graph.indices / graph.data in canonical CSR order (csr.py:153-157)
int skb_csr_emit(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape, const int64_t* lin,
                 const uint32_t* keep, const int32_t* indptr, int64_t n, const double* wtab, const double* values,
                 int height, int dtype, int32_t* indices, double* data, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    if (height && !values) return skb_fail("height mode needs node values");
    SkbWeight wf{wtab, values, height, dtype};
    SkbCsrEmitItem it{skb_make_geom(ndim, shape), bits, pre256, lin, keep, indptr, wf, indices, data};
    return skb_foreach(n, (skb_stream_t)stream, it);
}

// ---- path tracing (csr.py:347-446): sampled list ranking, see skb_items.cuh ------------------
// rec i4[n] (n0, n1, indptr, degree|flags), the number of degree-2 nodes (sum of n_chain u32[64]), the
// phase-0 start offsets (startoff u32[n+1], scanned) and the initial state of the skeleton-id
// stage (par, cmin i32[n]; is_min u32[n+1])
int skb_path_records(const int32_t* indptr, const int32_t* indices, int64_t n, skb_i4* rec, uint32_t* n_chain,
                     uint32_t* startoff, int32_t* par, int32_t* cmin, uint32_t* is_min, void* scratch,
                     void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_fill_bytes(n_chain, 0, 64 * 4, s));
    SKB_TRY(skb_foreach(n, s, SkbPathRecItem{indptr, indices, rec, n_chain, startoff, par, cmin, is_min}));
    return skb_scan_u32_impl(startoff, startoff, n, scratch, s);
}

// walk starts per node, scanned in place: startoff u32[n+1].  phase 0: terminals + splitters;
// phase 1: cycle roots + uncovered splitters (covered u8[n] must be given).
int skb_path_start_count(const int32_t* indptr, const skb_i4* rec, const uint8_t* covered, int64_t n, int phase,
                         uint32_t* startoff, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (phase != 0 && !covered) return skb_fail("phase 1 needs the covered marks");
    SKB_TRY(skb_foreach(n, s, SkbStartCountItem{indptr, rec, covered, phase, startoff}));
    return skb_scan_u32_impl(startoff, startoff, n, scratch, s);
}

// su/se i32[S] (node and half-edge of every start) and the first local walk: ranked i4[S]
int skb_path_walk(const int32_t* indptr, const int32_t* indices, const skb_i4* rec, const uint32_t* startoff,
                  int64_t n, int64_t n_starts, int32_t* su, int32_t* se, skb_i4* ranked, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n, s, SkbStartFillItem{indptr, startoff, su, se}));
    return skb_foreach(n_starts, s, SkbWalkItem{indices, rec, startoff, su, se, ranked});
}

// pointer jumping over the starts (synchronous rounds, ping-pong between buf0 and buf1).
// Stops after the first round in which no start reaches its list end: every list that ends at
// a terminal is then finished, what is left cycles.  Returns 2*rounds + (index of the result buffer).
int skb_path_rank(skb_i4* buf0, skb_i4* buf1, int64_t n_starts, int32_t* flag, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (n_starts <= 0) return 0;
    skb_i4* b[2] = {buf0, buf1};
    int cur = 0, rounds = 0;
    const int blind = 3;  // rounds run before the first host check (lists of up to 8 starts)
    for (;; ++rounds) {
        if (rounds > 40) return skb_fail("list ranking did not converge");
        if (rounds >= blind || rounds == 0) SKB_TRY(skb_fill_bytes(flag, 0, 4, s));
        SKB_TRY(skb_foreach(n_starts, s, SkbJumpItem{b[cur], b[cur ^ 1], flag}));
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
}

// heads of emitted paths among the starts: sel u32[S+1] = exclusive scan of the head flags
int skb_path_select(const int32_t* su, const skb_i4* rec, const skb_i4* ranked, int64_t n_starts, int phase,
                    uint32_t* sel, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n_starts, s, SkbSelectItem{su, rec, ranked, phase, sel}));
    return skb_scan_u32_impl(sel, sel, n_starts, scratch, s);
}

// per path (ids pid_base + ...): start_node i32, plen u32 (entries), pmin i32; einfo i2[S]
int skb_path_heads(const int32_t* su, const skb_i4* rec, const skb_i4* ranked, const uint32_t* sel,
                   int64_t n_starts, int phase, int64_t pid_base, int32_t* start_node, uint32_t* plen,
                   int32_t* pmin, skb_i2* einfo, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_fill_bytes(einfo, 0xff, (size_t)n_starts * 8, s));
    return skb_foreach(n_starts, s,
                       SkbHeadsItem{su, rec, ranked, sel, phase, (int32_t)pid_base, start_node, plen, pmin, einfo});
}

// second local walk: paths.indices i32[L], paths.data f64[L] (csr.py:396,402; values NULL -> 1.0),
// pw f64[L] = weight of the edge arriving at each entry; chain nodes written are marked in
// covered u8[n] when it is given
int skb_path_emit(const int32_t* indices, const double* gdata, const skb_i4* rec, const int32_t* su,
                  const int32_t* se, const skb_i4* ranked, const skb_i2* einfo, const int32_t* pindptr,
                  const double* values, uint8_t* covered, int64_t n_starts, int32_t* pidx, double* pdata,
                  double* pw, void* stream) {
    return skb_foreach(n_starts, (skb_stream_t)stream,
                       SkbEmitItem{indices, gdata, rec, su, se, ranked, einfo, pindptr, values, covered, pidx, pdata, pw});
}

// junction-free cycles (csr.py:369-386): the minimum node of every component of uncovered
// degree-2 nodes is flagged in rec as a cycle root and in is_min.  Scratch: par i32[n].
int skb_cycle_roots(skb_i4* rec, const uint8_t* covered, int64_t n, int32_t* par, uint32_t* is_min, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n, s, SkbIotaItem{par}));
    SKB_TRY(skb_foreach(n, s, SkbCycleHookItem{rec, covered, par}));
    return skb_foreach(n, s, SkbCycleRootItem{rec, covered, par, is_min});
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
int skb_path_stats(const int32_t* pindptr, const double* pdata, const double* pw, int64_t n_paths, double* dist,
                   double* mean, double* stdev, void* stream) {
    return skb_foreach(n_paths, (skb_stream_t)stream, SkbPathStatsItem{pindptr, pdata, pw, dist, mean, stdev});
}

// the remaining columns of summarize() (csr.py:909-952)
int skb_branch_table(int64_t n_paths, int ndim, int height, int spacing_is_int, const double* spacing_f,
                     const int64_t* spacing_i, const int32_t* pindptr, const int32_t* pidx,
                     const int32_t* gindptr, const int32_t* skel_id, const int64_t* coords, const double* values,
                     int32_t* out_i32, int64_t* out_i64, double* out_f64, void* stream) {
    if (ndim < 1 || ndim > 3) return skb_fail("ndim must be 1, 2 or 3");
    if (height && !values) return skb_fail("height mode needs node values");
    SkbBranchItem it;
    it.P = n_paths;
    it.d = ndim;
    it.height = height;
    it.spacing_int = spacing_is_int;
    it.pindptr = pindptr;
    it.pidx = pidx;
    it.gindptr = gindptr;
    it.skel_id = skel_id;
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

// Skeleton.path_label_image (csr.py:776-790): image i64[V] must be zeroed by the caller
int skb_path_label_image(const int32_t* pindptr, const int32_t* pidx, const int64_t* lin, int64_t n_paths,
                         int64_t* image, void* stream) {
    return skb_foreach(n_paths, (skb_stream_t)stream, SkbLabelImageItem{pindptr, pidx, lin, image});
}

int skb_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* scratch, void* stream) {
    return skb_scan_u32_impl(in, out, n, scratch, (skb_stream_t)stream);
}

int skb_scan_u64(const uint64_t* in, uint64_t* out, int64_t n, void* scratch, void* stream) {
    return skb_scan_u64_impl(in, out, n, scratch, (skb_stream_t)stream);
}

}  // extern "C"
