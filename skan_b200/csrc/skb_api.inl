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

// connected components + node classes (csr.py:909; cycle detection for csr.py:369-386).
// Scratch: par i32[n], has_terminal u8[n].  Out: label i32[n] (= minimum node id of the
// component), ntype u8[n], root_rank u32[n+1] (rank of a component's minimum among all minima).
int skb_components(const int32_t* indptr, const int32_t* indices, int64_t n, int32_t* par, uint8_t* has_terminal,
                   int32_t* label, uint8_t* ntype, uint32_t* root_rank, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n, s, SkbIotaItem{par}));
    SKB_TRY(skb_fill_bytes(has_terminal, 0, (size_t)n, s));
    SKB_TRY(skb_foreach(n, s, SkbCcHookItem{indptr, indices, par}));
    SKB_TRY(skb_foreach(n, s, SkbCcFlattenItem{indptr, par, label, has_terminal}));
    SKB_TRY(skb_foreach(n, s, SkbNodeTypeItem{indptr, label, has_terminal, ntype, root_rank}));
    return skb_scan_u32_impl(root_rank, root_rank, n, scratch, s);
}

// list ranking of the half-edge lists (csr.py:347-409).  arcs i2[ne], flag i32[1]. Returns rounds.
int skb_rank_arcs(const int32_t* indptr, const int32_t* indices, const uint8_t* ntype, int64_t n, int64_t ne,
                  skb_i2* arcs, int32_t* flag, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    if (ne <= 0) return 0;
    SKB_TRY(skb_foreach(n, s, SkbArcInitItem{indptr, indices, ntype, arcs}));
    int rounds = 0;
    for (;; ++rounds) {
        if (rounds > 48) return skb_fail("list ranking did not converge");
        SKB_TRY(skb_fill_bytes(flag, 0, 4, s));
        // three launches of two jumps each between host checks: 2^6 arcs per check at least
        SKB_TRY(skb_foreach(ne, s, SkbArcJumpItem{arcs, flag, 2}));
        SKB_TRY(skb_fill_bytes(flag, 0, 4, s));
        SKB_TRY(skb_foreach(ne, s, SkbArcJumpItem{arcs, flag, 2}));
        SKB_TRY(skb_fill_bytes(flag, 0, 4, s));
        SKB_TRY(skb_foreach(ne, s, SkbArcJumpItem{arcs, flag, 2}));
        int32_t any = 0;
        SKB_TRY(skb_fetch_i32(flag, &any, s));
        if (!any) break;
    }
    return rounds;
}

// number of paths starting at every node: count u64[n+1] -> exclusive scan in place;
// count[n] = (cycle paths << 32) | regular paths
int skb_path_count(const int32_t* indptr, const uint8_t* ntype, const skb_i2* arcs, int64_t n, uint64_t* count,
                   void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_foreach(n, s, SkbPathCountItem{indptr, ntype, arcs, count}));
    return skb_scan_u64_impl(count, count, n, scratch, s);
}

// path ids: start_node i32[P], plen u32[P+1] -> scanned into paths.indptr i32[P+1];
// einfo i2[ne] tags every emitted list by its (reversed) last half-edge
int skb_path_starts(const int32_t* indptr, const uint8_t* ntype, const skb_i2* arcs, const uint64_t* prefix,
                    int64_t n, int64_t ne, int64_t n_regular, int64_t n_paths, int32_t* start_node,
                    int32_t* pindptr, skb_i2* einfo, void* scratch, void* stream) {
    skb_stream_t s = (skb_stream_t)stream;
    SKB_TRY(skb_fill_bytes(einfo, 0xff, (size_t)ne * 8, s));
    SKB_TRY(skb_foreach(
        n, s, SkbPathStartItem{indptr, ntype, arcs, prefix, (uint32_t)n_regular, start_node, (uint32_t*)pindptr, einfo}));
    return skb_scan_u32_impl((const uint32_t*)pindptr, (uint32_t*)pindptr, n_paths, scratch, s);
}

// paths.indices i32[L], paths.data f64[L] (csr.py:396,402: node values, 1.0 without values)
int skb_path_emit(const int32_t* indices, const skb_i2* arcs, const skb_i2* einfo, const int32_t* pindptr,
                  const int32_t* start_node, const double* values, int64_t ne, int32_t* pidx, double* pdata,
                  void* stream) {
    return skb_foreach(ne, (skb_stream_t)stream,
                       SkbPathEmitItem{indices, arcs, einfo, pindptr, start_node, values, pidx, pdata});
}

// branch_distance / mean / stdev per path (csr.py:962-977, 792-816); null outputs are skipped
int skb_path_stats(const int32_t* gindptr, const int32_t* gindices, const double* gdata, const int32_t* pindptr,
                   const int32_t* pidx, const double* pdata, int64_t n_paths, double* dist, double* mean,
                   double* stdev, void* stream) {
    return skb_foreach(n_paths, (skb_stream_t)stream,
                       SkbPathStatsItem{gindptr, gindices, gdata, pindptr, pidx, pdata, dist, mean, stdev});
}

// the remaining columns of summarize() (csr.py:909-952)
int skb_branch_table(int64_t n_paths, int ndim, int height, int spacing_is_int, const double* spacing_f,
                     const int64_t* spacing_i, const int32_t* pindptr, const int32_t* pidx,
                     const int32_t* gindptr, const int32_t* label, const uint32_t* root_rank,
                     const int64_t* coords, const double* values, int32_t* out_i32, int64_t* out_i64,
                     double* out_f64, void* stream) {
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
    it.label = label;
    it.root_rank = root_rank;
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
