/* skan_b200.h -- C-ABI of libskan_b200.so: the skeleton-to-graph hot path of jni/skan on B200.
 *
 * The reference (pure Python; /root/reference/src/skan/csr.py) has no FFI for this path: its
 * "native" side is NumPy / SciPy / numba reached in-process.  Each entry point below replaces
 * the reference lines cited beside it; the Python mirror of skan.csr (skan_b200/csr.py) calls
 * them through ctypes with torch.Tensor.data_ptr() as the only array interop.
 *
 * Conventions
 *   - every array pointer is DEVICE memory of the current CUDA device unless marked (host);
 *   - `stream` is a cudaStream_t; calls are asynchronous unless noted "synchronises";
 *   - no allocation inside: count -> caller allocates -> fill;
 *   - return value: >= 0 success (some calls return an iteration count), < 0 failure with the
 *     message available from skb_last_error() (thread-local);
 *   - no global state; safe to call from several host threads on different streams;
 *   - images are embedded as (nz, ny, nx), C order, 1 <= ndim <= 3; `shape` (host) has ndim entries;
 *   - node ids are 0-based raster ranks of the nonzero voxels (csr.py:131-132), int32;
 *   - neighbour code k27 = (dz+1)*9 + (dy+1)*3 + (dx+1), 0..26, centre 13 never set.
 */
#ifndef SKAN_B200_H
#define SKAN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element codes of the skeleton image (numpy bool is SKB_U8) */
enum skb_dtype {
    SKB_U8 = 0, SKB_I8 = 1, SKB_U16 = 2, SKB_I16 = 3, SKB_U32 = 4, SKB_I32 = 5,
    SKB_U64 = 6, SKB_I64 = 7, SKB_F32 = 8, SKB_F64 = 9
};

typedef struct skb_i2 { int32_t x, y; } skb_i2;

int skb_abi_version(void);
const char* skb_last_error(void);
int64_t skb_scan_scratch_bytes(int64_t n); /* scratch needed by any call taking `scratch` over n items */
int64_t skb_tile_voxels(void);             /* voxels per mask tile (16384) */
int skb_set_option(const char* name, int64_t value); /* tunables with no effect on results: "emit_both_ways" (1,
                                                        default: skb_run_image's second walk writes every chain from
                                                        both of its ends, which halves its longest dependent-load
                                                        chain; 0: from its head only), "splitter_bits" (0..12: one chain
                                                        voxel in 2^bits splits the chains into walks; 0, the default:
                                                        3 below 2.5 M nodes per GPU, else 4; every rank of a Z-slab
                                                        run must use the same value), "small_cycles", "junction_append",
                                                        "mst_keys", "mst_blocks_per_sm", "emit_blocks_per_sm" (A/B
                                                        switches of round 2), "pdl" (1, default: kernels
                                                        are launched with programmatic stream serialisation; also
                                                        SKAN_B200_PDL=0 in the environment), "arena_guard" (debugging
                                                        aid, see skb_debug_guard_offsets).  The
                                                        defaults were measured on B200 (profiles/r02_summary.md).
                                                        One option does change fp64 rounding: "long_path_exact" (0,
                                                        default: the statistics of a path of more than 32768 entries
                                                        are summed by a thread block in a fixed tree order; 1: by one
                                                        thread in the reference's order, bit-exact, O(longest path)) */
/* Debugging aid: after skb_set_option("arena_guard", 1) the runners leave a 256-byte gap behind every allocation they make
 * in the caller's arenas; this returns the number of gaps of the calling thread's last runner call and writes their byte
 * offsets (which = 0: work arena, 1: out arena).  A caller that filled the arenas with a pattern finds out-of-bounds writes
 * as damaged gaps (tests/test_gpu_guards.py). */
int64_t skb_debug_guard_offsets(int which, int64_t* out /*host*/, int64_t cap);
int64_t skb_kernel_launches(void);         /* kernels launched by the library since it was loaded */

/* Mask sweep -- replaces image.astype(bool), np.pad and both np.flatnonzero passes
 * (csr.py:1070, 122-123, 131).  Reads the image once; writes 1 bit per voxel (u64 words,
 * little-endian bit order, ntiles*256 words; the caller keeps two zeroed pad words on either
 * side) and the foreground count of every 16384-voxel tile.  `image` must be 16-byte aligned.
 * variant 0: 128-bit streaming loads; variant 1: TMA bulk copies staged in shared memory. */
int skb_pack_count(const void* image, int dtype, int64_t nvox, uint64_t* bits, uint32_t* tile_counts,
                   int64_t ntiles, int variant, void* stream);

/* Exclusive scan with n+1 outputs (out[n] = total); in == out allowed.  Single pass (decoupled
 * look-back).  `scratch` (here and in every call that takes one): skb_scan_scratch_bytes(n) bytes,
 * zeroed ONCE by the caller and then reusable by later calls on the same stream. */
int skb_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, void* scratch, void* stream);

/* Nodes in raster order -- replaces np.flatnonzero / np.unravel_index / _extract_values
 * (csr.py:131-132, 1088, 554-562).  tile_prefix = scanned tile counts (ntiles+1).
 * Out: lin i64[N] raveled index, coords i64[N][ndim], values f64[N] (or NULL for bool images),
 * pre256 u32[ntiles*64] = popcount prefix of `bits` per 256-bit block (rank structure that
 * stands in for skimage.util.map_array, csr.py:135,142-144).  z_origin is added to the first
 * coordinate of 3-D images (the image may be one Z-slab of a larger volume). */
int skb_emit_nodes(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                   const int64_t* shape /*host*/, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                   int64_t* coords, double* values, uint32_t* pre256, void* stream);
/* The same when lin / coords / values have room for node_cap nodes only: nodes with id >= node_cap are not
 * written (pre256 is always complete). */
int skb_emit_nodes_cap(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                       const int64_t* shape /*host*/, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                       int64_t* coords, double* values, uint32_t* pre256, int64_t node_cap, void* stream);
/* Mask sweep + scan + node emission in ONE pass over the image (csr.py:1070,122-123,131-132,1088,
 * 554-562): writes what skb_pack_count + skb_scan_u32 + skb_emit_nodes write (bits, pre256, lin,
 * coords, values) except the tile counts, without reading the mask back.  The node arrays have room
 * for node_cap nodes; *n_nodes (device) receives the true count N.  When N > node_cap only the first
 * node_cap nodes were written: bits and pre256 are complete, so allocate N and call skb_tile_prefix +
 * skb_emit_nodes.  scratch: skb_sweep_scratch_bytes(nvox) bytes, zeroed by the call itself.
 * n_nodes: device i32[2]: [0] = N, [1] = number of waits inside the kernel that timed out (always 0
 * unless the chunk hand-off protocol failed; the result is invalid then and the runner raises). */
int64_t skb_sweep_scratch_bytes(int64_t nvox);
int skb_sweep_nodes(const void* image, int dtype, int ndim, const int64_t* shape /*host*/, int64_t z_origin,
                    uint64_t* bits, int64_t ntiles, uint32_t* pre256, int64_t node_cap, int64_t* lin, int64_t* coords,
                    double* values, void* scratch, int32_t* n_nodes, void* stream);
int skb_tile_prefix(const uint32_t* pre256, const int32_t* n_nodes, int64_t ntiles, uint32_t* tile_prefix,
                    void* stream);

/* Raw 3^d neighbourhood of every node as a 27-bit mask -- replaces the N x (3^d-1) neighbour
 * table and membership test of pixel_graph (csr.py:122-144).  planes_independent != 0: the image
 * is a stack of independent 2-D images (pipe.py:140-148 processes them one by one), no neighbours
 * across the leading axis. */
int skb_raw_neighbors(const uint64_t* bits, int ndim, const int64_t* shape /*host*/, const int64_t* lin, int64_t n,
                      int planes_independent, uint32_t* nb, void* stream);
/* nb[i] &= allowed: keeps the neighbour directions k27 whose bit is set -- pixel_graph(connectivity=c) with c < ndim
 * (csr.py:51-158; scipy.ndimage.generate_binary_structure: at most c non-zero offset components). */
int skb_mask_neighbors(uint32_t* nb, int64_t n, uint32_t allowed, void* stream);
/* make_degree_image (csr.py:1173-1208): image i64[V], zeroed by the caller, receives popcount(nb[i]) at lin[i]. */
int skb_degree_image(const uint32_t* nb, const int64_t* lin, int64_t n, int64_t* image, void* stream);

/* Junction-junction edges (both ends of raw degree >= 3; csr.py:1004-1016): per-node count of
 * forward edges (only nodes a in [a0, a1): the nodes a Z-slab owns), then (after an exclusive scan
 * into `offset`) emission in (a, k27) order with a < b.  ew = fp64 bit pattern of the weight: wtab[k27] (host-computed per nputil.py:278-279) or,
 * with height != 0, hypot(values[a]-values[b] in the image dtype, wtab[k27]) (csr.py:1023-1025). */
int skb_junction_edge_count(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape /*host*/,
                            const int64_t* lin, const uint32_t* nb, int64_t n, int64_t a0, int64_t a1,
                            uint32_t* count, void* stream);
int skb_junction_edge_emit(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape /*host*/,
                           const int64_t* lin, const uint32_t* nb, const uint32_t* offset, int64_t n,
                           const double* wtab, const double* values, int height, int dtype, int32_t* ea,
                           int32_t* eb, uint8_t* ek, uint64_t* ew, void* stream);

/* Junction MST -- replaces _mst_junctions (csr.py:980-1020; scipy minimum_spanning_tree):
 * Boruvka under the strict order (weight, min id, max id).  Endpoints are node ids < n (global ids
 * when the edges of several Z-slabs were gathered).  Scratch: par i32[n], bestw u64[n], beste u32[n]
 * (only entries of edge endpoints are touched), era/erb i32[nej], flag i32[4].  Out: tree u8[nej].
 * Returns the number of rounds; synchronises once per round. */
int skb_mst_junctions(const int32_t* ea, const int32_t* eb, const uint64_t* ew, int64_t nej, int64_t n,
                      int32_t* par, uint64_t* bestw, uint32_t* beste, int32_t* era, int32_t* erb, uint8_t* tree,
                      int32_t* flag, void* stream);
/* keep[] (a copy of nb[]; n_local entries for node ids id_shift ..) loses both directions of every
 * non-tree junction edge (csr.py:1018-1019); endpoints outside the local range are skipped */
int skb_mst_apply(const int32_t* ea, const int32_t* eb, const uint8_t* ek, const uint8_t* tree, int64_t nej,
                  int64_t id_shift, int64_t n_local, uint32_t* keep, void* stream);

/* Skeleton.degrees (csr.py:660) and graph.indptr: deg i32[n], indptr i32[n+1] */
int skb_degrees_indptr(const uint32_t* keep, int64_t n, int32_t* deg, int32_t* indptr, void* scratch, void* stream);
/* graph.indices i32[E] / graph.data f64[E], canonical CSR (csr.py:153-157).  With rec != NULL the
 * same pass writes the outputs of skb_path_records (lin_splitters: sample splitters by voxel position). */
int skb_csr_emit(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape /*host*/,
                 const int64_t* lin, const uint32_t* keep, const int32_t* indptr, int64_t n, const double* wtab,
                 const double* values, int height, int dtype, int32_t* indices, double* data, skb_i4* rec,
                 const int64_t* slab /*host*/, int lin_splitters, uint32_t* startoff, int32_t* par, int32_t* cmin,
                 uint32_t* is_min, void* scratch, void* stream);

/* ---- Path tracing -- replaces _build_paths / _walk_path (csr.py:347-409) ---------------------
 * Sampled list ranking: every half-edge leaving a terminal (degree 1 or >= 3) and both half-edges
 * of every "splitter" (a 1/32 hash sample of the degree-2 nodes) start a bounded local walk to the
 * next terminal or splitter; the list of starts is ranked by pointer jumping; a second local walk
 * writes the path entries.  Phase 0 handles paths between terminals, phase 1 junction-free cycles
 * (csr.py:369-386), whose minimum node is flagged as a "cycle root" and acts as a terminal. */
typedef struct skb_i4 { int32_t x, y, z, w; } skb_i4; /* 16-byte aligned */

/* `slab` (host, 13 x i64, or NULL on a single GPU): own0, own1 (owned nodes), cnt0, cnt1 (nodes whose
 * starts are numbered), ga0, ga1, gb0, gb1 (forced-splitter ranges: slab boundary planes), id_shift
 * (global node id = local + id_shift), start_shift, start_lo, start_hi (global start ids),
 * lin_offset (global raveled index = lin + lin_offset).  See the Z-slab section below. */

/* Node records from an existing CSR: rec = 2 x i4 per node (32 B): (n0, n1, indptr[v], degree | flags)
 * and the two edge weights of a degree-2 node as (f64, f64); startoff u32[n+1] = scanned phase-0 walk
 * starts per node; par, cmin i32[n] (may be NULL) and is_min u32[n+1] = initial state of
 * skb_skeleton_ids (isolated pixels already flagged).  With lin != NULL splitters are sampled by voxel
 * position instead of node id. */
int skb_path_records(const int32_t* indptr, const int32_t* indices, const double* gdata, const int64_t* lin,
                     const int64_t* slab /*host*/, int64_t n, skb_i4* rec, uint32_t* startoff, int32_t* par,
                     int32_t* cmin, uint32_t* is_min, void* scratch, void* stream);
/* startoff u32[n+1] for phase 1 (cycle roots + uncovered splitters; covered u8[n]) */
int skb_path_start_count(const skb_i4* rec, const uint8_t* covered, const int64_t* slab /*host*/, int64_t n,
                         uint32_t* startoff, void* scratch, void* stream);
/* su/se i32[S]: node and half-edge of each start; ranked i4[S]: result of the first local walk;
 * aux (48 B per start, NULL on a single GPU): running sums and end node; stamp i32[n] (may be NULL) */
int skb_path_walk(const int32_t* indptr, const int32_t* indices, const double* gdata, const double* values,
                  const int64_t* lin, const skb_i4* rec, const uint32_t* startoff, const int64_t* slab /*host*/,
                  int64_t n, int64_t n_starts, int32_t* su, int32_t* se, skb_i4* ranked, void* aux, int32_t* stamp,
                  void* stream);
/* pointer jumping over the starts with global ids [lo, hi), ping-pong buf0/aux0 <-> buf1/aux1;
 * returns 2*rounds + index of the buffer holding the result; synchronises every round after the third */
int skb_path_rank(skb_i4* buf0, skb_i4* buf1, void* aux0, void* aux1, int64_t n_starts, int64_t lo, int64_t hi,
                  int32_t* flag, void* stream);
/* sel u32[S+1]: exclusive scan of "this start is the head of an emitted path"  */
int skb_path_select(const int32_t* su, const skb_i4* rec, const skb_i4* ranked, int64_t n_starts, int phase,
                    int64_t start_lo, uint32_t* sel, void* scratch, void* stream);
/* per path pid_base + k: start_node i32 (global id), plen u32 (entries), pmin i32 (minimum node id);
 * einfo i4[S] (single GPU) or head_end i32[P] + head_aux (48 B each) (slabs) */
int skb_path_heads(const int32_t* su, const skb_i4* rec, const skb_i4* ranked, const void* aux,
                   const uint32_t* sel, int64_t n_starts, int phase, int64_t pid_base, int64_t start_lo,
                   int64_t id_shift, int32_t* start_node, uint32_t* plen, int32_t* pmin, skb_i4* einfo,
                   int32_t* head_end, void* head_aux, void* stream);
/* paths.indices i32[L], paths.data f64[L] (csr.py:396,402,442-445; values NULL -> 1.0) and
 * pw f64[L] = weight of the edge arriving at each entry; covered u8[n] (optional) marks chain nodes;
 * pindptr NULL: the entry offset of a path is einfo.z */
int skb_path_emit(const int32_t* indices, const double* gdata, const skb_i4* rec, const int32_t* su,
                  const int32_t* se, const skb_i4* ranked, const skb_i4* einfo, const int32_t* pindptr,
                  const double* values, uint8_t* covered, int64_t id_shift, int64_t n_starts, int32_t* pidx,
                  double* pdata, double* pw, void* stream);
/* flags the minimum node of every component of uncovered degree-2 nodes in rec and in is_min;
 * scratch par i32[n] */
int skb_cycle_roots(skb_i4* rec, const uint8_t* covered, int64_t n, int32_t* par, uint32_t* is_min, void* stream);

/* Skeleton ids -- replaces csgraph.connected_components (csr.py:909) on the graph contracted to
 * terminals and paths; skel_id i32[P] = rank of the component's minimum node id.
 * par, cmin, is_min as initialised by skb_path_records / skb_cycle_roots (is_min becomes the rank table). */
int skb_skeleton_ids(const int32_t* pindptr, const int32_t* pidx, const int32_t* pmin, int64_t n,
                     int64_t n_regular, int64_t n_paths, int32_t* par, int32_t* cmin, uint32_t* is_min,
                     int32_t* skel_id, void* scratch, void* stream);

/* Per-path statistics -- replaces _compute_distances (csr.py:962-977, left-to-right sum of pw),
 * path_means and path_stdev (csr.py:792-816; NumPy reduceat summation order).  NULL outputs are skipped.
 * pdata NULL = boolean image: all of paths.data is 1.0, so mean = 1 and stdev = 0 exactly, nothing is read. */
int skb_path_stats(const int32_t* pindptr, const double* pdata, const double* pw, int64_t n_paths, double* dist,
                   double* mean, double* stdev, void* stream);

/* Branch table -- the remaining columns of summarize (csr.py:909-952) as three SoA buffers:
 *   out_i32[3][P]            skeleton_id, node_id_src, node_id_dst
 *   out_i64[1+2d][P]         branch_type, image_coord_src_*, image_coord_dst_*
 *   out_f64[2(d+h)+1][P]     coord_src_*(+height), coord_dst_*(+height), euclidean_distance
 * (coord_* carry int64 bit patterns when spacing_is_int).  spacing_f / spacing_i are host arrays.
 * Slabs: head_aux / start_node / id_shift / global_shape describe paths whose last node lives on
 * another rank (NULL / 0 on a single GPU). */
int skb_branch_table(int64_t n_paths, int ndim, int height, int spacing_is_int, const double* spacing_f /*host*/,
                     const int64_t* spacing_i /*host*/, const int32_t* pindptr, const int32_t* pidx,
                     const int32_t* gindptr, const int32_t* skel_id, const int64_t* coords, const double* values,
                     const void* head_aux, const int32_t* start_node, int64_t id_shift,
                     const int64_t* global_shape /*host*/, int32_t* out_i32, int64_t* out_i64, double* out_f64,
                     void* stream);

/* ---- Z-slab mode: one volume split over several GPUs -----------------
 * Every rank runs the voxel and node stages on its owned planes plus 2 halo planes on either side,
 * the junction MST on the gathered junction edges of all ranks, numbers walk starts for its owned
 * nodes and the adjacent halo plane exactly as the neighbour rank does, ranks its own lists, and
 * exchanges only (a) counts, (b) junction edges, (c) the ranked starts of its two boundary planes
 * ("inter-slab table"), (d) one head + one link record per path. */
int skb_rank_query(const uint64_t* bits, const uint32_t* pre256, const int64_t* pos, int64_t n, int32_t* out,
                   void* stream);
int skb_slab_covered(const skb_i4* rec, const uint32_t* startoff, const int32_t* stamp, const skb_i4* ranked,
                     int64_t n, int64_t own0, int64_t own1, uint8_t* covered, uint32_t* n_uncovered, void* stream);
/* n_rows >= n_block rows are written; the extra ones are padding (-1) */
int skb_slab_export(const skb_i4* ranked, const void* aux, int64_t n_first, int64_t last0, int64_t n_block,
                    int64_t n_rows, int64_t lo, int64_t hi, skb_i4* out, void* out_aux, void* stream);
/* table_desc (host i64): n_ranks, start_off[n_ranks+1], n_first[n_ranks], last0[n_ranks], t_off[n_ranks] */
int skb_slab_table_rank(skb_i4* buf0, skb_i4* buf1, void* aux0, void* aux1, int64_t n_table,
                        const int64_t* table_desc /*host*/, int32_t* flag, void* stream);
int skb_slab_apply(skb_i4* ranked, void* aux, const skb_i4* table, const void* table_aux, int64_t n_starts,
                   const int64_t* table_desc /*host*/, int64_t lo, int64_t hi, void* stream);
int skb_slab_head_records(const int32_t* head_end, const uint32_t* plen, const int32_t* pindptr,
                          const int32_t* pmin, const int32_t* start_node, const uint32_t* startoff,
                          const void* head_aux, const int64_t* slab /*host*/, int64_t n, int64_t n_paths,
                          int64_t n_rows, int64_t pid_off, int64_t entry_off, skb_i4* head, skb_i4* link,
                          void* stream);
/* regroup gathered buffers: desc (host i64) = n_seg, then n_seg x (src address, dst address, bytes), 16-byte units */
int skb_copy_segments(const int64_t* desc /*host*/, void* stream);
/* gathered junction edges (cap rows per rank, a < 0 = padding): local ids of rank k += shifts[k] */
int skb_slab_edge_shift(int32_t* ea, int32_t* eb, int64_t n_rows, int64_t cap, const int64_t* shifts /*host*/,
                        int64_t n_ranks, void* stream);
int skb_slab_einfo(const skb_i4* head, int64_t n_records, int64_t n_total_starts, skb_i4* einfo, void* stream);
int skb_slab_components(const skb_i4* link, int64_t n_records, int64_t n_keys, int32_t* par, int32_t* cmin,
                        int64_t node_lo, int64_t node_hi, int64_t id_shift, uint32_t* is_min, int64_t n,
                        void* scratch, void* stream);
int skb_slab_skeleton_ids(const skb_i4* link, int64_t n_records, int32_t* par, const int32_t* cmin,
                          const uint32_t* rank, int64_t node_lo, int64_t node_hi, int64_t id_shift,
                          int64_t comp_off, int64_t own0, int32_t* skel, void* stream);
int skb_slab_path_stats(const uint32_t* plen, const void* head_aux, int64_t n_paths, double* dist, double* mean,
                        double* stdev, void* stream);

/* Skeleton.path_label_image (csr.py:776-790); image i64[V] zeroed by the caller */
int skb_path_label_image(const int32_t* pindptr, const int32_t* pidx, const int64_t* lin, int64_t n_paths,
                         int64_t* image, void* stream);

/* Sholl analysis (csr.py:1332-1390).  skb_sholl_bins: per node, the distance of
 * coordinates * spacing from `center` (fp64, the arithmetic of scipy.spatial.distance_matrix, csr.py:1376-1377)
 * and its shell index np.digitize(d, radii) (csr.py:1378-1379).  center / spacing_f / spacing_i: host arrays of 3
 * (ndim used); radii: device f64[n_radii], monotonic (decreasing != 0 when decreasing).  bins i32[n], dist f64[n],
 * max_bits u64[1] (pre-zeroed; bit pattern of the largest distance, csr.py:1314-1316) may each be NULL.
 * skb_sholl_count: hist u64[n_radii+1] (pre-zeroed) += 1 at min(shell of i, shell of j) for every directed edge
 * (i, j) whose ends lie in different shells (csr.py:1380-1385; halve for the undirected count). */
int skb_sholl_bins(const int64_t* coords, int64_t n, int ndim, int spacing_is_int, const double* spacing_f /*host*/,
                   const int64_t* spacing_i /*host*/, const double* center /*host*/, const double* radii,
                   int64_t n_radii, int decreasing, int32_t* bins, double* dist, uint64_t* max_bits, void* stream);
int skb_sholl_count(const int32_t* indptr, const int32_t* indices, const int32_t* bins, int64_t n,
                    unsigned long long* hist, void* stream);

/* image_stats.mesh_sizes / image_summary (image_stats.py:8-90): the 4-connected background
 * components of a 2-D mask (scipy.ndimage.label(~skeleton), image_stats.py:31-32) as unions of horizontal runs of zero
 * bits.  skb_mesh_runs: rs u64[nwords], rsp u32[nwords+1] (nwords = ceil(ny*nx/64)); rsp[nwords] = number of runs R.
 * skb_mesh_link: par i32[R], len u32[R], border u8[R], size u64[R] and root_border u32[R] (both pre-zeroed),
 * keepoff u32[R+1]; keepoff[R] = number of components that do not touch the border (image_stats.py:33-44).
 * skb_mesh_emit: their sizes i64[keepoff[R]] in label order.  skb_mesh_border_foreground: flag i32[1] (pre-zeroed) is
 * raised when a foreground pixel lies on the border.  skb_cc_count: number of connected components of a symmetric CSR
 * graph, isolated nodes included (ndi.label(skeleton, full structure)[1], image_stats.py:86-87); par i32[n] scratch,
 * count u64[1] pre-zeroed.  scratch: skb_scan_scratch_bytes(nwords + 1) resp. (R + 1). */
int skb_mesh_runs(const uint64_t* bits, int64_t ny, int64_t nx, uint64_t* rs, uint32_t* rsp, void* scratch, void* stream);
int skb_mesh_link(const uint64_t* bits, int64_t ny, int64_t nx, const uint64_t* rs, const uint32_t* rsp, int64_t n_runs,
                  int32_t* par, uint32_t* len, uint8_t* border, unsigned long long* size, uint32_t* root_border,
                  uint32_t* keepoff, void* scratch, void* stream);
int skb_mesh_emit(const int32_t* par, const uint32_t* root_border, const uint32_t* keepoff, const unsigned long long* size,
                  int64_t n_runs, int64_t* out, void* stream);
int skb_mesh_border_foreground(const uint64_t* bits, int64_t ny, int64_t nx, int32_t* flag, void* stream);
int skb_cc_count(const int32_t* indptr, const int32_t* indices, int64_t n, int32_t* par, unsigned long long* count,
                 void* stream);

/* ---- The whole hot path for one image (or one stack of 2-D images) in one call ---------------------
 * Native twin of the per-stage sequencing: same kernels in the same order; allocation is a bump
 * pointer into two caller-provided device arenas (`work`: scratch, `out`: results), the counts that
 * size allocations are read back inside (~8 stream synchronisations).  Returns 0, or 1 when an arena is
 * too small (result->work_need / out_need tell how much to provide on the retry), or < 0 on error.
 * result->off[SKB_O_*] are byte offsets of the result arrays inside `out` (-1 = absent),
 * result->counts[SKB_C_*] the sizes (see skan_b200/csrc/skb_run.inl for the two enums). */
typedef struct skb_run_params {
    const void* image; int32_t dtype, ndim; int64_t shape[3]; int64_t z_origin;
    int32_t image_stack, height, pack_variant, stop_after_graph, want_values, time_sweep;
    const double* wtab;                 /* device: 27 neighbour weights (nputil.py:278-279, host-computed) */
    double spacing_f[3]; int64_t spacing_i[3]; int32_t spacing_is_int, table_ndim;
    void* work; int64_t work_bytes; void* out; int64_t out_bytes; void* stream;
    int64_t node_cap_hint;              /* capacity hint for the node arrays (fused sweep: 0 = V/256 + 4096; else 0 = none) */
    int64_t hint[8];                    /* capacity hints: [0] junction edges, [1] edges, [2] walk starts, [3] regular
                                           paths; 0 = none.  With a hint, the read-back of the count overlaps the first
                                           kernel that fills the arrays it sizes (counts[10] such overlaps per call);
                                           a hint that proves too small costs that kernel a second launch (counts[11]).
                                           Results never depend on the hints. */
} skb_run_params;
typedef struct skb_run_result { int64_t counts[16]; int64_t off[32]; int64_t work_need, out_need; int64_t stage_ns[32]; } skb_run_result;
int skb_run_image(const skb_run_params* params /*host*/, skb_run_result* result /*host*/);

/* ---- One volume over the GPUs of a node: Z-slab decomposition in one call per rank ------------------
 * (the reference has no distributed layer.)  A communicator is created
 * collectively by the ranks of ONE node: it maps a control block in POSIX shared memory (host
 * integers: counts and barriers) and every rank's exchange arena, device memory exported with CUDA
 * IPC, so that the kernels of a rank load the rows a peer produced in place over NVLink -- no
 * collective library on the data path.  skb_run_slab is the native twin of skan_b200/slab.py: same
 * kernels, same order; returns 0, 1 (an arena is too small: work_need / out_need), 2 (the exchange
 * arena is too small: exchange_need; destroy and recreate the communicator on every rank), 3 (another
 * rank returned 1 or failed: call skb_comm_reset on every rank between two barriers of the host's own
 * process group and run again) or < 0. */
int skb_comm_create(int rank, int world, const char* name /* same unique string on every rank */,
                    int64_t exchange_bytes, void** comm_out);
int skb_comm_destroy(void* comm);
int skb_comm_reset(void* comm);
int skb_comm_allgather_i64(void* comm, const int64_t* vals /*host*/, int n, int64_t* out /*host, world x n*/);
typedef struct skb_slab_params {
    const void* image;                  /* planes [ze0, ze0 + nz_ext) of the volume */
    int32_t dtype, pack_variant; int64_t global_shape[3]; int64_t z0, z1, ze0, nz_ext;
    int32_t want_values, time_sweep; const double* wtab;
    double spacing_f[3]; int64_t spacing_i[3]; int32_t spacing_is_int, halo_exchange;
    int32_t height, reserved;           /* height: value_is_height (csr.py:148-151, 932-948); needs want_values */
    void* work; int64_t work_bytes; void* out; int64_t out_bytes; void* stream;
} skb_slab_params;
typedef struct skb_slab_result { int64_t counts[32]; int64_t off[32]; int64_t work_need, out_need, exchange_need; int64_t host_ns[16]; int64_t dev_ns[16]; } skb_slab_result;
int skb_run_slab(const skb_slab_params* params /*host*/, void* comm, skb_slab_result* result /*host*/);

#ifdef __cplusplus
}
#endif
#endif /* SKAN_B200_H */
