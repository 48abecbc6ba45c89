// skan_b200 -- one volume over the GPUs of a node: Z-slab decomposition, native sequencing.
//
// C++ twin of skan_b200/slab.py: the same entry points in the same order, but
//   * allocation is a bump pointer into caller-provided device arenas, and the counts that size the next
//     allocation are read back here (no interpreter between the kernels);
//   * what crosses ranks does not go through a collective library.  Every rank owns an *exchange arena*
//     in device memory that all ranks of the node have mapped (CUDA IPC); producers write their rows
//     there, and after a host barrier the consumers' kernels load the peers' rows in place over NVLink
//     (skb_copy_segments / SkbPeerSumItem with peer pointers).  Host integers (counts, and the barrier
//     itself) go through a sequence-numbered table in POSIX shared memory: a few microseconds instead of
//     a collective launch plus a device round trip.
// Included after skb_run.inl by skb_cuda.cu and tests/emul/skb_emul.cpp; the including file provides
// skb_xalloc / skb_xopen / skb_xclose / skb_xfree / skb_xunlink (the arena: CUDA IPC, or shared memory
// in the emulation harness) and skb_stream_sync.
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#define SKB_COMM_MAXV 15
#define SKB_RUN_NEED_EXCHANGE 2  // skb_run_slab: the exchange arena is too small (result->exchange_need)
#define SKB_RUN_RESTART 3        // skb_run_slab: another rank gave up (its arenas were too small, or it failed)
#define SKB_XFLAG_BYTES 8192     // tail of every exchange arena: the flags of the device-side barriers

struct SkbComm {
    int rank, world;
    char name[48];
    int64_t* ctl;  // control block in shared memory: handles, ready flags, count table
    size_t ctl_bytes;
    int64_t seq;
    void* xbase[SKB_MAX_RANKS];  // every rank's exchange arena as seen from this process
    int64_t xbytes;
    int64_t hint_max_e, hint_edges;  // capacities learnt from the previous call (0: none): the largest per-rank junction
                                     // edge count (the same on every rank) and this rank's edge count
    int64_t epoch;   // calls of skb_run_slab so far: consecutive calls use alternate halves of the exchange arena, so a
    int64_t xlimit;  // region is rewritten only after a whole intervening call (whose barriers every reader has passed)
    int64_t xgen, xcount;  // device-side barriers (skb_xbar_launch): generation (advanced by skb_comm_reset) and count
};

// The last SKB_XFLAG_BYTES of every exchange arena hold the flags of the device-side barriers.
static int64_t skb_comm_xdata(const SkbComm* c) { return c->xbytes - SKB_XFLAG_BYTES; }

// Barrier between the ranks in stream order, no host thread involved: what this rank's earlier kernels wrote into its
// exchange arena can be read by the peers' kernels that follow THEIR barrier, and vice versa.  The kernels launched
// until the next read-back are "guarded": should a peer have given the call up, they return at once (g_skb_poison)
// and the read-back reports it.
static int skb_comm_xbar(SkbComm* c, skb_stream_t s) {
    if (c->world == 1) return 0;
    const uint64_t seq = ((uint64_t)c->xgen << 32) | (uint64_t)(++c->xcount);
    SKB_TRY(skb_xbar_launch(c->xbase, c->rank, c->world, skb_comm_xdata(c), seq, 0, s));
    g_skb_guard = 1;
    return 0;
}

static double skb_now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int64_t* skb_comm_handles(SkbComm* c) { return c->ctl; }                              // world x 8 words
static int64_t* skb_comm_ready(SkbComm* c) { return c->ctl + c->world * 8; }                 // world words
static int64_t* skb_comm_abort(SkbComm* c) { return c->ctl + c->world * 9; }                 // 1 word
static int64_t* skb_comm_table(SkbComm* c, int b) {                                          // 2 x world x 16 words
    return c->ctl + c->world * 9 + 1 + (int64_t)b * c->world * (1 + SKB_COMM_MAXV);
}

// all-gather of n <= 15 host integers per rank; also the barrier of the exchanges (a rank gets here only
// after it has synchronised its stream, so what it produced before is complete when its row appears)
static int skb_comm_allgather(SkbComm* c, const int64_t* vals, int n, int64_t* out) {
    if (n > SKB_COMM_MAXV) return skb_fail("too many values in a count exchange");
    const int64_t seq = ++c->seq;
    int64_t* tab = skb_comm_table(c, (int)(seq & 1));
    int64_t* row = tab + (int64_t)c->rank * (1 + SKB_COMM_MAXV);
    for (int i = 0; i < n; ++i) row[1 + i] = vals[i];
    __atomic_store_n(row, seq, __ATOMIC_RELEASE);
    const double t0 = skb_now_s();
    for (int k = 0; k < c->world; ++k) {
        const int64_t* r = tab + (int64_t)k * (1 + SKB_COMM_MAXV);
        int64_t spins = 0;
        while (__atomic_load_n(r, __ATOMIC_ACQUIRE) != seq) {
            if ((++spins & 63) == 0 && __atomic_load_n(skb_comm_abort(c), __ATOMIC_ACQUIRE) != 0) return SKB_RUN_RESTART;
            if ((spins & 1023) == 0) {
                if (skb_now_s() - t0 > 60.0) return skb_fail("count exchange timed out: a rank of the slab decomposition died");
                if (spins > (1 << 16)) sched_yield();
            }
        }
        for (int i = 0; i < n; ++i) out[(int64_t)k * n + i] = r[1 + i];
    }
    return 0;
}

extern "C" {

// Collective over the ranks of one node.  name: the same unique string on every rank (<= 40 characters).
int skb_comm_create(int rank, int world, const char* name, int64_t exchange_bytes, void** comm_out) {
    if (world < 1 || world > SKB_MAX_RANKS || rank < 0 || rank >= world) return skb_fail("bad rank / world size");
    if (!name || strlen(name) > 40) return skb_fail("bad communicator name");
    SkbComm* c = new SkbComm();
    c->rank = rank;
    c->world = world;
    snprintf(c->name, sizeof(c->name), "%s", name);
    c->seq = 0;
    c->epoch = 0;
    c->xgen = c->xcount = 0;
    c->hint_max_e = c->hint_edges = 0;
    c->xbytes = ((exchange_bytes + 8191) & ~(int64_t)8191) + SKB_XFLAG_BYTES;
    c->xlimit = c->xbytes;
    for (int k = 0; k < SKB_MAX_RANKS; ++k) c->xbase[k] = nullptr;
    char shm[64];
    snprintf(shm, sizeof(shm), "/skbc_%s", name);
    c->ctl_bytes = (size_t)(world * 9 + 1 + 2 * world * (1 + SKB_COMM_MAXV)) * 8;
    int fd = shm_open(shm, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return skb_fail("shm_open of the control block failed");
    if (ftruncate(fd, (off_t)c->ctl_bytes) != 0) {
        close(fd);
        return skb_fail("ftruncate of the control block failed");
    }
    void* p = mmap(nullptr, c->ctl_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return skb_fail("mmap of the control block failed");
    c->ctl = (int64_t*)p;
    uint8_t handle[64];
    SKB_TRY(skb_xalloc(name, rank, c->xbytes, &c->xbase[rank], handle));
    memcpy(skb_comm_handles(c) + rank * 8, handle, 64);
    __atomic_store_n(skb_comm_ready(c) + rank, (int64_t)1, __ATOMIC_RELEASE);
    const double t0 = skb_now_s();
    for (int k = 0; k < world; ++k) {
        while (__atomic_load_n(skb_comm_ready(c) + k, __ATOMIC_ACQUIRE) != 1) {
            if (skb_now_s() - t0 > 120.0) return skb_fail("communicator creation timed out");
            sched_yield();
        }
        if (k == rank) continue;
        memcpy(handle, skb_comm_handles(c) + k * 8, 64);
        SKB_TRY(skb_xopen(handle, c->xbytes, &c->xbase[k]));
    }
    int64_t one = 1, all[SKB_MAX_RANKS];
    SKB_TRY(skb_comm_allgather(c, &one, 1, all));  // everybody has mapped everything: the names can go
    if (rank == 0) shm_unlink(shm);
    skb_xunlink(name, rank);
    *comm_out = c;
    return 0;
}

int skb_comm_destroy(void* comm) {
    SkbComm* c = (SkbComm*)comm;
    if (!c) return 0;
    for (int k = 0; k < c->world; ++k) {
        if (!c->xbase[k]) continue;
        if (k == c->rank) skb_xfree(c->xbase[k], c->xbytes, c->name, c->rank);
        else skb_xclose(c->xbase[k], c->xbytes);
    }
    if (c->ctl) munmap(c->ctl, c->ctl_bytes);
    delete c;
    return 0;
}

// After skb_run_slab returned 1 or 3 on some rank: every rank calls this between two barriers of the
// host's own process group (the ranks no longer agree on the exchange sequence), then runs again.
int skb_comm_reset(void* comm) {
    SkbComm* c = (SkbComm*)comm;
    c->seq = 0;
    c->epoch = 0;
    ++c->xgen;  // the flags of the abandoned call (and its abort marks) are smaller than anything this generation waits for
    c->xcount = 0;
    g_skb_guard = 0;
    SKB_TRY(skb_poison_reset());
    c->hint_max_e = c->hint_edges = 0;
    for (int b = 0; b < 2; ++b) {
        int64_t* row = skb_comm_table(c, b) + (int64_t)c->rank * (1 + SKB_COMM_MAXV);
        for (int i = 0; i <= SKB_COMM_MAXV; ++i) row[i] = 0;
    }
    if (c->rank == 0) __atomic_store_n(skb_comm_abort(c), (int64_t)0, __ATOMIC_RELEASE);
    return 0;
}

// out: world x n host integers (tests, and host code that wants the same channel)
int skb_comm_allgather_i64(void* comm, const int64_t* vals, int n, int64_t* out) {
    return skb_comm_allgather((SkbComm*)comm, vals, n, out);
}

}  // extern "C"

// ---- the runner -----------------------------------------------------------------------------------------
// indices into skb_slab_result.off (byte offsets into the `out` arena; -1 = absent)
enum {
    SKB_SO_LIN = 0, SKB_SO_COORDS, SKB_SO_VALUES, SKB_SO_DEGREES, SKB_SO_INDPTR, SKB_SO_INDICES, SKB_SO_DATA,
    SKB_SO_PLEN, SKB_SO_PIDX, SKB_SO_PDATA, SKB_SO_PW, SKB_SO_DIST, SKB_SO_MEAN, SKB_SO_STDEV, SKB_SO_T32,
    SKB_SO_T64, SKB_SO_TF, SKB_SO_START_NODE, SKB_SO_COUNT
};
// indices into skb_slab_result.counts
enum {
    SKB_SC_NODES_EXT = 0, SKB_SC_OWN0, SKB_SC_OWN1, SKB_SC_ID_SHIFT, SKB_SC_NODE_OFFSET, SKB_SC_EDGE_OFFSET,
    SKB_SC_EDGES_OWNED, SKB_SC_NODES_TOTAL, SKB_SC_EDGES_TOTAL, SKB_SC_REGULAR, SKB_SC_CYCLES, SKB_SC_PATHS_TOTAL,
    SKB_SC_ENTRIES_TOTAL, SKB_SC_REGULAR_TOTAL, SKB_SC_REGULAR_OFFSET, SKB_SC_CYCLE_OFFSET, SKB_SC_ENTRY_OFF_REGULAR,
    SKB_SC_ENTRY_OFF_CYCLES, SKB_SC_STARTS, SKB_SC_TABLE_RECORDS, SKB_SC_SYNCS, SKB_SC_EXCHANGES, SKB_SC_EDGES_EXT,
    SKB_SC_SWEEP_NS, SKB_SC_DEVICE_BARRIERS, SKB_SC_COUNT
};

struct skb_slab_params {
    const void* image;      // planes [ze0, ze0 + nz_ext) of the volume, C order
    int32_t dtype, pack_variant;
    int64_t global_shape[3];
    int64_t z0, z1, ze0, nz_ext;
    int32_t want_values, time_sweep;
    const double* wtab;     // device, 27 weights
    double spacing_f[3];
    int64_t spacing_i[3];
    int32_t spacing_is_int, halo_exchange;
    int32_t height, reserved;  // height: value_is_height (edge weights hypot(height difference, distance), the pixel
                               // value as an extra coordinate of the branch table); needs want_values
    void* work;
    int64_t work_bytes;
    void* out;
    int64_t out_bytes;
    void* stream;
};

struct skb_slab_result {
    int64_t counts[32];
    int64_t off[32];
    int64_t work_need, out_need, exchange_need;
    int64_t host_ns[16];  // host clock at the section boundaries (the sections end in a synchronisation), from the call's start
    int64_t dev_ns[16];   // time_sweep == 2: device time of every section (events on the stream at the same boundaries)
};

#define SKB_SNEED(cond)                                                        \
    do {                                                                       \
        if (!(cond)) {                                                         \
            res->work_need = work.need + work.need / 2 + (1 << 20);            \
            res->out_need = out.need + out.need / 2 + (1 << 20);               \
            return SKB_RUN_NEED_MORE;                                          \
        }                                                                      \
    } while (0)

// fetch up to SKB_MAX_FETCH device ints in one synchronisation
#define SKB_SFETCH(n_, ptrs_, vals_)                           \
    do {                                                       \
        SKB_TRY(skb_fetch_i32s((ptrs_), (n_), (vals_), s));    \
        ++nsync;                                               \
        g_skb_guard = 0; /* the read-back reported no abandoned barrier */ \
    } while (0)

static int64_t skb_round16(int64_t x) { return x < 16 ? 16 : (x + 15) / 16 * 16; }

// One all-gather of several per-rank arrays of `cap` rows each (the C++ twin of slab._Exchange): the
// producers write into this rank's block of the exchange arena; gather() = stream sync + host barrier +
// one copy kernel whose sources are the peers' blocks.
struct SkbExchange {
    SkbComm* c;
    int64_t cap, chunk, base;  // rows per rank, bytes per rank, offset of the block in every arena
    int nf;
    int64_t row_bytes[4], foff[4];
    bool ok;
    void init(SkbComm* comm, int64_t* xoff, int64_t cap_rows, int n_fields, const int64_t* rb) {
        c = comm;
        cap = skb_round16(cap_rows);
        nf = n_fields;
        int64_t o = 0;
        for (int f = 0; f < nf; ++f) {
            row_bytes[f] = rb[f];
            foff[f] = o;
            o += cap * rb[f];
        }
        chunk = o;
        base = (*xoff + 255) & ~(int64_t)255;
        *xoff = base + chunk;
        ok = *xoff <= comm->xlimit;
    }
    void* field(int f, int k) const { return (uint8_t*)c->xbase[k] + base + foff[f]; }
    void* mine(int f) const { return field(f, c->rank); }
};

static int skb_run_slab_impl(const skb_slab_params* prm, SkbComm* comm, skb_slab_result* res);

// A rank that cannot go on (an arena too small: 1, or an error: < 0) raises the abort word of the control
// block, so that the ranks waiting for it in an exchange return SKB_RUN_RESTART instead of timing out.
extern "C" int skb_run_slab(const skb_slab_params* prm, void* comm_, skb_slab_result* res) {
    SkbComm* comm = (SkbComm*)comm_;
    if (!comm) return skb_fail("skb_run_slab needs a communicator");
    g_skb_guard = 0;
    int rc = skb_run_slab_impl(prm, comm, res);
    g_skb_guard = 0;
    if (rc < 0 || rc == SKB_RUN_NEED_MORE) __atomic_store_n(skb_comm_abort(comm), (int64_t)1, __ATOMIC_RELEASE);
    if ((rc < 0 || rc == SKB_RUN_NEED_MORE || rc == SKB_RUN_RESTART) && comm->world > 1) {
        // peers may be waiting on the device for this rank's next barrier: end every wait of this generation
        const uint64_t seq = ((uint64_t)comm->xgen << 32) | 0xffffffffull;
        skb_xbar_launch(comm->xbase, comm->rank, comm->world, skb_comm_xdata(comm), seq, 1, (skb_stream_t)prm->stream);
    }
    return rc;
}

// offsets (into the work arena: which = 0, the out arena: which = 1) of the guard gaps of this thread's last runner call
extern "C" int64_t skb_debug_guard_offsets(int which, int64_t* out, int64_t cap) {
    if (which < 0 || which > 1) return 0;
    const int n = g_skb_guards.n[which];
    for (int i = 0; i < n && i < cap; ++i) out[i] = g_skb_guards.off[which][i];
    return n;
}

static int skb_run_slab_impl(const skb_slab_params* prm, SkbComm* comm, skb_slab_result* res) {
    skb_stream_t s = (skb_stream_t)prm->stream;
    const int world = comm->world, rank = comm->rank;
    const int64_t NZ = prm->global_shape[0], NY = prm->global_shape[1], NX = prm->global_shape[2];
    const int64_t z0 = prm->z0, z1 = prm->z1, ze0 = prm->ze0, ze1 = prm->ze0 + prm->nz_ext;
    if (!(0 <= ze0 && ze0 <= z0 && z0 < z1 && z1 <= ze1 && ze1 <= NZ) || z1 - z0 < 2)
        return skb_fail("bad slab bounds (each rank needs at least two owned planes)");
    const bool has_lo = z0 > 0, has_hi = z1 < NZ;
    if ((has_lo && z0 - ze0 < 2) || (has_hi && ze1 - z1 < 2)) return skb_fail("slab mode needs a halo of at least 2 planes");
    const bool xhalo = prm->halo_exchange == 1;  // halo exchange: the image holds the owned planes [z0, z1) only
    if (xhalo) {
        if (ze0 != (has_lo ? z0 - 2 : z0) || ze1 != (has_hi ? z1 + 2 : z1))
            return skb_fail("halo exchange: the extended slab is the owned planes plus two planes towards every neighbour");
        if ((NY * NX) % SKB_TILE_VOX) return skb_fail("halo exchange needs planes of a multiple of 16384 voxels (pass the halo planes with the slab otherwise)");
        if (prm->want_values) return skb_fail("halo exchange is implemented for boolean volumes (pass the halo planes with the slab for valued images)");
    }
    const int height = prm->height == 1 ? 1 : 0;
    if (height && !prm->want_values) return skb_fail("value_is_height needs a numeric (non-boolean) image");
    const int64_t shape[3] = {prm->nz_ext, NY, NX};
    SKB_TRY(skb_check_geom(3, shape));
    SkbGeom g = skb_make_geom(3, shape);
    const int64_t sz = NY * NX, V = g.nvox;
    const int64_t ntiles = (V + SKB_TILE_VOX - 1) / SKB_TILE_VOX;
    int64_t nsync = 0, nexch = 0, nxbar = 0;
    const double t_call = skb_now_s();
    int n_mark = 0;
    for (int i = 0; i < 16; ++i) res->host_ns[i] = res->dev_ns[i] = 0;
    const bool dev_clock = prm->time_sweep == 2;
    uint8_t dev_marked[16] = {};
    if (dev_clock) {
        SKB_TRY(skb_stage_mark(0, s));
        dev_marked[0] = 1;
    }
#define SKB_HOST_MARK()                                                                         \
    do {                                                                                        \
        if (n_mark < 15) {                                                                      \
            res->host_ns[n_mark++] = (int64_t)((skb_now_s() - t_call) * 1e9);                     \
            if (dev_clock) {                                                                    \
                SKB_TRY(skb_stage_mark(n_mark, s));                                             \
                dev_marked[n_mark] = 1;                                                         \
            }                                                                                   \
        }                                                                                       \
    } while (0)
    for (int i = 0; i < 32; ++i) res->counts[i] = 0;
    for (int i = 0; i < 32; ++i) res->off[i] = -1;
    res->work_need = res->out_need = res->exchange_need = 0;
    SkbArena work, out;
    work.init(prm->work, prm->work_bytes, 0);
    out.init(prm->out, prm->out_bytes, 1);
    auto out_off = [&](const void* p) -> int64_t { return p ? (int64_t)((const uint8_t*)p - out.base) : -1; };
    // bump pointer into the exchange arena, the same on every rank; this call's half of the arena
    const int64_t xhalf = (skb_comm_xdata(comm) / 2) & ~(int64_t)255;
    int64_t xoff = (comm->epoch & 1) ? xhalf : 0;
    const int64_t xoff0 = xoff;
    comm->xlimit = xoff + xhalf;
    ++comm->epoch;
    int64_t vals[SKB_MAX_FETCH];
    int64_t all[SKB_MAX_RANKS * SKB_COMM_MAXV];
    const int variant = prm->pack_variant == 0 ? 0 : 1;

    // ---- voxel and node stages on the extended slab -------------------------------------------------------
    uint64_t* store = work.take<uint64_t>(ntiles * SKB_TILE_WORDS + SKB_MASK_PAD + 2);
    uint32_t* tile_prefix = work.take<uint32_t>(ntiles + 1);
    void* scratch = work.take<uint8_t>(skb_scan_scratch_bytes(ntiles + 1));
    int32_t* flag = work.take<int32_t>(4);
    int32_t* small = work.take<int32_t>(32);  // device scalars: rank queries, sums, flags
    int64_t* posd = work.take<int64_t>(8);
    uint32_t* pre256 = work.take<uint32_t>(ntiles * 64);
    SKB_SNEED(work.ok);
    // the mask starts one 128-byte line into its (256-byte aligned) allocation: every 256-bit block of the rank
    // structure is one 32-byte sector, every tile a whole number of lines; the words before and after it are zero
    uint64_t* bits = store + SKB_MASK_PAD;
    SKB_TRY(skb_fill_bytes(store, 0, SKB_MASK_PAD * 8, s));
    SKB_TRY(skb_fill_bytes(store + SKB_MASK_PAD + ntiles * SKB_TILE_WORDS, 0, 16, s));
    SKB_TRY(skb_fill_bytes(scratch, 0, (size_t)skb_scan_scratch_bytes(ntiles + 1), s));
    void* ev = nullptr;
    if (!xhalo) {
        if (prm->time_sweep) SKB_TRY(skb_timer_start(&ev, s));
        SKB_TRY(skb_pack_count(prm->image, prm->dtype, V, bits, tile_prefix, ntiles, variant, s));
        if (prm->time_sweep) SKB_TRY(skb_timer_stop(ev, s));
    } else {
        // Halo exchange: the image holds the owned planes only.  They are swept into
        // their place in the extended mask; every rank parks the packed mask of its first and last two owned planes
        // in its exchange arena, and after the barrier reads its neighbours' planes from theirs over NVLink (2 planes
        // of 1 bit per voxel per face: 1 MB at 2048^2); the halo tiles are counted from the bits that arrived.
        const int64_t lo_planes = z0 - ze0, hi_planes = ze1 - z1, own_planes = z1 - z0;
        const int64_t face_bytes = 2 * sz / 8;  // two planes of mask
        const int64_t tile_lo = lo_planes * sz / SKB_TILE_VOX, tile_own = own_planes * sz / SKB_TILE_VOX;
        uint64_t* own_bits = bits + lo_planes * sz / 64;
        if (prm->time_sweep) SKB_TRY(skb_timer_start(&ev, s));
        SKB_TRY(skb_pack_count(prm->image, prm->dtype, own_planes * sz, own_bits, tile_prefix + tile_lo, tile_own, variant, s));
        if (prm->time_sweep) SKB_TRY(skb_timer_stop(ev, s));
        SkbExchange exh;
        const int64_t rbh[2] = {1, 1};
        exh.init(comm, &xoff, face_bytes, 2, rbh);
        if (!exh.ok) {
            res->exchange_need = 8 * (xoff - xoff0) + (1 << 23);
            return SKB_RUN_NEED_EXCHANGE;
        }
        SKB_TRY(skb_copy_bytes(exh.mine(0), own_bits, (size_t)face_bytes, s));
        SKB_TRY(skb_copy_bytes(exh.mine(1), own_bits + (own_planes - 2) * sz / 64, (size_t)face_bytes, s));
        SKB_TRY(skb_comm_xbar(comm, s));
        ++nxbar;
        int64_t desc[1 + 3 * 2];
        int n = 0;
        if (has_lo) {  // the last two planes of the rank below
            desc[1 + 3 * n] = (int64_t)(uintptr_t)exh.field(1, rank - 1);
            desc[2 + 3 * n] = (int64_t)(uintptr_t)bits;
            desc[3 + 3 * n] = face_bytes;
            ++n;
        }
        if (has_hi) {  // the first two planes of the rank above
            desc[1 + 3 * n] = (int64_t)(uintptr_t)exh.field(0, rank + 1);
            desc[2 + 3 * n] = (int64_t)(uintptr_t)(own_bits + own_planes * sz / 64);
            desc[3 + 3 * n] = face_bytes;
            ++n;
        }
        desc[0] = n;
        if (n) SKB_TRY(skb_copy_segments(desc, s));
        if (has_lo) SKB_TRY(skb_foreach(tile_lo, s, SkbTileCountItem{bits, tile_prefix, 0}));
        if (has_hi) SKB_TRY(skb_foreach(hi_planes * sz / SKB_TILE_VOX, s, SkbTileCountItem{bits, tile_prefix, tile_lo + tile_own}));
    }
    SKB_TRY(skb_scan_u32_impl(tile_prefix, tile_prefix, ntiles, scratch, s));
    // node ranks of the plane boundaries: owned range, first / last owned plane, the halo planes next to them
    const int64_t p_own0 = (z0 - ze0) * sz, p_own1 = (z1 - ze0) * sz;
    int64_t pos[6] = {p_own0, p_own1, p_own0 + sz, p_own1 - sz, has_lo ? p_own0 - sz : p_own0, has_hi ? p_own1 + sz : p_own1};
    bool aligned = true;
    for (int i = 0; i < 6; ++i) {
        if (pos[i] > V) pos[i] = V;
        if (pos[i] != V && pos[i] % SKB_TILE_VOX) aligned = false;
    }
    int64_t Ne, pr[6];
    {
        const int32_t* ptrs[7] = {(const int32_t*)tile_prefix + ntiles};
        int n = 1;
        if (aligned)
            for (int i = 0; i < 6; ++i) ptrs[n++] = (const int32_t*)tile_prefix + (pos[i] == V ? ntiles : pos[i] / SKB_TILE_VOX);
        SKB_SFETCH(n, ptrs, vals);
        Ne = vals[0];
        for (int i = 0; i < 6; ++i) pr[i] = aligned ? vals[1 + i] : 0;
    }
    if (prm->time_sweep) SKB_TRY(skb_timer_read_ns(ev, &res->counts[SKB_SC_SWEEP_NS]));
    {
        const int64_t bytes = skb_scan_scratch_bytes(27 * Ne + ntiles + 1);
        scratch = work.take<uint8_t>(bytes);
        SKB_SNEED(work.ok);
        SKB_TRY(skb_fill_bytes(scratch, 0, (size_t)bytes, s));
    }
    const bool want_values = prm->want_values != 0;
    int64_t* lin = out.take<int64_t>(Ne);
    int64_t* coords = out.take<int64_t>(Ne * 3);
    double* values = want_values ? out.take<double>(Ne) : nullptr;
    uint32_t* nb = work.take<uint32_t>(Ne);
    uint32_t* jcount = work.take<uint32_t>(Ne + 1);
    SKB_SNEED(work.ok && out.ok);
    SKB_TRY(skb_emit_nodes(bits, tile_prefix, ntiles, 3, shape, prm->image, prm->dtype, ze0, lin, coords, values, pre256, s));
    SKB_TRY(skb_raw_neighbors(bits, 3, shape, lin, Ne, 0, nb, s));
    if (!aligned) {
        if (Ne) {
            // host positions -> device, through the pinned staging of a fill-free path: six 8-byte memsets would
            // cost more launches than one tiny kernel, so the positions travel as a kernel argument
            SkbSetI64Item set{};
            for (int i = 0; i < 6; ++i) set.v[i] = pos[i] < V ? pos[i] : 0;
            set.out = posd;
            SKB_TRY(skb_foreach(6, s, set));
            SKB_TRY(skb_rank_query(bits, pre256, posd, 6, small, s));
            const int32_t* ptrs[6] = {small, small + 1, small + 2, small + 3, small + 4, small + 5};
            SKB_SFETCH(6, ptrs, vals);
        }
        for (int i = 0; i < 6; ++i) pr[i] = pos[i] >= V ? Ne : (Ne ? vals[i] : 0);
    }
    const int64_t own0 = pr[0], own1 = pr[1], f1 = pr[2], l0 = pr[3], hl0 = pr[4], hu1 = pr[5];
    const int64_t N_k = own1 - own0;
    SKB_HOST_MARK();
    // ---- junction MST on the junction edges of the whole volume ----------------------------------------------
    bool counted = false;  // jcount holds the scanned per-node edge counts (count + scan + emit form)
    auto count_edges = [&]() -> int {
        if (counted) return 0;
        SKB_TRY(skb_junction_edge_count(bits, pre256, 3, shape, lin, nb, Ne, own0, own1, jcount, s));
        SKB_TRY(skb_scan_u32_impl(jcount, jcount, Ne, scratch, s));
        counted = true;
        return 0;
    };
    // The junction edges of every rank go to every rank.  With a capacity learnt from the previous call (the largest
    // per-rank count, plus slack: the same number on every rank) the edges are emitted straight away and ONE
    // synchronisation + exchange carries the counts and doubles as the barrier of the data; without one (first call,
    // or the capacity proved too small: every rank sees that in the exchanged counts) the counts are exchanged first.
    int64_t nej_k = -1;
    int64_t Noff[SKB_MAX_RANKS + 1], own0_of[SKB_MAX_RANKS];
    int64_t max_e = 0;
    auto exchange_counts = [&]() -> int {
        const int64_t v[3] = {N_k, nej_k, own0};
        SKB_TRY(skb_comm_allgather(comm, v, 3, all));
        ++nexch;
        Noff[0] = 0;
        max_e = 0;
        for (int k = 0; k < world; ++k) {
            Noff[k + 1] = Noff[k] + all[k * 3];
            if (all[k * 3 + 1] > max_e) max_e = all[k * 3 + 1];
            own0_of[k] = all[k * 3 + 2];
        }
        return 0;
    };
    const int32_t* nej_src = nullptr;  // where the number of this rank's edges is: the scan's total, or the append counter
    auto fetch_nej = [&]() -> int {
        const int32_t* ptrs[1] = {nej_src};
        SKB_SFETCH(1, ptrs, vals);
        nej_k = vals[0];
        return 0;
    };
    int32_t *gea = nullptr, *geb = nullptr;
    uint64_t* gew = nullptr;
    uint8_t *gek = nullptr, *tree = nullptr;
    int32_t *mpar = nullptr, *era = nullptr, *erb = nullptr;
    uint64_t* bestw = nullptr;
    int64_t nej_rows = 0;
    SkbExchange ex;
    const int64_t rb[4] = {4, 4, 8, 1};
    bool speculative = comm->hint_max_e > 0;
    const int64_t work_mark_j = work.off, xoff_mark_j = xoff;
    if (!speculative) {
        SKB_TRY(count_edges());
        nej_src = (const int32_t*)jcount + Ne;
        SKB_TRY(fetch_nej());
        SKB_TRY(exchange_counts());
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int64_t rows = speculative ? comm->hint_max_e + comm->hint_max_e / 8 + 1024 : max_e;
        if (!rows) break;
        work.off = work_mark_j;
        xoff = xoff_mark_j;
        ex.init(comm, &xoff, rows, 4, rb);
        if (!ex.ok) {
            res->exchange_need = 8 * (xoff - xoff0) + (1 << 23);
            return SKB_RUN_NEED_EXCHANGE;
        }
        nej_rows = world * ex.cap;
        gea = work.take<int32_t>(nej_rows);
        geb = work.take<int32_t>(nej_rows);
        gew = work.take<uint64_t>(nej_rows);
        gek = work.take<uint8_t>(nej_rows);
        era = work.take<int32_t>(nej_rows);
        erb = work.take<int32_t>(nej_rows);
        tree = work.take<uint8_t>(nej_rows);
        SKB_SNEED(work.ok);
        SKB_TRY(skb_fill_bytes(ex.mine(0), 0xff, (size_t)ex.cap * 4, s));  // a < 0 marks the padding rows
        bool appended = false;
        if (speculative && !height) {  // one unordered pass (the keyed Boruvka needs no edge order; table weights only)
            int r = skb_junction_append(bits, pre256, 3, shape, lin, nb, Ne, own0, own1, (int32_t*)ex.mine(0),
                                        (int32_t*)ex.mine(1), (uint8_t*)ex.mine(3), ex.cap, small + 11, s);
            if (r < 0) return r;
            appended = r == 0;
            if (appended) nej_src = small + 11;
        }
        if (!appended) {
            SKB_TRY(count_edges());
            nej_src = (const int32_t*)jcount + Ne;
            if (nej_k != 0)
                SKB_TRY(skb_junction_edge_emit_cap(bits, pre256, 3, shape, lin, nb, jcount, Ne, prm->wtab, values, height,
                                                   prm->dtype, (int32_t*)ex.mine(0), (int32_t*)ex.mine(1),
                                                   (uint8_t*)ex.mine(3), (uint64_t*)ex.mine(2), ex.cap, s));
        }
        if (speculative) {
            SKB_TRY(fetch_nej());        // the edges are complete when their count has arrived (stream order)
            SKB_TRY(exchange_counts());  // ... and every peer's are when its row of counts has
            if (max_e > ex.cap) {        // some rank's edges did not fit: every rank sees it, every rank goes again
                speculative = false;
                continue;
            }
        } else {
            SKB_TRY(skb_stream_sync(s));
            ++nsync;
            const int64_t one = 1;
            int64_t dummy[SKB_MAX_RANKS];
            SKB_TRY(skb_comm_allgather(comm, &one, 1, dummy));
            ++nexch;
        }
        break;
    }
    comm->hint_max_e = max_e;
    if (!max_e) nej_rows = 0;
    const int64_t N_total = Noff[world];
    if (N_total >= 0x7fffffffLL) return skb_fail("more than 2^31 nodes: node ids no longer fit the reference's int32 index arrays");
    const int64_t id_shift = Noff[rank] - own0;
    if (nej_rows) {
        mpar = work.take<int32_t>(N_total);
        bestw = work.take<uint64_t>(2 * N_total);
        SKB_SNEED(work.ok);
        {   // the peers' rows, in rank order (= the global (a, b) order the tie-break needs), read in place
            int64_t desc[1 + 3 * 4 * SKB_MAX_RANKS];
            void* dst[4] = {gea, geb, gew, gek};
            int n = 0;
            for (int f = 0; f < 4; ++f)
                for (int k = 0; k < world; ++k) {
                    const int64_t seg = ex.cap * rb[f];
                    desc[1 + 3 * n] = (int64_t)(uintptr_t)ex.field(f, k);
                    desc[2 + 3 * n] = (int64_t)(uintptr_t)((uint8_t*)dst[f] + k * seg);
                    desc[3 + 3 * n] = seg;
                    ++n;
                }
            desc[0] = n;
            SKB_TRY(skb_copy_segments(desc, s));
        }
        int64_t shifts[SKB_MAX_RANKS];
        for (int k = 0; k < world; ++k) shifts[k] = Noff[k] - own0_of[k];
        SKB_TRY(skb_slab_edge_shift(gea, geb, nej_rows, ex.cap, shifts, world, s));
        // Boruvka on the gathered edges (the same on every rank) and, in the same launch, the dropped edges cleared from
        // this rank's raw masks in place
        SKB_TRY(skb_mst_junctions_apply(gea, geb, gek, gew, prm->wtab, height, nej_rows, N_total, mpar, bestw, era, erb, tree,
                                        flag, nb, id_shift, Ne, s));
    }
    SKB_HOST_MARK();
    // ---- CSR + node records + start numbering (one pass) ------------------------------------------------------
    int64_t slab[13] = {own0, own1, hl0, hu1, has_lo ? hl0 : 0, has_lo ? f1 : 0, has_hi ? l0 : 0, has_hi ? hu1 : 0,
                        0, 0, 0, 0, ze0 * sz};
    uint32_t* keep = nb;
    int32_t* deg = out.take<int32_t>(Ne);
    int32_t* indptr = out.take<int32_t>(Ne + 1);
    SKB_SNEED(out.ok);
    SKB_TRY(skb_degrees_indptr(keep, Ne, deg, indptr, scratch, s));
    // graph.indices / data sized by the previous call's edge count (plus slack) when there is one: the emission runs
    // without waiting for the exact count, which then arrives with the six offsets below (one synchronisation less)
    int64_t E = -1;
    const int64_t e_cap = comm->hint_edges > 0 ? comm->hint_edges + comm->hint_edges / 16 + 4096 : 0;
    if (!e_cap) {
        const int32_t* ptrs[1] = {indptr + Ne};
        SKB_SFETCH(1, ptrs, vals);
        E = vals[0];
    }
    const int64_t out_mark_e = out.off, work_mark_e = work.off;
    int32_t* indices = out.take<int32_t>(e_cap ? e_cap : E);
    double* data = out.take<double>(e_cap ? e_cap : E);
    skb_i4* rec = work.take<skb_i4>(2 * Ne);
    uint32_t* startoff = work.take<uint32_t>(Ne + 1);
    uint32_t* is_min = work.take<uint32_t>(Ne + 1);
    SKB_SNEED(work.ok && out.ok);
    SKB_TRY(skb_csr_emit_cap(bits, pre256, 3, shape, lin, keep, indptr, Ne, prm->wtab, values, height, prm->dtype, indices, data,
                             rec, slab, 1, startoff, nullptr, nullptr, is_min, scratch, e_cap ? e_cap : 0x7fffffffLL, s,
                             N_total / world + 1));
    int64_t so[4], ei[2];
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int32_t* ptrs[7] = {(const int32_t*)startoff + own0, (const int32_t*)startoff + own1,
                                  (const int32_t*)startoff + f1,   (const int32_t*)startoff + l0,
                                  indptr + own0,                   indptr + own1,
                                  indptr + Ne};
        SKB_SFETCH(7, ptrs, vals);
        for (int i = 0; i < 4; ++i) so[i] = vals[i];
        ei[0] = vals[4];
        ei[1] = vals[5];
        E = vals[6];
        if (!e_cap || E <= e_cap || attempt) break;
        out.off = out_mark_e;  // the capacity was too small: emit again with the exact size
        work.off = work_mark_e;
        indices = out.take<int32_t>(E);
        data = out.take<double>(E);
        rec = work.take<skb_i4>(2 * Ne);
        startoff = work.take<uint32_t>(Ne + 1);
        is_min = work.take<uint32_t>(Ne + 1);
        SKB_SNEED(work.ok && out.ok);
        SKB_TRY(skb_csr_emit_cap(bits, pre256, 3, shape, lin, keep, indptr, Ne, prm->wtab, values, height, prm->dtype, indices, data,
                                 rec, slab, 1, startoff, nullptr, nullptr, is_min, scratch, 0x7fffffffLL, s, N_total / world + 1));
    }
    comm->hint_edges = E;
    const int64_t S = so[1] - so[0], SF = so[2] - so[0], SL0 = so[3] - so[0];
    const int64_t E_k = ei[1] - ei[0];
    int64_t Eoff[SKB_MAX_RANKS + 1], Soff[SKB_MAX_RANKS + 1], cnt_sf[SKB_MAX_RANKS], cnt_sl0[SKB_MAX_RANKS];
    int64_t nblk[SKB_MAX_RANKS], max_blk = 0, n_table_records = 0;
    {
        const int64_t v[5] = {N_k, E_k, S, SF, SL0};
        SKB_TRY(skb_comm_allgather(comm, v, 5, all));
        ++nexch;
        Eoff[0] = Soff[0] = 0;
        for (int k = 0; k < world; ++k) {
            Eoff[k + 1] = Eoff[k] + all[k * 5 + 1];
            Soff[k + 1] = Soff[k] + all[k * 5 + 2];
            cnt_sf[k] = all[k * 5 + 3];
            cnt_sl0[k] = all[k * 5 + 4];
            nblk[k] = all[k * 5 + 3] + (all[k * 5 + 2] - all[k * 5 + 4]);
            if (nblk[k] > max_blk) max_blk = nblk[k];
            n_table_records += nblk[k];
        }
    }
    const int64_t S_total = Soff[world];
    const int64_t start_lo = Soff[rank], start_hi = Soff[rank + 1];
    slab[8] = id_shift;
    slab[9] = start_lo - so[0];
    slab[10] = start_lo;
    slab[11] = start_hi;
    SKB_HOST_MARK();
    // ---- phase 0: lists between terminals, first inside the slab ---------------------------------------------
    const int64_t S1 = S > 0 ? S : 1, Ne1 = Ne > 0 ? Ne : 1;
    int32_t* su = work.take<int32_t>(S1);
    int32_t* se = work.take<int32_t>(S1);
    skb_i4* buf0 = work.take<skb_i4>(S1);
    skb_i4* buf1 = work.take<skb_i4>(S1);
    SkbAux* aux0 = work.take<SkbAux>(S1);
    SkbAux* aux1 = work.take<SkbAux>(S1);
    int32_t* stamp = work.take<int32_t>(Ne1);
    SKB_SNEED(work.ok);
    SKB_TRY(skb_fill_bytes(stamp, 0xff, (size_t)Ne1 * 4, s));
    SKB_TRY(skb_path_walk(indptr, indices, data, values, lin, rec, startoff, slab, Ne, S, su, se, buf0, aux0, stamp, s));
    skb_i4* ranked = buf0;
    SkbAux* raux = aux0;
    {
        int r = skb_path_rank(buf0, buf1, aux0, aux1, S, start_lo, start_hi, flag, s);
        if (r < 0) return r;
        if (r & 1) {
            ranked = buf1;
            raux = aux1;
        }
    }
    SKB_HOST_MARK();
    // ---- inter-slab table ------------------------------------------------------------------------------------------
    int32_t* cross = small + 8;  // != 0: a list of the inter-slab table never ends (a cycle crossing slabs)
    SKB_TRY(skb_fill_bytes(cross, 0, 4, s));
    if (world > 1 && max_blk > 0) {
        SkbExchange ex;
        const int64_t rb[2] = {16, (int64_t)sizeof(SkbAux)};
        ex.init(comm, &xoff, max_blk, 2, rb);
        if (!ex.ok) {
            res->exchange_need = 8 * (xoff - xoff0) + (1 << 23);
            return SKB_RUN_NEED_EXCHANGE;
        }
        const int64_t n_table = world * ex.cap;
        skb_i4* T = work.take<skb_i4>(n_table);
        skb_i4* T2 = work.take<skb_i4>(n_table);
        SkbAux* TA = work.take<SkbAux>(n_table);
        SkbAux* TA2 = work.take<SkbAux>(n_table);
        SKB_SNEED(work.ok);
        SKB_TRY(skb_slab_export(ranked, raux, SF, SL0, nblk[rank], ex.cap, start_lo, start_hi, (skb_i4*)ex.mine(0),
                                ex.mine(1), s));
        SKB_TRY(skb_comm_xbar(comm, s));
        ++nxbar;
        {
            int64_t desc[1 + 3 * 2 * SKB_MAX_RANKS];
            void* dst[2] = {T, TA};
            int n = 0;
            for (int f = 0; f < 2; ++f)
                for (int k = 0; k < world; ++k) {
                    const int64_t seg = ex.cap * rb[f];
                    desc[1 + 3 * n] = (int64_t)(uintptr_t)ex.field(f, k);
                    desc[2 + 3 * n] = (int64_t)(uintptr_t)((uint8_t*)dst[f] + k * seg);
                    desc[3 + 3 * n] = seg;
                    ++n;
                }
            desc[0] = n;
            SKB_TRY(skb_copy_segments(desc, s));
        }
        int64_t td[1 + 4 * SKB_MAX_RANKS + 1];
        {
            int64_t* p = td;
            *p++ = world;
            for (int k = 0; k <= world; ++k) *p++ = Soff[k];
            for (int k = 0; k < world; ++k) *p++ = cnt_sf[k];
            for (int k = 0; k < world; ++k) *p++ = cnt_sl0[k];
            for (int k = 0; k < world; ++k) *p++ = k * ex.cap;
        }
        skb_i4* Tf = T;
        SkbAux* TAf = TA;
        int r2 = skb_slab_table_rank(T, T2, TA, TA2, n_table, td, flag, s);
        if (r2 < 0) return r2;
        if (r2 & 1) {
            Tf = T2;
            TAf = TA2;
        }
        SKB_TRY(skb_slab_apply(ranked, raux, Tf, TAf, S, td, start_lo, start_hi, s));
        SKB_TRY(skb_foreach(n_table, s, SkbAnyUnfinishedItem{Tf, cross}));
    }
    SKB_HOST_MARK();
    // ---- covered chain voxels, heads of the regular paths -----------------------------------------------------------
    uint8_t* covered = work.take<uint8_t>(Ne1);
    uint32_t* n_unc = work.take<uint32_t>(64);
    uint32_t* sel = work.take<uint32_t>(S + 1);
    SKB_SNEED(work.ok);
    SKB_TRY(skb_slab_covered(rec, startoff, stamp, ranked, Ne, own0, own1, covered, n_unc, s));
    SKB_TRY(skb_path_select(su, rec, ranked, S, 0, start_lo, sel, scratch, s));
    SKB_TRY(skb_foreach(1, s, SkbSumSlotsItem{n_unc, 64, small + 9}));
    int64_t P0, n_uncovered, cc;
    {
        const int32_t* ptrs[3] = {(const int32_t*)sel + S, small + 9, cross};
        SKB_SFETCH(3, ptrs, vals);
        P0 = vals[0];
        n_uncovered = vals[1];
        cc = vals[2] != 0;
    }
    SKB_HOST_MARK();
    // ---- phase 1: junction-free cycles -----------------------------------------------------------------------------------
    // All ranks agree first whether some ring touches a slab boundary (a list of the inter-slab table that never
    // ends).  If none does, every ring lies inside one slab and the parallel machinery of phase 0 runs again on
    // the uncovered voxels.  Otherwise (rare) every rank exports its uncovered chain voxels, all ranks gather them
    // and the rank that owns a ring's minimum node emits the whole ring (SkbRing*Item).
    int64_t Pc = 0, Sb = 0;
    int32_t *sub = nullptr, *seb = nullptr;
    skb_i4* rankedb = nullptr;
    SkbAux* rauxb = nullptr;
    uint32_t* selb = nullptr;
    bool ring_mode = false;
    int64_t ring_m = 0, unc_of[SKB_MAX_RANKS], max_unc = 0;
    SkbRingRec* ring_rec = nullptr;
    uint32_t *ring_size = nullptr, *ring_oidx = nullptr;
    const int64_t node_lo = Noff[rank], node_hi = Noff[rank + 1];
    {
        const int64_t v[2] = {cc, n_uncovered};
        SKB_TRY(skb_comm_allgather(comm, v, 2, all));
        ++nexch;
        for (int k = 0; k < world; ++k) {
            if (all[2 * k]) ring_mode = true;
            unc_of[k] = all[2 * k + 1];
            ring_m += unc_of[k];
            if (unc_of[k] > max_unc) max_unc = unc_of[k];
        }
    }
    if (ring_mode && ring_m > 0) {
        SkbExchange exq;
        const int64_t rbq[1] = {(int64_t)sizeof(SkbRingRec)};
        exq.init(comm, &xoff, max_unc, 1, rbq);
        if (!exq.ok) {
            res->exchange_need = 8 * (xoff - xoff0) + (1 << 23);
            return SKB_RUN_NEED_EXCHANGE;
        }
        uint32_t* rflag = work.take<uint32_t>(Ne + 1);
        ring_rec = work.take<SkbRingRec>(ring_m);
        int32_t* rpar = work.take<int32_t>(ring_m);
        ring_size = work.take<uint32_t>(ring_m);
        ring_oidx = work.take<uint32_t>(ring_m + 1);
        SKB_SNEED(work.ok);
        SKB_TRY(skb_foreach(Ne, s, SkbRingFlagItem{rec, covered, (int32_t)own0, (int32_t)own1, rflag}));
        SKB_TRY(skb_scan_u32_impl(rflag, rflag, Ne, scratch, s));
        SKB_TRY(skb_foreach(Ne, s, SkbRingExportItem{rec, rflag, values, (int32_t)id_shift, exq.cap, (SkbRingRec*)exq.mine(0)}));
        SKB_TRY(skb_stream_sync(s));
        ++nsync;
        {
            const int64_t one = 1;
            SKB_TRY(skb_comm_allgather(comm, &one, 1, all));
            ++nexch;
        }
        {   // exactly unc_of[k] records of rank k, in rank order = ascending global id
            int64_t desc[1 + 3 * SKB_MAX_RANKS];
            int n = 0;
            int64_t at = 0;
            for (int k = 0; k < world; ++k) {
                if (!unc_of[k]) continue;
                desc[1 + 3 * n] = (int64_t)(uintptr_t)exq.field(0, k);
                desc[2 + 3 * n] = (int64_t)(uintptr_t)(ring_rec + at);
                desc[3 + 3 * n] = unc_of[k] * (int64_t)sizeof(SkbRingRec);
                at += unc_of[k];
                ++n;
            }
            desc[0] = n;
            SKB_TRY(skb_copy_segments(desc, s));
        }
        int32_t* bad = small + 10;
        SKB_TRY(skb_fill_bytes(bad, 0, 4, s));
        SKB_TRY(skb_fill_bytes(ring_size, 0, (size_t)ring_m * 4, s));
        SKB_TRY(skb_foreach(ring_m, s, SkbIotaItem{rpar}));
        SKB_TRY(skb_foreach(ring_m, s, SkbRingHookItem{ring_rec, (int32_t)ring_m, rpar, bad}));
        SKB_TRY(skb_foreach(ring_m, s, SkbRingSizeItem{ring_rec, rpar, (int32_t)node_lo, (int32_t)node_hi, ring_size, ring_oidx}));
        SKB_TRY(skb_scan_u32_impl(ring_oidx, ring_oidx, ring_m, scratch, s));
        const int32_t* ptrs[2] = {(const int32_t*)ring_oidx + ring_m, bad};
        SKB_SFETCH(2, ptrs, vals);
        Pc = vals[0];
        if (vals[1]) return skb_fail("slab mode: the uncovered chain voxels do not form closed rings");
    }
    // rings inside this rank's slab: few uncovered chain voxels (the usual case) are one thread block's work -- heads now,
    // entries once the exchange below has told where this rank's paths start
    bool small_cycles = false;
    int32_t *clist = nullptr, *t_start = nullptr, *t_pmin = nullptr;
    uint32_t* t_plen = nullptr;
    SkbAux* t_aux = nullptr;
    if (!ring_mode && n_uncovered > 0 && n_uncovered <= SKB_SMALL_CYCLE_LIMIT) {
        const int64_t pc_cap = n_uncovered / 3 + 1;  // a ring has at least three voxels
        clist = work.take<int32_t>(n_uncovered);
        t_start = work.take<int32_t>(pc_cap);
        t_plen = work.take<uint32_t>(pc_cap);
        t_pmin = work.take<int32_t>(pc_cap);
        t_aux = work.take<SkbAux>(pc_cap);
        SKB_SNEED(work.ok);
        int r = skb_cycles_small_heads(rec, covered, Ne, n_uncovered, own0, own1, values, lin, ze0 * sz, id_shift, t_start,
                                       t_plen, t_pmin, is_min, t_aux, clist, small + 12, s);
        if (r < 0) return r;
        if (r == 0) {
            const int32_t* ptrs[2] = {small + 12, small + 14};
            SKB_SFETCH(2, ptrs, vals);
            if (vals[1]) return skb_fail("slab mode: the uncovered chain voxels do not form closed rings");
            Pc = vals[0];
            small_cycles = true;
        }
    }
    if (!ring_mode && n_uncovered > 0 && !small_cycles) {
        int32_t* cpar = work.take<int32_t>(Ne1);
        uint32_t* startoff_b = work.take<uint32_t>(Ne + 1);
        SKB_SNEED(work.ok);
        SKB_TRY(skb_cycle_roots(rec, covered, Ne, cpar, is_min, s));
        SKB_TRY(skb_path_start_count(rec, covered, slab, Ne, startoff_b, scratch, s));
        const int32_t* ptrs[2] = {(const int32_t*)startoff_b + own0, (const int32_t*)startoff_b + own1};
        SKB_SFETCH(2, ptrs, vals);
        const int64_t sob0 = vals[0];
        Sb = vals[1] - vals[0];
        if (Sb) {
            int64_t slab_b[13];
            for (int i = 0; i < 13; ++i) slab_b[i] = slab[i];
            slab_b[9] = -sob0;
            slab_b[10] = 0;
            slab_b[11] = Sb;
            sub = work.take<int32_t>(Sb);
            seb = work.take<int32_t>(Sb);
            skb_i4* bb0 = work.take<skb_i4>(Sb);
            skb_i4* bb1 = work.take<skb_i4>(Sb);
            SkbAux* ab0 = work.take<SkbAux>(Sb);
            SkbAux* ab1 = work.take<SkbAux>(Sb);
            selb = work.take<uint32_t>(Sb + 1);
            SKB_SNEED(work.ok);
            SKB_TRY(skb_path_walk(indptr, indices, data, values, lin, rec, startoff_b, slab_b, Ne, Sb, sub, seb, bb0, ab0,
                                  nullptr, s));
            int rb_ = skb_path_rank(bb0, bb1, ab0, ab1, Sb, 0, Sb, flag, s);
            if (rb_ < 0) return rb_;
            rankedb = (rb_ & 1) ? bb1 : bb0;
            rauxb = (rb_ & 1) ? ab1 : ab0;
            SKB_TRY(skb_path_select(sub, rec, rankedb, Sb, 1, 0, selb, scratch, s));
            const int32_t* p1[1] = {(const int32_t*)selb + Sb};
            SKB_SFETCH(1, p1, vals);
            Pc = vals[0];
        }
    }
    const int64_t Pk = P0 + Pc, Pk1 = Pk > 0 ? Pk : 1;
    int32_t* start_node = out.take<int32_t>(Pk1);
    uint32_t* plen = out.take<uint32_t>(Pk + 1);
    int32_t* pmin = work.take<int32_t>(Pk1);
    int32_t* head_end = work.take<int32_t>(Pk1);
    SkbAux* head_aux = work.take<SkbAux>(Pk1);
    int32_t* pindptr = work.take<int32_t>(Pk + 1);
    skb_i4* einfo_b = (Pc && !ring_mode && !small_cycles) ? work.take<skb_i4>(Sb) : nullptr;
    SKB_SNEED(work.ok && out.ok);
    SKB_TRY(skb_path_heads(su, rec, ranked, raux, sel, S, 0, 0, start_lo, id_shift, start_node, plen, pmin, nullptr,
                           head_end, head_aux, s));
    if (Pc && ring_mode)
        SKB_TRY(skb_foreach(ring_m, s, SkbRingHeadItem{ring_rec, ring_size, ring_oidx, (int32_t)P0, (int32_t)id_shift,
                                                      start_node, plen, pmin, is_min}));
    else if (Pc && small_cycles)
        SKB_TRY(skb_foreach(Pc, s, SkbCopyCycleHeadsItem{t_start, t_plen, t_pmin, t_aux, (int32_t)P0, start_node, plen, pmin,
                                                        head_aux}));
    else if (Pc)
        SKB_TRY(skb_path_heads(sub, rec, rankedb, rauxb, selb, Sb, 1, P0, 0, id_shift, start_node, plen, pmin, einfo_b,
                               head_end, head_aux, s));
    SKB_TRY(skb_scan_u32_impl(plen, (uint32_t*)pindptr, Pk, scratch, s));
    int64_t L0, Lc;
    {
        const int32_t* ptrs[2] = {pindptr + P0, pindptr + Pk};
        SKB_SFETCH(2, ptrs, vals);
        L0 = vals[0];
        Lc = vals[1] - vals[0];
    }
    int64_t P0off[SKB_MAX_RANKS + 1], L0off[SKB_MAX_RANKS + 1], Pcoff[SKB_MAX_RANKS + 1], Lcoff[SKB_MAX_RANKS + 1];
    int64_t maxP = 0;
    bool any_cc = false;
    {
        const int64_t v[5] = {P0, L0, Pc, Lc, ring_mode ? 0 : cc};
        SKB_TRY(skb_comm_allgather(comm, v, 5, all));
        ++nexch;
        P0off[0] = L0off[0] = Pcoff[0] = Lcoff[0] = 0;
        for (int k = 0; k < world; ++k) {
            P0off[k + 1] = P0off[k] + all[k * 5];
            L0off[k + 1] = L0off[k] + all[k * 5 + 1];
            Pcoff[k + 1] = Pcoff[k] + all[k * 5 + 2];
            Lcoff[k + 1] = Lcoff[k] + all[k * 5 + 3];
            if (all[k * 5] > maxP) maxP = all[k * 5];
            if (all[k * 5 + 4]) any_cc = true;
        }
    }
    if (any_cc) return skb_fail("a junction-free cycle crosses a slab boundary (not supported in slab mode)");
    const int64_t P0_total = P0off[world], L0_total = L0off[world], Pc_total = Pcoff[world], Lc_total = Lcoff[world];
    const int64_t P_total = P0_total + Pc_total, L_total = L0_total + Lc_total;
    res->counts[SKB_SC_NODES_EXT] = Ne;
    res->counts[SKB_SC_EDGES_EXT] = E;
    res->counts[SKB_SC_OWN0] = own0;
    res->counts[SKB_SC_OWN1] = own1;
    res->counts[SKB_SC_ID_SHIFT] = id_shift;
    res->counts[SKB_SC_NODE_OFFSET] = Noff[rank];
    res->counts[SKB_SC_EDGE_OFFSET] = Eoff[rank];
    res->counts[SKB_SC_EDGES_OWNED] = E_k;
    res->counts[SKB_SC_NODES_TOTAL] = N_total;
    res->counts[SKB_SC_EDGES_TOTAL] = Eoff[world];
    res->counts[SKB_SC_REGULAR] = P0;
    res->counts[SKB_SC_CYCLES] = Pc;
    res->counts[SKB_SC_PATHS_TOTAL] = P_total;
    res->counts[SKB_SC_ENTRIES_TOTAL] = L_total;
    res->counts[SKB_SC_REGULAR_TOTAL] = P0_total;
    res->counts[SKB_SC_REGULAR_OFFSET] = P0off[rank];
    res->counts[SKB_SC_CYCLE_OFFSET] = P0_total + Pcoff[rank];
    res->counts[SKB_SC_ENTRY_OFF_REGULAR] = L0off[rank];
    res->counts[SKB_SC_ENTRY_OFF_CYCLES] = L0_total + Lcoff[rank];
    res->counts[SKB_SC_STARTS] = S;
    res->counts[SKB_SC_TABLE_RECORDS] = n_table_records;
    res->off[SKB_SO_LIN] = out_off(lin);
    res->off[SKB_SO_COORDS] = out_off(coords);
    res->off[SKB_SO_VALUES] = out_off(values);
    res->off[SKB_SO_DEGREES] = out_off(deg);
    res->off[SKB_SO_INDPTR] = out_off(indptr);
    res->off[SKB_SO_INDICES] = out_off(indices);
    res->off[SKB_SO_DATA] = out_off(data);
    if (P_total == 0) {  // the caller raises the reference's ValueError (csr.py:442)
        res->counts[SKB_SC_SYNCS] = nsync;
        res->counts[SKB_SC_EXCHANGES] = nexch;
        return 0;
    }
    SKB_HOST_MARK();
    // ---- path records to every rank --------------------------------------------------------------------------------------
    SkbExchange exr;
    {
        const int64_t rb[2] = {16, 16};
        exr.init(comm, &xoff, maxP, 2, rb);
    }
    SkbExchange exs;  // skeleton-id contributions: one int32 per gathered path record and rank
    {
        const int64_t rb[1] = {4};
        exs.init(comm, &xoff, world * exr.cap, 1, rb);
    }
    if (!exr.ok || !exs.ok) {
        res->exchange_need = 8 * (xoff - xoff0) + (1 << 23);
        return SKB_RUN_NEED_EXCHANGE;
    }
    const int64_t nrec = world * exr.cap;
    const int64_t S_total1 = S_total > 0 ? S_total : 1;
    skb_i4* heads_all = work.take<skb_i4>(nrec);
    skb_i4* links_all = work.take<skb_i4>(nrec);
    skb_i4* einfo = work.take<skb_i4>(S_total1);
    int32_t* pidx = out.take<int32_t>(L_total);
    double* pdata = out.take<double>(L_total);
    double* pw = out.take<double>(L_total);
    int32_t* kpar = work.take<int32_t>(S_total1);
    int32_t* kmin = work.take<int32_t>(S_total1);
    SKB_SNEED(work.ok && out.ok);
    SKB_TRY(skb_slab_head_records(head_end, plen, pindptr, pmin, start_node, startoff, head_aux, slab, Ne, P0, exr.cap,
                                  P0off[rank], L0off[rank], (skb_i4*)exr.mine(0), (skb_i4*)exr.mine(1), s));
    SKB_TRY(skb_comm_xbar(comm, s));
    ++nxbar;
    {
        int64_t desc[1 + 3 * 2 * SKB_MAX_RANKS];
        void* dst[2] = {heads_all, links_all};
        int n = 0;
        for (int f = 0; f < 2; ++f)
            for (int k = 0; k < world; ++k) {
                const int64_t seg = exr.cap * 16;
                desc[1 + 3 * n] = (int64_t)(uintptr_t)exr.field(f, k);
                desc[2 + 3 * n] = (int64_t)(uintptr_t)((uint8_t*)dst[f] + k * seg);
                desc[3 + 3 * n] = seg;
                ++n;
            }
        desc[0] = n;
        SKB_TRY(skb_copy_segments(desc, s));
    }
    SKB_TRY(skb_slab_einfo(heads_all, nrec, S_total, einfo, s));
    SKB_HOST_MARK();
    // ---- path entries: every rank writes the entries of its own voxels at their global positions ----------------------------
    SKB_TRY(skb_fill_bytes(pidx, 0, (size_t)L_total * 4, s));
    SKB_TRY(skb_fill_bytes(pdata, 0, (size_t)L_total * 8, s));
    SKB_TRY(skb_fill_bytes(pw, 0, (size_t)L_total * 8, s));
    SKB_TRY(skb_path_emit(indices, data, rec, su, se, ranked, einfo, nullptr, values, nullptr, id_shift, S, pidx, pdata, pw, s));
    if (Pc) {
        int32_t* gptr = work.take<int32_t>(Pk + 1);
        SKB_SNEED(work.ok);
        SKB_TRY(skb_foreach(Pk + 1, s, SkbAddOffsetItem{pindptr, gptr, (int32_t)((L0_total + Lcoff[rank]) - L0)}));
        if (ring_mode)
            SKB_TRY(skb_foreach(ring_m, s, SkbRingWalkItem{ring_rec, (int32_t)ring_m, ring_oidx, (int32_t)P0, (int32_t)id_shift,
                                                          gptr, lin, ze0 * sz, pidx, pdata, pw, head_aux}));
        else if (small_cycles)
            SKB_TRY(skb_cycles_small_entries(rec, n_uncovered, values, id_shift, P0, gptr, pidx, pdata, pw, clist, small + 12, s));
        else
            SKB_TRY(skb_path_emit(indices, data, rec, sub, seb, rankedb, einfo_b, gptr, values, nullptr, id_shift, Sb, pidx,
                                  pdata, pw, s));
    }
    SKB_HOST_MARK();
    // ---- skeleton ids ------------------------------------------------------------------------------------------------------------
    SKB_TRY(skb_slab_components(links_all, nrec, S_total, kpar, kmin, node_lo, node_hi, id_shift, is_min, Ne, scratch, s));
    // the last allocations come before the last exchange through the host: a rank that gets past it knows that no
    // rank will give the call up (the barrier that follows is on the device and has nobody to report to)
    int32_t* skel = work.take<int32_t>(Pk1);
    int32_t* skel_reg = work.take<int32_t>(P0 > 0 ? P0 : 1);
    double* dist = out.take<double>(Pk1);
    double* mean = out.take<double>(Pk1);
    double* stdev = out.take<double>(Pk1);
    int32_t* t32 = out.take<int32_t>(3 * Pk1);
    int64_t* t64 = out.take<int64_t>(7 * Pk1);
    double* tf = out.take<double>((2 * (3 + height) + 1) * Pk1);
    SKB_SNEED(work.ok && out.ok);
    int64_t C_k;
    {
        const int32_t* ptrs[2] = {(const int32_t*)is_min + own0, (const int32_t*)is_min + own1};
        SKB_SFETCH(2, ptrs, vals);
        C_k = vals[1] - vals[0];
    }
    int64_t Coff[SKB_MAX_RANKS + 1];
    {
        SKB_TRY(skb_comm_allgather(comm, &C_k, 1, all));
        ++nexch;
        Coff[0] = 0;
        for (int k = 0; k < world; ++k) Coff[k + 1] = Coff[k] + all[k];
    }
    SKB_TRY(skb_slab_skeleton_ids(links_all, nrec, kpar, kmin, is_min, node_lo, node_hi, id_shift, Coff[rank], own0,
                                  (int32_t*)exs.mine(0), s));
    SKB_TRY(skb_comm_xbar(comm, s));
    ++nxbar;
    if (P0) {  // sum of every rank's contribution to the ids of the paths headed here, read in place
        SkbPeerSumItem ps{};
        for (int k = 0; k < world; ++k) ps.src[k] = (const int32_t*)exs.field(0, k) + rank * exr.cap;
        ps.n_src = world;
        ps.dst = skel_reg;
        SKB_TRY(skb_foreach(P0, s, ps));
    }
    SKB_TRY(skb_foreach(Pk, s, SkbSlabSkelFinalItem{skel_reg, 0, (int32_t)P0, start_node, (int32_t)id_shift,
                                                   (int32_t)Coff[rank], (int32_t)own0, is_min, skel}));
    SKB_HOST_MARK();
    // ---- statistics and branch table of the paths headed here -----------------------------------------------------------------------
    SKB_TRY(skb_slab_path_stats(plen, head_aux, Pk, dist, mean, stdev, s));
    if (Pk)
        SKB_TRY(skb_branch_table(Pk, 3, height, prm->spacing_is_int, prm->spacing_f, prm->spacing_i, nullptr, nullptr, indptr, skel,
                                 coords, values, head_aux, start_node, id_shift, prm->global_shape, t32, t64, tf, s));
    res->off[SKB_SO_PLEN] = out_off(plen);
    res->off[SKB_SO_PIDX] = out_off(pidx);
    res->off[SKB_SO_PDATA] = out_off(pdata);
    res->off[SKB_SO_PW] = out_off(pw);
    res->off[SKB_SO_DIST] = out_off(dist);
    res->off[SKB_SO_MEAN] = out_off(mean);
    res->off[SKB_SO_STDEV] = out_off(stdev);
    res->off[SKB_SO_T32] = out_off(t32);
    res->off[SKB_SO_T64] = out_off(t64);
    res->off[SKB_SO_TF] = out_off(tf);
    res->off[SKB_SO_START_NODE] = out_off(start_node);
    SKB_HOST_MARK();
    if (dev_clock) {  // dev_ns[i] = device time between boundary i and boundary i + 1
        int64_t ns[16];
        SKB_TRY(skb_stage_read(dev_marked, 16, ns));
        for (int i = 0; i + 1 < 16; ++i) res->dev_ns[i] = ns[i + 1];
    }
    res->counts[SKB_SC_SYNCS] = nsync;
    res->counts[SKB_SC_EXCHANGES] = nexch;
    res->counts[SKB_SC_DEVICE_BARRIERS] = nxbar;
    res->work_need = work.need;
    res->out_need = out.need;
    return 0;
}
