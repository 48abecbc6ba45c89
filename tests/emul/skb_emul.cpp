// TEST INFRASTRUCTURE -- host emulation of libskan_b200's C-ABI for CPU-only unit tests.
//
// Compiles the SAME per-item device functions (skan_b200/csrc/skb_items.cuh) and the SAME
// C-ABI bodies (skan_b200/csrc/skb_api.inl) with g++, running every "one thread per item"
// kernel as a serial loop.  This lets `pytest -m "not gpu"` exercise the index arithmetic of
// the node/edge/path stages and the python host logic in a container without a GPU.  The
// three block-cooperative CUDA kernels (mask sweep, node emission, scans) are replaced by
// plain serial restatements below and are therefore NOT covered here -- the `-m gpu` tests
// cover them.  The skan_b200 package never loads this library: it is built and injected only
// by tests/ (tests/emul_util.py).
//
//   g++ -O2 -ffp-contract=off -shared -fPIC -I skan_b200/csrc -o tests/emul/_build/libskb_emul.so tests/emul/skb_emul.cpp
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "skb_items.cuh"

typedef void* skb_stream_t;

static thread_local char g_skb_err[512] = "";

static int skb_fail(const char* what) {
    snprintf(g_skb_err, sizeof(g_skb_err), "%s", what);
    return -2;
}

// abort word of the device-side barriers and the "guarded launch" flag, as in the CUDA build
#define SKB_POISONED 3
static int g_skb_poison = 0;
static thread_local int g_skb_guard = 0;
#define SKB_GUARDED_OUT() (g_skb_guard && g_skb_poison)

template <typename F>
static int skb_foreach(int64_t n, skb_stream_t, F f) {
    if (SKB_GUARDED_OUT()) return 0;
    for (int64_t i = 0; i < n; ++i) f(i);
    return 0;
}

// stepwise items (the chain walks): the CUDA build refills idle lanes from a queue; here the same begin / step /
// finish functions run one item after the other
template <typename F>
static int skb_foreach_steps(int64_t n, skb_stream_t, F f) {
    if (SKB_GUARDED_OUT()) return 0;
    for (int64_t i = 0; i < n; ++i) {
        typename F::State st;
        if (!f.begin(i, st)) continue;
        while (f.step(st)) {
        }
        f.finish(i, st);
    }
    return 0;
}

template <typename F>
static int skb_foreach_unmarked(int64_t n, const uint8_t* marks, skb_stream_t, F f) {
    if (SKB_GUARDED_OUT()) return 0;
    for (int64_t i = 0; i < n; ++i)
        if (!marks[i]) f(i);
    return 0;
}

static int skb_fill_bytes(void* p, int byte, size_t nbytes, skb_stream_t) {
    memset(p, byte, nbytes);
    return 0;
}

static int skb_copy_bytes(void* dst, const void* src, size_t nbytes, skb_stream_t) {
    memcpy(dst, src, nbytes);
    return 0;
}

static int skb_poisoned();
static int skb_fetch_i32s(const int32_t* const* ptrs, int n, int64_t* out, skb_stream_t) {
    for (int i = 0; i < n; ++i) out[i] = *ptrs[i];
    return skb_poisoned();
}
// two-part read-back: the values are taken when the request is posted (on the device the mail kernel runs in
// stream order, i.e. before anything enqueued after the post), and handed over by the wait
static thread_local int64_t g_skb_posted[16];
static int skb_fetch_post(const int32_t* const* ptrs, int n, skb_stream_t) {
    for (int i = 0; i < n; ++i) g_skb_posted[i] = *ptrs[i];
    return 0;
}
static int skb_fetch_wait(int n, int64_t* out, skb_stream_t) {
    for (int i = 0; i < n; ++i) out[i] = g_skb_posted[i];
    return skb_poisoned();
}
static int skb_stream_sync(skb_stream_t) { return 0; }

// exchange arena of the Z-slab mode: POSIX shared memory stands in for CUDA IPC device memory
#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
static void skb_xname(char* out, size_t n, const char* name, int rank) { snprintf(out, n, "/skbx_%s_%d", name, rank); }
static int skb_xmap(const char* shm, int64_t bytes, bool create, void** ptr) {
    int fd = shm_open(shm, create ? (O_CREAT | O_RDWR) : O_RDWR, 0600);
    if (fd < 0) return skb_fail("shm_open of an exchange arena failed");
    if (create && ftruncate(fd, bytes) != 0) { close(fd); return skb_fail("ftruncate of an exchange arena failed"); }
    void* p = mmap(nullptr, (size_t)bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return skb_fail("mmap of an exchange arena failed");
    *ptr = p;
    return 0;
}
static int skb_xalloc(const char* name, int rank, int64_t bytes, void** ptr, uint8_t handle[64]) {
    memset(handle, 0, 64);
    skb_xname((char*)handle, 64, name, rank);
    return skb_xmap((const char*)handle, bytes, true, ptr);
}
static int skb_xopen(const uint8_t handle[64], int64_t bytes, void** ptr) { return skb_xmap((const char*)handle, bytes, false, ptr); }
static int skb_xclose(void* ptr, int64_t bytes) { munmap(ptr, (size_t)bytes); return 0; }
static int skb_xfree(void* ptr, int64_t bytes, const char*, int) { munmap(ptr, (size_t)bytes); return 0; }
static int skb_xunlink(const char* name, int rank) {
    char shm[64];
    skb_xname(shm, sizeof(shm), name, rank);
    shm_unlink(shm);
    return 0;
}

// device-side barrier between the ranks (skb_xbar_kernel of the CUDA build): the same flag protocol, run by the host
// thread on the shared-memory arenas
#include <sched.h>
#include <time.h>
static int skb_xbar_launch(void* const* xbase, int rank, int world, int64_t flag_off, uint64_t seq, int poison, skb_stream_t) {
    const bool dead = poison || g_skb_poison;
    for (int k = 0; k < world; ++k) {
        uint64_t* slot = (uint64_t*)((uint8_t*)xbase[k] + flag_off) + rank;
        __atomic_store_n(slot, dead ? (seq | 0xffffffffull) : seq, __ATOMIC_RELEASE);
    }
    if (dead) return 0;
    const uint64_t* mine = (const uint64_t*)((uint8_t*)xbase[rank] + flag_off);
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < world; ++k) {
        for (uint64_t spins = 1;; ++spins) {
            const uint64_t v = __atomic_load_n(mine + k, __ATOMIC_ACQUIRE);
            if (v >= seq) {
                if ((uint32_t)v == 0xffffffffu) g_skb_poison = 1;
                break;
            }
            if ((spins & 1023) == 0) {
                timespec t1;
                clock_gettime(CLOCK_MONOTONIC, &t1);
                if (t1.tv_sec - t0.tv_sec > 60) {
                    g_skb_poison = 1;
                    break;
                }
                sched_yield();
            }
        }
        if (g_skb_poison) break;
    }
    return 0;
}
static int skb_poison_reset() {
    g_skb_poison = 0;
    return 0;
}
static int skb_poisoned() {  // every read-back reports (and clears) the abort word
    if (!g_skb_poison) return 0;
    g_skb_poison = 0;
    snprintf(g_skb_err, sizeof(g_skb_err), "a device-side barrier between the ranks was abandoned");
    return SKB_POISONED;
}

static int skb_stage_mark(int, skb_stream_t) { return 0; }
static int skb_stage_read(const uint8_t*, int n, int64_t* ns) {
    for (int i = 0; i < n; ++i) ns[i] = 0;
    return 0;
}
static int skb_timer_start(void** handle, skb_stream_t) { *handle = nullptr; return 0; }
static int skb_timer_stop(void*, skb_stream_t) { return 0; }
static int skb_timer_read_ns(void*, int64_t* ns) { *ns = 0; return 0; }

static int skb_fetch_i32(const int32_t* p, int32_t* host_out, skb_stream_t) {
    *host_out = *p;
    return 0;
}

template <typename T>
static int skb_scan_host(const T* in, T* out, int64_t n) {
    T acc = 0;
    for (int64_t i = 0; i < n; ++i) {
        T v = in[i];
        out[i] = acc;
        acc += v;
    }
    out[n] = acc;
    return 0;
}
static int skb_scan_u32_impl(const uint32_t* in, uint32_t* out, int64_t n, void*, skb_stream_t) {
    if (SKB_GUARDED_OUT()) return 0;
    return skb_scan_host<uint32_t>(in, out, n);
}
static int skb_scan_u32_dev(const uint32_t* in, uint32_t* out, int64_t n, const int32_t* n_ptr, uint32_t* total_out, void*,
                            skb_stream_t) {
    if (n_ptr && *n_ptr < n) n = *n_ptr < 0 ? 0 : *n_ptr;
    skb_scan_host<uint32_t>(in, out, n);
    if (total_out) *total_out = out[n];
    return 0;
}

#define SKB_TILE_VOX 16384
#define SKB_TILE_WORDS 256

// options of skb_set_option that reach shared code
static int g_skb_split_bits = 0;  // adaptive, like the CUDA build
static uint32_t skb_option_split_shift(int64_t population = 0) {
    int bits = g_skb_split_bits;
    if (bits == 0) bits = (population > 0 && population < 2500000) ? 3 : 4;
    return 32u - (uint32_t)bits;
}
static int g_skb_both_ways = 1;
static int g_skb_arena_guard = 0;
static int64_t g_skb_dense_mst_min_nodes = 12000000;
static int skb_option_both_ways() { return g_skb_both_ways; }
// the single-block handling of small junction-free cycles is a CUDA kernel: the harness always runs the general path
static int skb_cycles_small(const skb_i4*, const uint8_t*, int64_t, int64_t, const double*, int64_t, int32_t*, uint32_t*,
                            int32_t*, int32_t*, int32_t*, double*, double*, uint32_t*, int32_t*, int32_t*, skb_stream_t) {
    return 1;
}
static int skb_cycles_small_heads(const skb_i4*, const uint8_t*, int64_t, int64_t, int64_t, int64_t, const double*, const int64_t*,
                                  int64_t, int64_t, int32_t*, uint32_t*, int32_t*, uint32_t*, void*, int32_t*, int32_t*,
                                  skb_stream_t) {
    return 1;
}
static int skb_cycles_small_entries(const skb_i4*, int64_t, const double*, int64_t, int64_t, int32_t*, int32_t*, double*, double*,
                                    const int32_t*, int32_t*, skb_stream_t) {
    return -2;
}
// one-pass junction edge emission and the packed-key Boruvka launch are CUDA kernels: the harness counts, scans, emits
static int skb_junction_append(const uint64_t*, const uint32_t*, int, const int64_t*, const int64_t*, const uint32_t*, int64_t,
                               int64_t, int64_t, int32_t*, int32_t*, uint8_t*, int64_t, int32_t*, skb_stream_t,
                               const int32_t* = nullptr) {
    return 1;
}
static int skb_mst_keys_run(const int32_t*, const int32_t*, const uint8_t*, const double*, int64_t, const int32_t*, int32_t*,
                            unsigned long long*, int64_t, int32_t*, int32_t*, uint8_t*, int32_t*, uint32_t*, int32_t, int32_t,
                            skb_stream_t) {
    return -2;
}
#include "skb_api.inl"

static bool skb_host_nonzero(const void* img, int dt, int64_t r) {
    switch (dt) {
        case SKB_DT_U8: case SKB_DT_I8: return ((const uint8_t*)img)[r] != 0;
        case SKB_DT_U16: case SKB_DT_I16: return ((const uint16_t*)img)[r] != 0;
        case SKB_DT_U32: case SKB_DT_I32: return ((const uint32_t*)img)[r] != 0;
        case SKB_DT_U64: case SKB_DT_I64: return ((const uint64_t*)img)[r] != 0;
        case SKB_DT_F32: return (((const uint32_t*)img)[r] & 0x7fffffffu) != 0;  // -0.0 is zero, NaN is not
        default: return (((const uint64_t*)img)[r] & 0x7fffffffffffffffull) != 0;
    }
}

extern "C" {

const char* skb_last_error(void) { return g_skb_err; }
int64_t skb_scan_scratch_bytes(int64_t n) { return ((n + 2047) / 2048 + 2) * 8; }
int64_t skb_tile_voxels(void) { return SKB_TILE_VOX; }
int64_t skb_kernel_launches(void) { return 0; }
int skb_set_option(const char* name, int64_t value) {  // the scheduling tunables of the CUDA build do nothing here
    if (name && !strcmp(name, "emit_both_ways")) g_skb_both_ways = value ? 1 : 0;
    if (name && !strcmp(name, "arena_guard")) g_skb_arena_guard = value ? 1 : 0;
    if (name && !strcmp(name, "splitter_bits") && value >= 0 && value <= 12) g_skb_split_bits = (int)value;
    return 0;
}

int skb_pack_count(const void* image, int dtype, int64_t nvox, uint64_t* bits, uint32_t* tile_counts,
                   int64_t ntiles, int variant, void*) {
    (void)variant;
    if (dtype < 0 || dtype >= SKB_DT_COUNT) return skb_fail("unsupported image dtype");
    if (ntiles != (nvox + SKB_TILE_VOX - 1) / SKB_TILE_VOX) return skb_fail("ntiles does not match nvox");
    for (int64_t t = 0; t < ntiles; ++t) {
        uint32_t c = 0;
        for (int w = 0; w < SKB_TILE_WORDS; ++w) {
            uint64_t word = 0;
            for (int b = 0; b < 64; ++b) {
                int64_t r = (t * SKB_TILE_WORDS + w) * 64 + b;
                if (r < nvox && skb_host_nonzero(image, dtype, r)) word |= 1ull << b;
            }
            bits[t * SKB_TILE_WORDS + w] = word;
            c += (uint32_t)skb_popcll(word);
        }
        tile_counts[t] = c;
    }
    return 0;
}

int skb_emit_nodes_cap(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                       const int64_t* shape, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                       int64_t* coords, double* values, uint32_t* pre256, int64_t node_cap, void*);
int skb_emit_nodes(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                   const int64_t* shape, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                   int64_t* coords, double* values, uint32_t* pre256, void* stream) {
    return skb_emit_nodes_cap(bits, tile_prefix, ntiles, ndim, shape, image, dtype, z_origin, lin, coords, values,
                              pre256, 0xffffffffLL, stream);
}
int skb_emit_nodes_cap(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                       const int64_t* shape, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                       int64_t* coords, double* values, uint32_t* pre256, int64_t node_cap, void*) {
    SKB_TRY(skb_check_geom(ndim, shape));
    SkbGeom g = skb_make_geom(ndim, shape);
    for (int64_t t = 0; t < ntiles; ++t) {
        uint32_t id = tile_prefix[t];
        for (int w = 0; w < SKB_TILE_WORDS; ++w) {
            int64_t wi = t * SKB_TILE_WORDS + w;
            if ((w & 3) == 0) pre256[wi >> 2] = id;
            uint64_t word = bits[wi];
            while (word) {
                int b = __builtin_ctzll(word);
                word &= word - 1;
                int64_t r = wi * 64 + b;
                if ((int64_t)id >= node_cap) {  // beyond the capacity of the node arrays: counted, not written
                    ++id;
                    continue;
                }
                lin[id] = r;
                int32_t z, y, x;
                skb_coords(g, r, z, y, x);
                int64_t* c = coords + (int64_t)id * ndim;
                if (ndim == 3) { c[0] = z + z_origin; c[1] = y; c[2] = x; }
                else if (ndim == 2) { c[0] = y; c[1] = x; }
                else c[0] = x;
                if (values) values[id] = skb_load_value(image, dtype, r);
                ++id;
            }
        }
        if (id != tile_prefix[t + 1]) return skb_fail("tile prefix does not match the mask");
    }
    return 0;
}

int64_t skb_sweep_scratch_bytes(int64_t nvox) { return ((nvox + 65535) / 65536 + 2 + 64) * 8; }

int skb_sweep_nodes(const void* image, int dtype, int ndim, const int64_t* shape, int64_t z_origin, uint64_t* bits,
                    int64_t ntiles, uint32_t* pre256, int64_t node_cap, int64_t* lin, int64_t* coords, double* values,
                    void*, int32_t* n_nodes, void*) {
    SKB_TRY(skb_check_geom(ndim, shape));
    SkbGeom g = skb_make_geom(ndim, shape);
    if (dtype < 0 || dtype >= SKB_DT_COUNT) return skb_fail("unsupported image dtype");
    if (ntiles != (g.nvox + SKB_TILE_VOX - 1) / SKB_TILE_VOX) return skb_fail("ntiles does not match the shape");
    int64_t id = 0;
    for (int64_t wi = 0; wi < ntiles * SKB_TILE_WORDS; ++wi) {
        uint64_t word = 0;
        for (int b = 0; b < 64; ++b) {
            int64_t r = wi * 64 + b;
            if (r < g.nvox && skb_host_nonzero(image, dtype, r)) word |= 1ull << b;
        }
        bits[wi] = word;
        if ((wi & 3) == 0) pre256[wi >> 2] = (uint32_t)id;
        while (word) {
            int b = __builtin_ctzll(word);
            word &= word - 1;
            if (id < node_cap) {
                int64_t r = wi * 64 + b;
                if ((int64_t)id >= node_cap) {  // beyond the capacity of the node arrays: counted, not written
                    ++id;
                    continue;
                }
                lin[id] = r;
                int32_t z, y, x;
                skb_coords(g, r, z, y, x);
                int64_t* c = coords + id * ndim;
                if (ndim == 3) { c[0] = z + z_origin; c[1] = y; c[2] = x; }
                else if (ndim == 2) { c[0] = y; c[1] = x; }
                else c[0] = x;
                if (values) values[id] = skb_load_value(image, dtype, r);
            }
            ++id;
        }
    }
    n_nodes[0] = (int32_t)id;
    n_nodes[1] = 0;
    return 0;
}

}  // extern "C"

#include "skb_run.inl"
#include "skb_slab.inl"
