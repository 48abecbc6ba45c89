// skan_b200 -- sm_100a implementation of the skeleton-to-graph hot path of jni/skan.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -shared -Xcompiler -fPIC
//
// This file holds the block-cooperative kernels (the V-proportional mask sweep, node emission,
// device-wide scans) and the CUDA side of the platform layer; the per-item kernels of the
// node/edge/path proportional stages are the functors of skb_items.cuh launched through
// skb_foreach, sequenced by the C-ABI bodies in skb_api.inl.  There is no CPU fallback here:
// every entry point launches kernels on the given stream and reports CUDA errors.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "skb_items.cuh"

typedef cudaStream_t skb_stream_t;

// ------------------------------------------------------------------------------- errors
static thread_local char g_skb_err[512] = "";

static int skb_fail(const char* what) {
    snprintf(g_skb_err, sizeof(g_skb_err), "%s", what);
    return -2;
}

static int skb_cuda_check(cudaError_t e, const char* where) {
    if (e == cudaSuccess) return 0;
    snprintf(g_skb_err, sizeof(g_skb_err), "%s: %s", where, cudaGetErrorString(e));
    return -1;
}

#define SKB_CUDA(expr) \
    do { \
        int _e = skb_cuda_check((expr), #expr); \
        if (_e) return _e; \
    } while (0)

#define SKB_TRY(expr)            \
    do {                         \
        int _e = (expr);         \
        if (_e != 0) return _e;  \
    } while (0)

// kernels launched by this library since load (bench.py reports it as gpu_launches)
static unsigned long long g_skb_launches = 0;
#define SKB_COUNT_LAUNCH(k) __atomic_fetch_add(&g_skb_launches, (unsigned long long)(k), __ATOMIC_RELAXED)

// per-device facts, cached per device ordinal (a process may use several GPUs, from several threads)
#define SKB_MAX_DEV 64
static int skb_current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= SKB_MAX_DEV) return 0;
    return dev;
}
static int skb_sm_count() {
    static int sms[SKB_MAX_DEV] = {};
    const int dev = skb_current_device();
    int v = __atomic_load_n(&sms[dev], __ATOMIC_RELAXED);
    if (!v) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        __atomic_store_n(&sms[dev], v, __ATOMIC_RELAXED);
    }
    return v;
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: once per kernel AND device
template <typename K>
static int skb_optin_smem(K kernel, size_t bytes, uint8_t* done /* [SKB_MAX_DEV] */) {
    const int dev = skb_current_device();
    if (!__atomic_load_n(&done[dev], __ATOMIC_ACQUIRE)) {
        SKB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        __atomic_store_n(&done[dev], (uint8_t)1, __ATOMIC_RELEASE);
    }
    return 0;
}

// ------------------------------------------------------------------------------- abort word
// Z-slab mode synchronises the ranks of a node on the device (skb_xbar_kernel below): a rank's kernels wait for
// flags its peers write into its exchange arena.  A rank that gives up half-way (an arena too small, an error)
// or a wait that times out sets this word instead; kernels launched as "guarded" (between such a barrier and the
// host's next read-back, which reports the word) then return at once rather than interpret rows that never arrived.
__device__ int g_skb_poison = 0;
static thread_local int g_skb_guard = 0;  // host: launch the following kernels guarded

// ------------------------------------------------------------------------------- programmatic dependent launch
// The node/edge/path proportional part of a step is a chain of 40-60 short kernels, each consuming what the previous
// one wrote.  Launched plainly, a kernel's blocks are scheduled only after its predecessor has drained and flushed;
// with programmatic stream serialisation (sm_90+) they are made resident while the predecessor's last wave is still
// running and park at griddepcontrol.wait until it has completed, which hides the launch latency of every link of
// the chain.  Every kernel launched this way executes SKB_PDL_ENTRY() before it touches global memory.  Option
// "pdl" (skb_set_option; default 1) switches the attribute off for A/B measurements.
#include <utility>
static int g_skb_pdl = -1;  // -1: not set yet -> SKAN_B200_PDL in the environment ("0" switches it off), else on
static int skb_option_pdl() {
    if (g_skb_pdl < 0) {
        const char* e = getenv("SKAN_B200_PDL");
        g_skb_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_skb_pdl;
}
#define SKB_PDL_ENTRY()                                                   \
    do {                                                                  \
        asm volatile("griddepcontrol.wait;" ::: "memory");                \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   \
    } while (0)
template <typename... KArgs, typename... Args>
static cudaError_t skb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, skb_stream_t s,
                                  Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = skb_option_pdl() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------- foreach
template <typename F>
__global__ void __launch_bounds__(256) skb_foreach_kernel(F f, int64_t n, int guard) {
    SKB_PDL_ENTRY();
    if (guard && g_skb_poison) return;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

template <typename F>
static int skb_foreach(int64_t n, skb_stream_t s, F f) {
    if (n <= 0) return 0;
    int64_t blocks = (n + 255) / 256;
    int64_t cap = (int64_t)skb_sm_count() * 32;  // grid-stride beyond 32 CTAs per SM
    if (blocks > cap) blocks = cap;
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(skb_launch_pdl(skb_foreach_kernel<F>, dim3((unsigned)blocks), dim3(256), 0, s, f, n, g_skb_guard),
                          "skb_foreach launch");
}

// tunables (skb_set_option, or the environment at first use), measured on B200 in round 2 (profiles/r02_ab_walks_*.txt):
// "emit_both_ways": 1 (default) = the second walk writes every chain from both of its ends; 0 = from its head only.
static int g_skb_both_ways = -1;
static int skb_option_both_ways() {
    if (g_skb_both_ways < 0) {
        const char* e = getenv("SKAN_B200_EMIT_BOTH_WAYS");
        g_skb_both_ways = (e && e[0] == '0') ? 0 : 1;
    }
    return g_skb_both_ways;
}
// "splitter_bits": the chain voxels whose hash has this many leading zero bits split the chains into walks
// (4 = one in 16, the measured default: a walk kernel lasts as long as its longest walk, ~16 ln(starts) hops;
// more splitters = shorter walks, more starts to rank)
static int g_skb_split_bits = -1;  // -1: not set -> by the environment, else adaptive; >= 1: fixed by skb_set_option
// population = nodes (per GPU) the walks will run over, 0 = unknown.  A walk kernel lasts as long as its longest walk
// (~2^bits * ln(starts) dependent loads), whatever the volume: measured (profiles/r02_ab_walks_*.txt) 1/16 wins at
// 7.5 M nodes (2048^3), 1/8 at 0.8 M (1024^3, a Z-slab of an eighth of 2048^3), where ranking the extra starts is cheap.
static uint32_t skb_option_split_shift(int64_t population = 0) {
    if (g_skb_split_bits < 0) {
        const char* e = getenv("SKAN_B200_SPLITTER_BITS");
        const int v = e ? atoi(e) : 0;
        g_skb_split_bits = (v >= 1 && v <= 12) ? v : 0;  // 0 = adaptive
    }
    int bits = g_skb_split_bits;
    if (bits == 0) bits = (population > 0 && population < 2500000) ? 3 : 4;
    return 32u - (uint32_t)bits;
}

// stepwise items (the chain walks: begin / step / finish): one item per thread.  A variant with persistent warps
// whose idle lanes drew new walks from a ticket queue was measured slower twice (r01: 3.73 -> 4.02 ms, r02: 3.68 ->
// 3.97 ms at 2048^3) and has been removed: a walk kernel lasts as long as its longest walk, not as its lane count.
template <typename F>
static int skb_foreach_steps(int64_t n, skb_stream_t s, F f) {
    return skb_foreach(n, s, f);  // F::operator() runs begin / step / finish in place
}

// ------------------------------------------------------------------------------- foreach over unmarked items
// f(i) for every i < n with marks[i] == 0, when almost every item is marked (the nodes no terminal-headed path
// covers are a handful out of millions): a thread tests 16 marks with one 128-bit load.
template <typename F>
__global__ void __launch_bounds__(256) skb_foreach_unmarked_kernel(F f, const uint8_t* __restrict__ marks, int64_t n,
                                                                   int guard) {
    SKB_PDL_ENTRY();
    if (guard && g_skb_poison) return;
    const int64_t groups = (n + 15) >> 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t gi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += stride) {
        const int64_t v0 = gi << 4;
        if (v0 + 16 <= n) {
            const uint4 c = *(const uint4*)(marks + v0);
            const uint32_t w[4] = {c.x, c.y, c.z, c.w};
            uint32_t any_zero = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) any_zero |= (w[k] - 0x01010101u) & ~w[k] & 0x80808080u;
            if (!any_zero) continue;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (((w[j >> 2] >> (8 * (j & 3))) & 0xffu) == 0) f(v0 + j);
        } else {
            for (int64_t v = v0; v < n; ++v)
                if (!marks[v]) f(v);
        }
    }
}

template <typename F>
static int skb_foreach_unmarked(int64_t n, const uint8_t* marks, skb_stream_t s, F f) {
    if (n <= 0) return 0;
    if (((uintptr_t)marks & 15) != 0)  // unaligned marks: every item tests its own byte (the items do)
        return skb_foreach(n, s, f);
    int64_t blocks = ((n + 15) / 16 + 255) / 256;
    int64_t cap = (int64_t)skb_sm_count() * 32;
    if (blocks > cap) blocks = cap;
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(skb_launch_pdl(skb_foreach_unmarked_kernel<F>, dim3((unsigned)blocks), dim3(256), 0, s, f, marks, n,
                                         g_skb_guard),
                          "skb_foreach_unmarked launch");
}

static int skb_fill_bytes(void* p, int byte, size_t nbytes, skb_stream_t s) {
    if (!nbytes) return 0;
    return skb_cuda_check(cudaMemsetAsync(p, byte, nbytes, s), "cudaMemsetAsync");
}

static int skb_copy_bytes(void* dst, const void* src, size_t nbytes, skb_stream_t s) {
    if (!nbytes) return 0;
    return skb_cuda_check(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync");
}

// Host read-back of a few device ints -- the only places where the host waits for the device.  A one-warp
// kernel writes the values straight into mapped pinned host memory and then a sequence number; the host
// spins on that word.  Compared with cudaMemcpyAsync + cudaStreamSynchronize this saves the copy engine
// round trip and the driver's wake-up on each of the 8-13 read-backs of a step.  The stream is checked
// from time to time so that a failed kernel is reported instead of being waited for.
struct SkbMailbox {
    volatile int32_t* host;  // [0..SKB_MAX_FETCH) values, [SKB_MAX_FETCH] sequence number, [SKB_MAX_FETCH + 1] abort word
    int32_t* dev;            // the same memory as the device sees it
    int32_t seq;
};
static int skb_mailbox(SkbMailbox** out) {
    static thread_local SkbMailbox box = {nullptr, nullptr, 0};
    if (!box.host) {
        void* h = nullptr;
        SKB_CUDA(cudaHostAlloc(&h, (SKB_MAX_FETCH + 2) * 4, cudaHostAllocMapped | cudaHostAllocPortable));
        memset(h, 0, (SKB_MAX_FETCH + 2) * 4);
        void* d = nullptr;
        SKB_CUDA(cudaHostGetDevicePointer(&d, h, 0));
        box.host = (volatile int32_t*)h;
        box.dev = (int32_t*)d;
    }
    *out = &box;
    return 0;
}

#define SKB_POISONED 3  // = SKB_RUN_RESTART of skb_slab.inl: the call is to be repeated after skb_comm_reset
__global__ void skb_poison_clear_kernel() { g_skb_poison = 0; }

__global__ void skb_mail_kernel(SkbGatherI32Item it, int n, int32_t seq) {
    SKB_PDL_ENTRY();
    const int lane = threadIdx.x;
    if (lane < n) it.out[lane] = *it.ptr[lane];
    if (lane == 31) it.out[SKB_MAX_FETCH + 1] = g_skb_poison;  // every read-back also reports the abort word
    __threadfence_system();
    __syncwarp();
    if (lane == 0) *(volatile int32_t*)(it.out + SKB_MAX_FETCH) = seq;
}

// A read-back in two halves: skb_fetch_post enqueues the mail kernel (it runs when everything enqueued before it
// has finished), skb_fetch_wait spins until its values have arrived.  Kernels enqueued between the two keep the
// device busy while the values travel: the runner posts a count, launches the first consumer sized by a
// capacity hint, and only then waits.  One read-back per host thread may be in flight.
static int skb_fetch_post(const int32_t* const* ptrs, int n, skb_stream_t s) {
    if (n > SKB_MAX_FETCH) return skb_fail("too many values to fetch");
    SkbMailbox* box;
    {
        int e = skb_mailbox(&box);
        if (e) return e;
    }
    SkbGatherI32Item it;
    for (int i = 0; i < SKB_MAX_FETCH; ++i) it.ptr[i] = n ? ptrs[i < n ? i : 0] : nullptr;
    it.out = box->dev;
    const int32_t seq = ++box->seq;
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(skb_launch_pdl(skb_mail_kernel, dim3(1), dim3(32), 0, s, it, n, seq), "read-back launch");
}

static int skb_fetch_wait(int n, int64_t* out, skb_stream_t s) {
    SkbMailbox* box;
    {
        int e = skb_mailbox(&box);
        if (e) return e;
    }
    const int32_t seq = box->seq;
    for (unsigned spins = 1;; ++spins) {
        if (box->host[SKB_MAX_FETCH] == seq) break;
        if ((spins & 0xfffffu) == 0) {  // ~ every few milliseconds: is the stream still alive?
            cudaError_t e = cudaStreamQuery(s);
            if (e != cudaSuccess && e != cudaErrorNotReady) return skb_cuda_check(e, "read-back");
        }
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    for (int i = 0; i < n; ++i) out[i] = box->host[i];
    if (box->host[SKB_MAX_FETCH + 1] != 0) {  // a device-side barrier of the Z-slab mode was given up (a peer left, or a
        skb_poison_clear_kernel<<<1, 1, 0, s>>>();  // wait timed out): guarded kernels have been skipped since
        SKB_COUNT_LAUNCH(1);
        snprintf(g_skb_err, sizeof(g_skb_err), "a device-side barrier between the ranks was abandoned");
        return SKB_POISONED;
    }
    return 0;
}

static int skb_fetch_i32s(const int32_t* const* ptrs, int n, int64_t* out, skb_stream_t s) {
    int e = skb_fetch_post(ptrs, n, s);
    if (e) return e;
    return skb_fetch_wait(n, out, s);
}

// everything enqueued on the stream so far has completed (an empty read-back)
static int skb_stream_sync(skb_stream_t s) { return skb_fetch_i32s(nullptr, 0, nullptr, s); }

// Barrier between the ranks of a node on the device, in stream order (Z-slab mode): every rank owns a block of one
// 64-bit flag per peer at the end of its exchange arena.  The kernel stores this barrier's sequence number into its
// slot of EVERY rank's block (st.release.sys after a system fence: what earlier kernels of this stream wrote into this
// rank's arena is complete, and the peers read it there, through this GPU's L2), then one lane per peer polls the own
// block (ld.acquire.sys) until every flag has reached the number.  Sequence numbers only grow: (generation << 32) |
// count, the generation advanced by skb_comm_reset.  A rank that leaves a call half-way stores the generation's
// largest number (low word all ones) instead: every wait of the generation ends at once and sets g_skb_poison, as
// does a wait that lasts longer than SKB_XBAR_TIMEOUT_NS (a peer died).  No host thread takes part.
#define SKB_XBAR_TIMEOUT_NS 20000000000ull
struct SkbXbarArgs {
    unsigned long long* peer_slot[SKB_MAX_RANKS];  // this rank's slot in every rank's flag block
    const unsigned long long* mine;                // this rank's flag block
    unsigned long long seq;
    int world, poison;
};
__device__ __forceinline__ unsigned long long skb_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void skb_xbar_kernel(const SkbXbarArgs a) {
    const int lane = threadIdx.x;
    const bool on = lane < a.world;
    const bool dead = a.poison || g_skb_poison != 0;  // an abandoned call abandons the barriers that follow as well
    if (on) {
        const unsigned long long v = dead ? (a.seq | 0xffffffffull) : a.seq;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.peer_slot[lane]), "l"(v) : "memory");
    }
    if (dead) return;
    const unsigned long long t0 = skb_globaltimer();
    bool ok = !on, bad = false;
    for (unsigned spins = 1;; ++spins) {
        if (!ok) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.mine + lane) : "memory");
            if (v >= a.seq) {
                ok = true;
                bad = (uint32_t)v == 0xffffffffu;
            }
        }
        if ((spins & 255u) == 0 && skb_globaltimer() - t0 > SKB_XBAR_TIMEOUT_NS) bad = true;
        if (__any_sync(0xffffffffu, bad) || __all_sync(0xffffffffu, ok)) break;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) g_skb_poison = 1;
}
static int skb_xbar_launch(void* const* xbase, int rank, int world, int64_t flag_off, uint64_t seq, int poison,
                           skb_stream_t s) {
    SkbXbarArgs a{};
    for (int k = 0; k < world; ++k) a.peer_slot[k] = (unsigned long long*)((uint8_t*)xbase[k] + flag_off) + rank;
    a.mine = (const unsigned long long*)((uint8_t*)xbase[rank] + flag_off);
    a.seq = seq;
    a.world = world;
    a.poison = poison;
    skb_xbar_kernel<<<1, 32, 0, s>>>(a);
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(cudaGetLastError(), "rank barrier launch");
}
// skb_comm_reset: the abort word is cleared once everything enqueued has run
static int skb_poison_reset() {
    SKB_CUDA(cudaDeviceSynchronize());
    const int zero = 0;
    SKB_CUDA(cudaMemcpyToSymbol(g_skb_poison, &zero, sizeof(int)));
    return 0;
}

// exchange arena of the Z-slab mode: device memory every rank of the node can read (CUDA IPC; the
// peers map it once, kernels then load from it directly over NVLink)
static int skb_xalloc(const char*, int, int64_t bytes, void** ptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle does not fit the control block");
    SKB_CUDA(cudaMalloc(ptr, (size_t)bytes));
    SKB_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    SKB_CUDA(cudaIpcGetMemHandle(&h, *ptr));
    memset(handle, 0, 64);
    memcpy(handle, &h, sizeof(h));
    return 0;
}
static int skb_xopen(const uint8_t handle[64], int64_t, void** ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return skb_cuda_check(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}
static int skb_xclose(void* ptr, int64_t) { return skb_cuda_check(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle"); }
static int skb_xfree(void* ptr, int64_t, const char*, int) { return skb_cuda_check(cudaFree(ptr), "cudaFree"); }
static int skb_xunlink(const char*, int) { return 0; }

// CUDA events around one kernel on its launching stream (bench.py's roofline figure)
struct SkbTimer {
    cudaEvent_t a, b;
};
static int skb_timer_start(void** handle, skb_stream_t s) {
    static thread_local SkbTimer per_dev[SKB_MAX_DEV] = {};  // an event belongs to the device it was created on
    SkbTimer& t = per_dev[skb_current_device()];
    if (!t.a) {
        SKB_CUDA(cudaEventCreate(&t.a));
        SKB_CUDA(cudaEventCreate(&t.b));
    }
    *handle = &t;
    return skb_cuda_check(cudaEventRecord(t.a, s), "cudaEventRecord");
}
static int skb_timer_stop(void* handle, skb_stream_t s) {
    return skb_cuda_check(cudaEventRecord(((SkbTimer*)handle)->b, s), "cudaEventRecord");
}
static int skb_timer_read_ns(void* handle, int64_t* ns) {  // call after the stream has been synchronised
    float ms = 0.f;
    SkbTimer* t = (SkbTimer*)handle;
    SKB_CUDA(cudaEventSynchronize(t->b));
    SKB_CUDA(cudaEventElapsedTime(&ms, t->a, t->b));
    *ns = (int64_t)((double)ms * 1e6);
    return 0;
}

// stage clock of the native runner (skb_run_params.time_sweep == 2): one event per stage boundary
#define SKB_MAX_MARKS 32
static thread_local cudaEvent_t g_skb_marks_dev[SKB_MAX_DEV][SKB_MAX_MARKS] = {};
static int skb_stage_mark(int idx, skb_stream_t s) {
    if (idx < 0 || idx >= SKB_MAX_MARKS) return skb_fail("stage mark out of range");
    cudaEvent_t* g_skb_marks = g_skb_marks_dev[skb_current_device()];
    if (!g_skb_marks[idx]) SKB_CUDA(cudaEventCreate(&g_skb_marks[idx]));
    return skb_cuda_check(cudaEventRecord(g_skb_marks[idx], s), "cudaEventRecord");
}
static int skb_stage_read(const uint8_t* marked, int n, int64_t* ns) {  // ns[i] = time from the previous recorded mark to mark i
    cudaEvent_t* g_skb_marks = g_skb_marks_dev[skb_current_device()];
    int prev = -1;
    for (int i = 0; i < n; ++i) {
        ns[i] = 0;
        if (!marked[i]) continue;
        if (prev >= 0) {
            float ms = 0.f;
            SKB_CUDA(cudaEventSynchronize(g_skb_marks[i]));
            SKB_CUDA(cudaEventElapsedTime(&ms, g_skb_marks[prev], g_skb_marks[i]));
            ns[i] = (int64_t)((double)ms * 1e6);
        }
        prev = i;
    }
    return 0;
}

static int skb_fetch_i32(const int32_t* p, int32_t* host_out, skb_stream_t s) {
    int64_t v = 0;
    int r = skb_fetch_i32s(&p, 1, &v, s);
    *host_out = (int32_t)v;
    return r;
}

// ------------------------------------------------------------------------------- scans
// Exclusive scan with n+1 outputs (out[n] = total).  `in == out` is allowed: every block holds
// its whole tile in registers before it writes.
#define SKB_SCAN_THREADS 512
#ifndef SKB_SCAN_ITEMS
#define SKB_SCAN_ITEMS 16  // 8192 items per tile: the look-back chain advances one window of tiles per L2 round trip, so fewer, larger tiles
#endif
#ifndef SKB_SCAN_WIDE
#define SKB_SCAN_WIDE 0    // 1: the look-back inspects 128 predecessors per round (4 per lane) instead of 32
#endif
#define SKB_SCAN_TILE (SKB_SCAN_THREADS * SKB_SCAN_ITEMS)

template <typename T>
__device__ __forceinline__ T skb_block_exclusive(T v, T* total) {
    // exclusive scan of one value per thread across a 256-thread block
    __shared__ T warp_sums[SKB_SCAN_THREADS / 32];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < SKB_SCAN_THREADS / 32 ? warp_sums[lane] : (T)0;
#pragma unroll
        for (int d = 1; d < SKB_SCAN_THREADS / 32; d <<= 1) {
            T o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += o;
        }
        if (lane < SKB_SCAN_THREADS / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    T base = warp ? warp_sums[warp - 1] : (T)0;
    *total = warp_sums[SKB_SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// Polling load of a status word another CTA publishes: the "memory" clobber keeps the compiler from
// hoisting it out of the spin loop (it does hoist __ldcg, whose asm has no clobber).
__device__ __forceinline__ unsigned long long skb_poll_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Single-pass u32 scan (decoupled look-back): every value is read once and written once.
// scratch (zeroed once by the caller, reused by later scans on the same stream): word 0 = dynamic
// tile counter (self-resetting), word 1 = unused, then one 64-bit status per tile =
// (epoch << 34) | (state << 32) | value; the epoch makes stale statuses of earlier scans invalid.
#define SKB_ST_AGGREGATE 1ull
#define SKB_ST_INCLUSIVE 2ull
static unsigned int g_skb_scan_epoch = 0;

__global__ void __launch_bounds__(SKB_SCAN_THREADS) skb_scan_lookback_kernel(const uint32_t* in, int64_t n,
                                                                             uint32_t* out,
                                                                             unsigned long long* scratch,
                                                                             unsigned long long epoch, int64_t nb,
                                                                             const int32_t* n_ptr, uint32_t* total_out,
                                                                             int guard) {
    // n_ptr != NULL: the length is read on the device (at most n, the capacity the grid was sized by); total_out != NULL
    // receives the total as well (a fixed address, where out[n] is not one)
    SKB_PDL_ENTRY();
    __shared__ int64_t s_tile;
    __shared__ uint32_t s_excl;
    unsigned int* counter = (unsigned int*)scratch;
    unsigned long long* status = scratch + 1;
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        if (t == gridDim.x - 1) *counter = 0u;  // every tile has been handed out: reset for the next scan
        s_tile = t;
    }
    __syncthreads();
    if (guard && g_skb_poison) return;  // (after the ticket: the counter still resets itself for the next scan)
    const int64_t tile = s_tile;
    if (n_ptr) {
        int64_t m = *n_ptr;
        if (m < 0) m = 0;
        if (m < n) n = m;
        nb = (n + SKB_SCAN_TILE - 1) / SKB_SCAN_TILE;
        if (n == 0 && tile == 0 && threadIdx.x == 0) {
            out[0] = 0u;
            if (total_out) *total_out = 0u;
        }
        if (tile >= nb) return;  // the same decision in every thread of the block
    }
    const int64_t base = tile * SKB_SCAN_TILE + (int64_t)threadIdx.x * SKB_SCAN_ITEMS;
    uint32_t v[SKB_SCAN_ITEMS];
    uint32_t acc = 0;
    // a thread's 16 values are 64 contiguous bytes: four 128-bit accesses when the arrays allow it
    const bool vec = base + SKB_SCAN_ITEMS <= n && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0;
    if (vec) {
#pragma unroll
        for (int j = 0; j < SKB_SCAN_ITEMS / 4; ++j) {
            const uint4 q = *(const uint4*)(in + base + 4 * j);
            v[4 * j] = q.x;
            v[4 * j + 1] = q.y;
            v[4 * j + 2] = q.z;
            v[4 * j + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SKB_SCAN_ITEMS; ++j) v[j] = (base + j < n) ? in[base + j] : 0u;
    }
#pragma unroll
    for (int j = 0; j < SKB_SCAN_ITEMS; ++j) acc += v[j];
    uint32_t total;
    uint32_t ex = skb_block_exclusive<uint32_t>(acc, &total);
    const unsigned long long tag = epoch << 34;
    if (threadIdx.x == 0) {
        unsigned long long st = tag | ((tile == 0 ? SKB_ST_INCLUSIVE : SKB_ST_AGGREGATE) << 32) | total;
        __stcg(status + tile, st);
        if (tile == 0) s_excl = 0;
    }
#if SKB_SCAN_WIDE
    if (tile > 0 && threadIdx.x < 32) {
        // warp 0 looks back over the 128 nearest predecessors at a time: lane l holds look - 4 l - j, j = 0..3 (j = 0 the
        // nearest).  The chain of inclusive prefixes advances one window per L2 round trip, so a wider window shortens it.
        const int lane = threadIdx.x;
        uint32_t excl = 0;
        int64_t look = tile - 1;
        for (;;) {
            unsigned long long st[4];
            bool ready;
            do {
                ready = true;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t idx = look - 4 * lane - j;
                    st[j] = idx >= 0 ? skb_poll_u64(status + idx) : (tag | (SKB_ST_INCLUSIVE << 32));
                    ready = ready && (st[j] >> 34) == epoch && ((st[j] >> 32) & 3ull) != 0ull;
                }
            } while (__any_sync(0xffffffffu, !ready));
            uint32_t lsum = 0;
            bool linc = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!linc) {
                    lsum += (uint32_t)(st[j] & 0xffffffffull);
                    linc = ((st[j] >> 32) & 3ull) == SKB_ST_INCLUSIVE;
                }
            }
            const unsigned incl = __ballot_sync(0xffffffffu, linc);
            const int stop = incl ? (__ffs((int)incl) - 1) : 31;
            uint32_t val = lane <= stop ? lsum : 0u;
#pragma unroll
            for (int d = 16; d; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
            excl += val;
            if (incl) break;
            look -= 128;
        }
        if (lane == 0) {
            __stcg(status + tile, tag | (SKB_ST_INCLUSIVE << 32) | (unsigned long long)(excl + total));
            s_excl = excl;
        }
    }
#else
    if (tile > 0 && threadIdx.x < 32) {
        // warp 0 looks back over the 32 nearest predecessors at a time
        const int lane = threadIdx.x;
        uint32_t excl = 0;
        int64_t look = tile - 1;
        for (;;) {
            int64_t idx = look - lane;
            unsigned long long st;
            do {
                st = idx >= 0 ? skb_poll_u64(status + idx) : (tag | (SKB_ST_INCLUSIVE << 32));
            } while (__any_sync(0xffffffffu, (st >> 34) != epoch || ((st >> 32) & 3ull) == 0ull));
            unsigned incl = __ballot_sync(0xffffffffu, ((st >> 32) & 3ull) == SKB_ST_INCLUSIVE);
            int stop = incl ? (__ffs((int)incl) - 1) : 31;
            uint32_t val = lane <= stop ? (uint32_t)(st & 0xffffffffull) : 0u;
#pragma unroll
            for (int d = 16; d; d >>= 1) val += __shfl_xor_sync(0xffffffffu, val, d);
            excl += val;
            if (incl) break;
            look -= 32;
        }
        if (lane == 0) {
            __stcg(status + tile, tag | (SKB_ST_INCLUSIVE << 32) | (unsigned long long)(excl + total));
            s_excl = excl;
        }
    }
#endif
    __syncthreads();
    ex += s_excl;
    if (vec) {
#pragma unroll
        for (int j = 0; j < SKB_SCAN_ITEMS / 4; ++j) {
            uint4 q;
            q.x = ex;
            q.y = q.x + v[4 * j];
            q.z = q.y + v[4 * j + 1];
            q.w = q.z + v[4 * j + 2];
            ex = q.w + v[4 * j + 3];
            *(uint4*)(out + base + 4 * j) = q;
        }
    } else {
#pragma unroll
        for (int j = 0; j < SKB_SCAN_ITEMS; ++j) {
            if (base + j < n) out[base + j] = ex;
            ex += v[j];
        }
    }
    if (tile == nb - 1 && threadIdx.x == SKB_SCAN_THREADS - 1) {
        out[n] = ex;
        if (total_out) *total_out = ex;
    }
}

// n_ptr (may be NULL): the length is *n_ptr, read on the device, at most n; total_out (may be NULL): also gets the total
static int skb_scan_u32_dev(const uint32_t* in, uint32_t* out, int64_t n, const int32_t* n_ptr, uint32_t* total_out,
                            void* scratch, skb_stream_t s) {
    if (n < 0) return skb_fail("scan: negative length");
    if (n == 0) {
        SKB_CUDA(cudaMemsetAsync(out, 0, sizeof(uint32_t), s));
        if (total_out) SKB_CUDA(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), s));
        return 0;
    }
    int64_t nb = (n + SKB_SCAN_TILE - 1) / SKB_SCAN_TILE;
    unsigned long long epoch = (unsigned long long)(__atomic_add_fetch(&g_skb_scan_epoch, 1u, __ATOMIC_RELAXED) & 0x3fffffffu);
    if (epoch == 0) epoch = (unsigned long long)(__atomic_add_fetch(&g_skb_scan_epoch, 1u, __ATOMIC_RELAXED) & 0x3fffffffu);
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(skb_launch_pdl(skb_scan_lookback_kernel, dim3((unsigned)nb), dim3(SKB_SCAN_THREADS), 0, s, in, n, out,
                                         (unsigned long long*)scratch, epoch, nb, n_ptr, total_out, g_skb_guard),
                          "scan launch");
}
static int skb_scan_u32_impl(const uint32_t* in, uint32_t* out, int64_t n, void* scratch, skb_stream_t s) {
    return skb_scan_u32_dev(in, out, n, nullptr, nullptr, scratch, s);
}
// ------------------------------------------------------------------------------- mask sweep
// The only V-proportional pass (replaces image.astype(bool), np.pad and both np.flatnonzero
// scans, csr.py:1070,122-123,131): read the image once, write 1 bit per voxel and one count
// per 16384-voxel tile.  A tile is 256 u64 words; each warp owns 2048 voxels of it.
#define SKB_TILE_VOX 16384
#define SKB_TILE_WORDS 256
#define SKB_PACK_THREADS 256

template <int ESIZE>
__device__ __forceinline__ uint32_t skb_nonzero_bits(uint4 v, bool is_float);

template <>
__device__ __forceinline__ uint32_t skb_nonzero_bits<1>(uint4 v, bool) {
    // 16 voxels: byte != 0 -> bit
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t ne = __vcmpne4(w[i], 0u) & 0x01010101u;  // 0x01 per nonzero byte
        m |= ((ne * 0x10204080u) >> 28) << (4 * i);       // gather bits 0,8,16,24 -> nibble
    }
    return m;
}
template <>
__device__ __forceinline__ uint32_t skb_nonzero_bits<2>(uint4 v, bool is_float) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t clr = is_float ? 0x7fff7fffu : 0xffffffffu;  // -0.0 is zero
    uint32_t m = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t x = w[i] & clr;
        m |= ((x & 0xffffu) ? 1u : 0u) << (2 * i);
        m |= ((x >> 16) ? 1u : 0u) << (2 * i + 1);
    }
    return m;
}
template <>
__device__ __forceinline__ uint32_t skb_nonzero_bits<4>(uint4 v, bool is_float) {
    uint32_t clr = is_float ? 0x7fffffffu : 0xffffffffu;
    return ((v.x & clr) ? 1u : 0u) | ((v.y & clr) ? 2u : 0u) | ((v.z & clr) ? 4u : 0u) | ((v.w & clr) ? 8u : 0u);
}
template <>
__device__ __forceinline__ uint32_t skb_nonzero_bits<8>(uint4 v, bool is_float) {
    uint32_t clr = is_float ? 0x7fffffffu : 0xffffffffu;
    return ((v.x | (v.y & clr)) ? 1u : 0u) | ((v.z | (v.w & clr)) ? 2u : 0u);
}

__device__ __forceinline__ uint4 skb_ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// guarded element-wise load of one 16-byte chunk that crosses the end of the image
template <int ESIZE>
__device__ __forceinline__ uint4 skb_load_tail(const uint8_t* img, int64_t byte_off, int64_t nbytes) {
    union {
        uint4 v;
        uint8_t b[16];
    } u;
    u.v = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < 16; ++i)
        if (byte_off + i < nbytes) u.b[i] = img[byte_off + i];
    return u.v;
}

// Variant 0: coalesced 128-bit streaming loads, 4 in flight per thread, bits gathered with
// xor-shuffles inside groups of 32/EPV lanes.
template <int ESIZE>
__global__ void __launch_bounds__(SKB_PACK_THREADS) skb_pack_ldg_kernel(const uint8_t* __restrict__ img,
                                                                        int64_t nvox, bool is_float,
                                                                        uint32_t* __restrict__ bits32,
                                                                        uint32_t* __restrict__ tile_counts,
                                                                        int64_t ntiles) {
    constexpr int EPV = 16 / ESIZE;             // voxels per 16-byte chunk
    constexpr int LANES_PER_WORD = 32 / EPV;    // lanes that share one 32-bit mask word
    constexpr int CHUNKS_PER_WARP = 2048 / EPV / 32;  // 16-byte chunks per lane per tile
    constexpr int UNROLL = CHUNKS_PER_WARP < 4 ? CHUNKS_PER_WARP : 4;
    __shared__ uint32_t warp_counts[SKB_PACK_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nbytes = nvox * ESIZE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t vox0 = tile * SKB_TILE_VOX + (int64_t)warp * 2048;  // first voxel of this warp
        uint32_t cnt = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < CHUNKS_PER_WARP; c0 += UNROLL) {
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                int64_t chunk_vox = vox0 + (int64_t)((c0 + u) * 32 + lane) * EPV;
                int64_t off = chunk_vox * ESIZE;
                if (off + 16 <= nbytes) v[u] = skb_ldg_stream((const uint4*)(img + off));
                else if (off < nbytes) v[u] = skb_load_tail<ESIZE>(img, off, nbytes);
                else v[u] = make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                uint32_t m = skb_nonzero_bits<ESIZE>(v[u], is_float);
                cnt += __popc(m);
                uint32_t word = m << ((lane % LANES_PER_WORD) * EPV);
#pragma unroll
                for (int d = 1; d < LANES_PER_WORD; d <<= 1) word |= __shfl_xor_sync(0xffffffffu, word, d);
                if ((lane % LANES_PER_WORD) == 0) {
                    int64_t chunk_vox = vox0 + (int64_t)((c0 + u) * 32 + lane) * EPV;
                    bits32[chunk_vox >> 5] = word;
                }
            }
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0) warp_counts[warp] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < SKB_PACK_THREADS / 32; ++w) t += warp_counts[w];
            tile_counts[tile] = t;
        }
        __syncthreads();
    }
}

// Variant 1: the tile is brought into shared memory by the TMA unit (cp.async.bulk, 1-D, one
// elected thread per CTA arms an mbarrier with the byte count), double-buffered so the next
// tile streams in while this one is packed.  Threads then read their own 32 voxels with
// 128-bit shared loads in a lane-rotated order that is bank-conflict free for every dtype.
__device__ __forceinline__ uint32_t skb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void skb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(skb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void skb_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(skb_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void skb_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(skb_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void skb_tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            skb_smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(skb_smem_u32(bar))
        : "memory");
}

// One pipeline stage is 16 KB of image = 16384 / ESIZE voxels; thread t owns bytes [64 t, 64 t + 64).
#define SKB_STAGE_BYTES 16384

template <int ESIZE, int STAGES>
__global__ void __launch_bounds__(SKB_PACK_THREADS) skb_pack_tma_kernel(const uint8_t* __restrict__ img,
                                                                        int64_t nvox, bool is_float,
                                                                        uint8_t* __restrict__ bits8,
                                                                        uint32_t* __restrict__ tile_counts,
                                                                        int64_t ntiles) {
    constexpr int NBITS = 64 / ESIZE;  // voxels (= mask bits) per thread and stage
    constexpr int EPV = 16 / ESIZE;    // voxels per 16-byte chunk
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full_bar[STAGES];
    __shared__ uint32_t warp_counts[2][SKB_PACK_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nbytes = nvox * ESIZE;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) skb_mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // this CTA owns tiles blockIdx.x, blockIdx.x + gridDim.x, ...; a tile is ESIZE stages
    int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t my_stages = my_tiles * ESIZE;
    auto stage_offset = [&](int64_t q) -> int64_t {
        int64_t tile = blockIdx.x + (q / ESIZE) * (int64_t)gridDim.x;
        return (tile * ESIZE + (q % ESIZE)) * (int64_t)SKB_STAGE_BYTES;
    };
    auto issue = [&](int64_t q) {  // thread 0 only
        int s = (int)(q % STAGES);
        int64_t off = stage_offset(q);
        int64_t avail = nbytes - off;
        uint32_t bytes = avail >= SKB_STAGE_BYTES ? (uint32_t)SKB_STAGE_BYTES : (uint32_t)(avail > 0 ? (avail & ~15LL) : 0);
        skb_mbar_expect_tx(&full_bar[s], bytes);
        if (bytes) skb_tma_load_1d(smem + (size_t)s * SKB_STAGE_BYTES, img + off, bytes, &full_bar[s]);
    };
    if (tid == 0)
        for (int64_t q = 0; q < STAGES - 1 && q < my_stages; ++q) issue(q);
    uint32_t tile_acc = 0;  // thread 0
    for (int64_t q = 0; q < my_stages; ++q) {
        const int s = (int)(q % STAGES);
        const uint32_t parity = (uint32_t)((q / STAGES) & 1);
        // the stage refilled here was read in iteration q-1; every thread has passed that
        // iteration's __syncthreads before thread 0 gets here
        if (tid == 0 && q + STAGES - 1 < my_stages) issue(q + STAGES - 1);
        skb_mbar_wait(&full_bar[s], parity);
        const uint8_t* st = smem + (size_t)s * SKB_STAGE_BYTES;
        const int64_t off = stage_offset(q);
        const int64_t copied = (nbytes - off >= SKB_STAGE_BYTES) ? SKB_STAGE_BYTES : ((nbytes - off) > 0 ? ((nbytes - off) & ~15LL) : 0);
        uint64_t word = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int jj = (j + (lane >> 1)) & 3;  // rotation: the 8 lanes of a quarter-warp hit 32 distinct banks
            int boff = tid * 64 + jj * 16;
            uint4 v;
            if (boff + 16 <= copied) v = *(const uint4*)(st + boff);
            else if (off + boff < nbytes) v = skb_load_tail<ESIZE>(img, off + boff, nbytes);
            else v = make_uint4(0, 0, 0, 0);
            word |= (uint64_t)skb_nonzero_bits<ESIZE>(v, is_float) << (jj * EPV);
        }
        // mask bytes of this thread: voxel index of the stage start is off / ESIZE
        const int64_t bit0 = off / ESIZE + (int64_t)tid * NBITS;
        if (NBITS == 64) *(uint64_t*)(bits8 + (bit0 >> 3)) = word;
        else if (NBITS == 32) *(uint32_t*)(bits8 + (bit0 >> 3)) = (uint32_t)word;
        else if (NBITS == 16) *(uint16_t*)(bits8 + (bit0 >> 3)) = (uint16_t)word;
        else bits8[bit0 >> 3] = (uint8_t)word;
        uint32_t cnt = __reduce_add_sync(0xffffffffu, (uint32_t)__popcll(word));
        if (lane == 0) warp_counts[q & 1][warp] = cnt;
        __syncthreads();  // stage s fully read; warp counts visible
        if (tid == 0) {
#pragma unroll
            for (int w = 0; w < SKB_PACK_THREADS / 32; ++w) tile_acc += warp_counts[q & 1][w];
            if ((q % ESIZE) == ESIZE - 1) {
                tile_counts[blockIdx.x + (q / ESIZE) * (int64_t)gridDim.x] = tile_acc;
                tile_acc = 0;
            }
        }
    }
}

template <int ESIZE>
static int skb_pack_launch(const void* img, int64_t nvox, bool is_float, uint64_t* bits, uint32_t* tile_counts,
                           int64_t ntiles, int variant, skb_stream_t s) {
    int sms = skb_sm_count();
    if (variant == 1) {
        constexpr int STAGES = 4;
        size_t smem = (size_t)STAGES * SKB_STAGE_BYTES;
        static uint8_t attr_set[SKB_MAX_DEV] = {};
        SKB_TRY(skb_optin_smem(skb_pack_tma_kernel<ESIZE, STAGES>, smem, attr_set));
        int64_t grid = (int64_t)sms * 3;
        if (grid > ntiles) grid = ntiles;
        skb_pack_tma_kernel<ESIZE, STAGES><<<(unsigned)grid, SKB_PACK_THREADS, smem, s>>>(
            (const uint8_t*)img, nvox, is_float, (uint8_t*)bits, tile_counts, ntiles);
    } else {
        int64_t grid = (int64_t)sms * 8;
        if (grid > ntiles) grid = ntiles;
        skb_pack_ldg_kernel<ESIZE><<<(unsigned)grid, SKB_PACK_THREADS, 0, s>>>(
            (const uint8_t*)img, nvox, is_float, (uint32_t*)bits, tile_counts, ntiles);
    }
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(cudaGetLastError(), "pack launch");
}

// ------------------------------------------------------------------------------- node emission
// One WARP per 16384-voxel tile of the mask (256 u64 words = 128 uint4, four per lane, loaded
// with coalesced 128-bit loads): rank structure (pre256), raveled index, coordinates
// (csr.py:1088) and fp64 values (csr.py:554-562) of every foreground voxel, in raster order
// (csr.py:131-132).  Reads V/8 bytes of mask; no block-wide barrier anywhere.
#define SKB_EMIT_THREADS 256

// coordinates of voxel (tile origin + off) from the tile origin's coordinates: 32-bit arithmetic only.
// (Folding the stencil stage in here -- the coordinates are at hand, the rows are being streamed -- was measured in
// round 2 and lost: 0.26 + 0.23 ms as two kernels, 0.52-0.80 ms as one; the emission is issue-bound and the stencil's
// dependent loads sit badly in its per-tile chain.)
__device__ __forceinline__ void skb_emit_one(uint32_t off, int64_t r0, int32_t zt, int64_t z_add, int32_t y0, int32_t x0,
                                             uint32_t id, const SkbGeom& g, const void* img, int dtype,
                                             int64_t* __restrict__ lin, int64_t* __restrict__ coords,
                                             double* __restrict__ values) {
    int64_t r = r0 + off;
    lin[id] = r;
    uint32_t t = (uint32_t)x0 + off;
    uint32_t q = skb_udiv(t, g.dx);
    uint32_t x = t - q * (uint32_t)g.nx;
    uint32_t t2 = (uint32_t)y0 + q;
    uint32_t q2 = skb_udiv(t2, g.dy);
    uint32_t y = t2 - q2 * (uint32_t)g.ny;
    const int32_t zl = zt + (int32_t)q2;  // plane inside this volume / slab
    int64_t z = z_add + zl;
    int64_t* c = coords + (int64_t)id * g.ndim;
    if (g.ndim == 3) {
        c[0] = z;
        c[1] = y;
        c[2] = x;
    } else if (g.ndim == 2) {
        c[0] = y;
        c[1] = x;
    } else {
        c[0] = x;
    }
    if (values) values[id] = skb_load_value(img, dtype, r);
}

// (z, y, x) += (dz, dy, dx) in mixed radix (ny, nx); all inputs are valid digits, so one carry per digit suffices
__device__ __forceinline__ void skb_advance_coords(const SkbGeom& g, int32_t& z, int32_t& y, int32_t& x, int32_t dz,
                                                   int32_t dy, int32_t dx) {
    x += dx;
    int32_t c = x >= g.nx ? 1 : 0;
    x -= c ? g.nx : 0;
    y += dy + c;
    c = y >= g.ny ? 1 : 0;
    y -= c ? g.ny : 0;
    z += dz + c;
}

// One tile of the node emission: the rank structure (pre256), and the tile's 256 mask words plus the inclusive count of
// set bits up to and including every word parked in shared memory for the emission loop.  Returns the tile's count.
__device__ __forceinline__ uint32_t skb_rank_tile(const int64_t tile, const uint32_t t0, const uint32_t total, const int lane,
                                                  const uint64_t* __restrict__ bits, uint4* __restrict__ s_words,
                                                  uint32_t* __restrict__ s_incl, uint32_t* __restrict__ pre256) {
    if (total == 0) {  // empty tile: only the rank structure (64 blocks of 256 bits)
        ((uint2*)(pre256 + tile * 64))[lane] = make_uint2(t0, t0);
        return 0u;
    }
    // coalesced 128-bit loads: lane holds words 2*(32 j + lane) and +1 for j = 0..3
    const uint4* src = (const uint4*)(bits + tile * SKB_TILE_WORDS) + lane;
    uint4 q[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) q[j] = src[32 * j];
    uint32_t c0[4], c1[4];
    // one warp scan for the four groups: 16-bit fields (a group holds at most 4096 set bits)
    unsigned long long packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_words[32 * j + lane] = q[j];
        c0[j] = (uint32_t)(__popc(q[j].x) + __popc(q[j].y));
        c1[j] = (uint32_t)(__popc(q[j].z) + __popc(q[j].w));
        packed |= (unsigned long long)(c0[j] + c1[j]) << (16 * j);
    }
    unsigned long long inc = packed;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        unsigned long long o = __shfl_up_sync(0xffffffffu, inc, s);
        if (lane >= s) inc += o;
    }
    const unsigned long long ex = inc - packed;
    const unsigned long long tot = __shfl_sync(0xffffffffu, inc, 31);
    uint32_t base = 0;              // nodes of the tile in groups before j
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t before = base + (uint32_t)((ex >> (16 * j)) & 0xffffu);  // set bits of the tile before this lane's pair
        base += (uint32_t)((tot >> (16 * j)) & 0xffffu);
        // words 2*(32 j + lane), +1: the even lane of each pair starts a 256-bit block
        if ((lane & 1) == 0) pre256[tile * 64 + j * 16 + (lane >> 1)] = t0 + before;
        s_incl[32 * j + lane] = (before + c0[j]) | ((before + c0[j] + c1[j]) << 16);  // two 16-bit inclusive counts
    }
    return total;
}

#define SKB_EMIT_WARPS (SKB_EMIT_THREADS / 32)

template <bool VALUES>  // false: boolean image, no node values (csr.py:554-562) and no dtype dispatch in the kernel
__global__ void __launch_bounds__(SKB_EMIT_THREADS) skb_emit_nodes_kernel(
    const uint64_t* __restrict__ bits, const uint32_t* __restrict__ tile_prefix, int64_t ntiles, SkbGeom g,
    const void* img, int dtype, int64_t z_origin, int64_t* __restrict__ lin, int64_t* __restrict__ coords,
    double* __restrict__ values_arg, uint32_t* __restrict__ pre256, uint32_t node_cap) {
    // node_cap: capacity of lin / coords / values (the runner may size them by a hint before the count has
    // arrived); nodes with id >= node_cap are not written, the rank structure always is
    SKB_PDL_ENTRY();
    double* __restrict__ values = VALUES ? values_arg : nullptr;
    __shared__ __align__(16) uint4 s_words[SKB_EMIT_WARPS][2][SKB_TILE_WORDS / 2];
    __shared__ uint32_t s_incl[SKB_EMIT_WARPS][2][SKB_TILE_WORDS / 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps = (int64_t)gridDim.x * SKB_EMIT_WARPS;
    const int64_t z_add = g.ndim == 3 ? z_origin : 0;
    // Coordinates of the tile origin: divided out once per warp (64-bit divisions when the volume has more than
    // 2^32 voxels), then advanced by the constant stride of the loop with two carries per tile.
    int64_t tile = (int64_t)blockIdx.x * SKB_EMIT_WARPS + warp;
    int32_t za = 0, ya = 0, xa = 0, dzt, dy0, dx0;
    if (tile < ntiles) skb_coords(g, tile * SKB_TILE_VOX, za, ya, xa);
    skb_coords(g, warps * SKB_TILE_VOX, dzt, dy0, dx0);  // may exceed the volume: only a displacement
    // The kernel is bound by the instructions it issues (ncu, round 2: 71 % issue slots busy, 524 instructions per tile
    // when every lane picked the set bits out of its own words in divergent loops, 18 % of them branches), and a sparse
    // tile holds fewer nodes (~14 at 2048^3) than a warp has lanes.  So: the words and their running counts go to
    // shared memory, TWO tiles per iteration (this one and the one a grid stride further), and then one loop over the
    // nodes of both, one node per lane -- a lane finds its node's word by binary search in the running counts and the
    // bit inside the word by clearing lower set bits.  The warp's next pair is pulled into L2 meanwhile.
    for (; tile < ntiles; tile += 2 * warps) {
        const int64_t tile_b = tile + warps;
        const bool has_b = tile_b < ntiles;
        const uint32_t ta = tile_prefix[tile], na_all = tile_prefix[tile + 1] - ta;
        const uint32_t tb = has_b ? tile_prefix[tile_b] : 0u, nb_all = has_b ? tile_prefix[tile_b + 1] - tb : 0u;
#pragma unroll
        for (int k = 2; k < 4; ++k) {
            const int64_t tp = tile + k * warps;
            if (tp < ntiles) {
                if (lane < 16)  // 16 lines of 128 bytes
                    asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)(bits + tp * SKB_TILE_WORDS) + lane * 128));
                else if (lane == 16)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(tile_prefix + tp));
            }
        }
        int32_t zb = za, yb = ya, xb = xa;
        skb_advance_coords(g, zb, yb, xb, dzt, dy0, dx0);
        const uint32_t na = skb_rank_tile(tile, ta, na_all, lane, bits, s_words[warp][0], s_incl[warp][0], pre256);
        const uint32_t nbs = has_b ? skb_rank_tile(tile_b, tb, nb_all, lane, bits, s_words[warp][1], s_incl[warp][1], pre256) : 0u;
        __syncwarp();
        for (uint32_t i = lane; i < na + nbs; i += 32) {
            const int second = i >= na ? 1 : 0;
            const uint32_t k = second ? i - na : i;  // the tile's k-th set bit
            const uint16_t* incl = (const uint16_t*)s_incl[warp][second];
            uint32_t pos = 0;  // first word whose inclusive count exceeds k
#pragma unroll
            for (uint32_t step = SKB_TILE_WORDS / 2; step; step >>= 1)
                if ((uint32_t)incl[pos + step - 1] <= k) pos += step;
            uint32_t skip = k - (pos ? (uint32_t)incl[pos - 1] : 0u);  // set bits of that word below ours
            uint64_t word = ((const uint64_t*)s_words[warp][second])[pos];
            for (; skip; --skip) word &= word - 1;
            const uint32_t off = pos * 64u + (uint32_t)(__ffsll((long long)word) - 1);
            const uint32_t id = (second ? tb : ta) + k;
            if (id < node_cap)
                skb_emit_one(off, (second ? tile_b : tile) * SKB_TILE_VOX, second ? zb : za, z_add, second ? yb : ya,
                             second ? xb : xa, id, g, img, dtype, lin, coords, values);
        }
        __syncwarp();
        za = zb;
        ya = yb;
        xa = xb;
        skb_advance_coords(g, za, ya, xa, dzt, dy0, dx0);
    }
}

// ------------------------------------------------------------------------------- fused sweep
// Mask sweep + device-wide scan + node emission in ONE pass over the image (variant 2, the default of
// the native runner).  The image is the only V-proportional array; the three-kernel path above reads
// it once and then reads the V/8-byte mask again to emit the nodes.  Here a CTA is warp-specialised:
//   * warps 2..9 (256 "converter" threads) stream 64 Ki-voxel chunks through the TMA ring exactly like
//     skb_pack_tma_kernel and assemble each chunk's 1024 mask words in shared memory (double-buffered);
//   * warps 0 and 1 ("scanners", one per buffer) take a finished chunk: write its words with coalesced
//     64-bit stores, count them, publish the count, obtain the number of foreground voxels before the
//     chunk by decoupled look-back over one 64-bit status word per chunk, then write the rank structure
//     and emit the chunk's nodes.
// The look-back polls L2 behind ~20 MB of bulk loads in flight (microseconds per round trip), which is
// why it runs beside the stream instead of inside it, and why it inspects 512 predecessors per round.
// Chunks are handed out by a ticket counter, so every predecessor a scanner may wait for belongs to a
// CTA that is already running.  Node arrays have a capacity (the count is only known afterwards):
// nodes with id >= node_cap are counted but not written, and the caller falls back to skb_emit_nodes.
#define SKB_CHUNK_VOX 65536
#define SKB_CHUNK_WORDS 1024
#define SKB_FUSED_THREADS 320
#define SKB_FUSED_ROWS 8   // look-back: rows of 32 chunk statuses per round
#define SKB_FUSED_STAGE_CAP 1536  // nodes of one chunk compacted in shared memory before emission (2.3 % density)

// Every wait of the fused sweep is bounded (~2 s): a protocol error must end the kernel with a report
// (n_out[1] = number of timed-out waits, details behind the chunk statuses), not hang the GPU.
#define SKB_WAIT_LIMIT_CYCLES 4000000000ll
// A waiting warp must not compete for issue slots with the scanner warps of its SM: the wait suspends the
// thread in hardware (it resumes when the phase completes, or after this many nanoseconds to look again).
#ifndef SKB_SUSPEND_HINT_NS
#define SKB_SUSPEND_HINT_NS 20000u
#endif
__device__ __forceinline__ bool skb_mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(skb_smem_u32(bar)), "r"(parity), "r"(SKB_SUSPEND_HINT_NS)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ void skb_sweep_report(unsigned long long* dbg, int32_t* n_out, int site, long long a, long long b,
                                              long long c) {
    const unsigned long long k = atomicAdd(dbg, 1ull);
    if (k < 15) {
        dbg[1 + 4 * k] = ((unsigned long long)site << 32) | blockIdx.x;
        dbg[2 + 4 * k] = (unsigned long long)a;
        dbg[3 + 4 * k] = (unsigned long long)b;
        dbg[4 + 4 * k] = (unsigned long long)c;
    }
    atomicAdd(n_out + 1, 1);
}
// false = timed out, or another wait of this launch has (n_out[1] != 0): give up quickly
__device__ __forceinline__ bool skb_mbar_wait_bounded(uint64_t* bar, uint32_t parity, const int32_t* n_out) {
    if (skb_mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    long long next_check = 1ll << 21;  // the abort flag is one global address: poll it once per millisecond at most
    for (;;) {
        for (int i = 0; i < 8; ++i)
            if (skb_mbar_try(bar, parity)) return true;
        const long long dt = clock64() - t0;
        if (dt > next_check) {
            if (dt > SKB_WAIT_LIMIT_CYCLES || *(volatile const int32_t*)(n_out + 1) != 0) return false;
            next_check = dt + (1ll << 21);
        }
    }
}
// barrier 1 over the 256 converter threads, returning the OR of `flag`
__device__ __forceinline__ bool skb_converters_sync_or(bool flag) {
    uint32_t out;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "bar.red.or.pred q, 1, 256, p;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(out)
        : "r"((uint32_t)flag)
        : "memory");
    return out != 0;
}

__device__ __forceinline__ void skb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(skb_smem_u32(bar)) : "memory");
}

#ifdef SKB_SWEEP_PROFILE
#define SKB_PROF_T(var) const long long var = clock64()
#define SKB_PROF_ADD(slot, cycles) \
    do { if (blockIdx.x == SKB_SWEEP_PROFILE) atomicAdd(dbg + 64 + (slot), (unsigned long long)(cycles)); } while (0)
#else
#define SKB_PROF_T(var)
#define SKB_PROF_ADD(slot, cycles)
#endif

template <int ESIZE, int STAGES>
__global__ void __launch_bounds__(SKB_FUSED_THREADS, 3) skb_sweep_nodes_kernel(
    const uint8_t* __restrict__ img, int64_t nvox, bool is_float, uint64_t* __restrict__ bits, int64_t nwords,
    int64_t nchunks, SkbGeom g, int dtype, int64_t z_origin, int64_t node_cap, int64_t* __restrict__ lin,
    int64_t* __restrict__ coords, double* __restrict__ values, uint32_t* __restrict__ pre256,
    unsigned long long* status /* [0] = ticket counter, [1 + c] = chunk c */, int32_t* n_out) {
    constexpr int EPV = 16 / ESIZE;                        // voxels per 16-byte piece
    constexpr int NBITS = 64 / ESIZE;                      // voxels per converter thread and stage
    constexpr int SV = SKB_STAGE_BYTES / ESIZE;            // voxels per stage
    constexpr int SPC = SKB_CHUNK_VOX / SV;                // stages per chunk
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t ready_bar[2];         // converters -> scanner: chunk buffer b is complete
    __shared__ __align__(8) uint64_t free_bar[2];          // scanner -> converters: chunk buffer b may be overwritten
    // word w of a chunk lives in slot w + (w >> 5): rows of 32 words and runs of 32 words are both conflict-free
    __shared__ __align__(16) uint64_t chunk_bits[2][SKB_CHUNK_WORDS + 32];
    __shared__ long long chunk_ring[4];
    __shared__ int32_t s_org[2][4][3];
    __shared__ uint16_t node_stage[2][SKB_FUSED_STAGE_CAP];  // per scanner: chunk-relative offsets of the set bits, in order
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t nbytes = nvox * ESIZE;
    unsigned long long* ticket = status;
    unsigned long long* chunk_status = status + 1;
    unsigned long long* dbg = status + 1 + nchunks;  // 64 words: timed-out waits (see skb_sweep_report)
    if (tid == 0) {
        for (int i = 0; i < STAGES; ++i) skb_mbar_init(&full_bar[i], 1);
        for (int i = 0; i < 2; ++i) {
            skb_mbar_init(&ready_bar[i], 256);
            skb_mbar_init(&free_bar[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= 2) {
        // ================================ converters (and the TMA producer, their thread 0) ================
        const int ct = tid - 64;
        bool producer_done = false;
        auto issue = [&](int64_t q) {  // converter thread 0 only
            if (producer_done) return;
            const int64_t n = q / SPC;
            const int k = (int)(q % SPC);
            long long c;
            if (k == 0) {
                c = (long long)atomicAdd(ticket, 1ull);
                chunk_ring[n & 3] = c;
                if (c >= nchunks) {
                    chunk_ring[(n + 1) & 3] = c;  // both scanners must see the end
                    producer_done = true;
                    return;
                }
            } else {
                c = chunk_ring[n & 3];
            }
            const int s = (int)(q % STAGES);
            const int64_t off = ((int64_t)c * SPC + k) * (int64_t)SKB_STAGE_BYTES;
            const int64_t avail = nbytes - off;
            const uint32_t bytes = avail >= SKB_STAGE_BYTES ? (uint32_t)SKB_STAGE_BYTES : (uint32_t)(avail > 0 ? (avail & ~15LL) : 0);
            skb_mbar_expect_tx(&full_bar[s], bytes);
            if (bytes) skb_tma_load_1d(smem + (size_t)s * SKB_STAGE_BYTES, img + off, bytes, &full_bar[s]);
        };
        if (ct == 0)
            for (int64_t q = 0; q < STAGES - 1; ++q) issue(q);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        int64_t c = 0;
        bool failed = false;
        for (int64_t q = 0;; ++q) {
            const int64_t n = q / SPC;
            const int k = (int)(q % SPC);
            const int s = (int)(q % STAGES);
            const uint32_t parity = (uint32_t)((q / STAGES) & 1);
            // the stage refilled here was read in iteration q-1; every converter has passed that
            // iteration's barrier before thread 0 gets here
            if (ct == 0) {
                SKB_PROF_T(pf3);
                issue(q + STAGES - 1);
                SKB_PROF_ADD(10, clock64() - pf3);
            }
            if (k == 0) {
                c = chunk_ring[n & 3];
                if (c >= nchunks) {  // no more chunks: release both scanners (local chunks n and n + 1)
                    for (int64_t e = n; e < n + 2; ++e) {
                        if (e >= 2 && !skb_mbar_wait_bounded(&free_bar[e & 1], (uint32_t)(((e >> 1) - 1) & 1), n_out) && ct == 0)
                            skb_sweep_report(dbg, n_out, 2, e, c, q);
                        skb_mbar_arrive(&ready_bar[e & 1]);
                    }
                    break;
                }
                SKB_PROF_T(pf0);
                if (n >= 2 && !skb_mbar_wait_bounded(&free_bar[n & 1], (uint32_t)(((n >> 1) - 1) & 1), n_out)) {
                    failed = true;
                    if (ct == 0) skb_sweep_report(dbg, n_out, 1, n, c, q);
                }
                if (ct == 0) { SKB_PROF_ADD(1, clock64() - pf0); SKB_PROF_ADD(8, 1); }
                if (ct < 4) {  // coordinates of the origins of the chunk's four 16384-voxel tiles
                    const int64_t r0 = (c * 4 + ct) * (int64_t)SKB_TILE_VOX;
                    int32_t z = 0, y = 0, x = 0;
                    if (r0 < nvox) skb_coords(g, r0, z, y, x);
                    s_org[n & 1][ct][0] = z;
                    s_org[n & 1][ct][1] = y;
                    s_org[n & 1][ct][2] = x;
                }
            }
            SKB_PROF_T(pf1);
            if (!failed && !skb_mbar_wait_bounded(&full_bar[s], parity, n_out)) {
                failed = true;
                if (ct == 0) skb_sweep_report(dbg, n_out, 3, n, c, q);
            }
            if (ct == 0) SKB_PROF_ADD(0, clock64() - pf1);
            const uint8_t* st = smem + (size_t)s * SKB_STAGE_BYTES;
            const int64_t off = ((int64_t)c * SPC + k) * (int64_t)SKB_STAGE_BYTES;
            const int64_t copied = (nbytes - off >= SKB_STAGE_BYTES) ? SKB_STAGE_BYTES : ((nbytes - off) > 0 ? ((nbytes - off) & ~15LL) : 0);
            uint64_t word = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int jj = (j + (lane >> 1)) & 3;  // rotation: the 8 lanes of a quarter-warp hit 32 distinct banks
                const int boff = ct * 64 + jj * 16;
                uint4 v;
                if (boff + 16 <= copied) v = *(const uint4*)(st + boff);
                else if (off + boff < nbytes) v = skb_load_tail<ESIZE>(img, off + boff, nbytes);
                else v = make_uint4(0, 0, 0, 0);
                word |= (uint64_t)skb_nonzero_bits<ESIZE>(v, is_float) << (jj * EPV);
            }
            {   // this thread's NBITS mask bits of the chunk
                const int bit0 = k * SV + ct * NBITS;
                uint8_t* cb = (uint8_t*)chunk_bits[n & 1] + (bit0 >> 11) * 8 + (bit0 >> 3);
                if (NBITS == 64) *(uint64_t*)cb = word;
                else if (NBITS == 32) *(uint32_t*)cb = (uint32_t)word;
                else if (NBITS == 16) *(uint16_t*)cb = (uint16_t)word;
                else *cb = (uint8_t)word;
            }
            if (k == SPC - 1) skb_mbar_arrive(&ready_bar[n & 1]);  // release: the chunk's words are complete
            SKB_PROF_T(pf2);
            if (skb_converters_sync_or(failed)) break;              // stage s fully read (or a wait timed out: stop)
            if (ct == 0) SKB_PROF_ADD(9, clock64() - pf2);
        }
        return;
    }

    // ==================================== scanner of chunk buffer `warp` ====================================
    for (int64_t n = warp;; n += 2) {
        const int b = warp;
        SKB_PROF_T(ps0);
        if (!__all_sync(0xffffffffu, skb_mbar_wait_bounded(&ready_bar[b], (uint32_t)((n >> 1) & 1), n_out))) {
            if (lane == 0) skb_sweep_report(dbg, n_out, 4, n, chunk_ring[n & 3], b);
            break;
        }
        const int64_t c = chunk_ring[n & 3];
        if (c >= nchunks) break;
        SKB_PROF_T(ps1);
        if (lane == 0) SKB_PROF_ADD(2, ps1 - ps0);
        const uint64_t* cb = chunk_bits[b];
        // pass 1: the words go to the global mask (coalesced), the chunk's count is published
        uint32_t cnt = 0;
        const int64_t wbase = c * SKB_CHUNK_WORDS + lane;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            const uint64_t w = cb[r * 33 + lane];
            if (wbase + r * 32 < nwords) bits[wbase + r * 32] = w;
        }
        // this lane owns words 32 lane .. 32 lane + 31 (slots 33 lane + i): their count, then one warp scan per chunk
#pragma unroll 8
        for (int i = 0; i < 32; ++i) cnt += (uint32_t)__popcll(cb[33 * lane + i]);
        uint32_t lane_incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, lane_incl, d);
            if (lane >= d) lane_incl += o;
        }
        const uint32_t chunk_total = __shfl_sync(0xffffffffu, lane_incl, 31);
        uint32_t excl = 0;
        SKB_PROF_T(ps2);
        if (lane == 0) SKB_PROF_ADD(6, ps2 - ps1);
        if (c == 0) {
            if (lane == 0) __stcg(chunk_status, (SKB_ST_INCLUSIVE << 32) | (unsigned long long)chunk_total);
        } else {
            if (lane == 0) __stcg(chunk_status + c, (SKB_ST_AGGREGATE << 32) | (unsigned long long)chunk_total);
            // SKB_FUSED_ROWS x 32 predecessors per round, nearest first: row r = chunks look - 32 r - lane, one
            // coalesced load per row.  Only rows that still hold an unpublished status are loaded again,
            // after a short sleep: ~900 warps poll the same few lines of L2.
            int64_t look = c - 1;
            const long long t0 = clock64();
            long long next_check = 1ll << 21;
            for (;;) {
                unsigned long long stw[SKB_FUSED_ROWS];
                unsigned reload = (1u << SKB_FUSED_ROWS) - 1u;
                uint32_t val;
                bool found;
                for (;;) {
                    if (lane == 0) SKB_PROF_ADD(4, 1);
#pragma unroll
                    for (int r = 0; r < SKB_FUSED_ROWS; ++r) {
                        if ((reload >> r) & 1u) {
                            const int64_t idx = look - (r * 32 + lane);
                            stw[r] = idx >= 0 ? skb_poll_u64(chunk_status + idx) : (SKB_ST_INCLUSIVE << 32);
                        }
                    }
                    found = false;
                    bool blocked = false;
                    val = 0;
                    reload = 0;
#pragma unroll
                    for (int r = 0; r < SKB_FUSED_ROWS; ++r) {
                        if (!found && !blocked) {  // warp-uniform
                            const unsigned incl = __ballot_sync(0xffffffffu, (stw[r] >> 32) == SKB_ST_INCLUSIVE);
                            const unsigned zero = __ballot_sync(0xffffffffu, (stw[r] >> 32) == 0ull);
                            const int stop = incl ? (__ffs((int)incl) - 1) : 31;
                            const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
                            if (zero & need) {  // something nearer than the first inclusive status is unpublished
                                blocked = true;
                                reload = 1u << r;
                            } else {
                                if (lane <= stop) val += (uint32_t)(stw[r] & 0xffffffffull);
                                found = incl != 0u;
                            }
                        }
                    }
                    if (!blocked) break;
                    __nanosleep(200);
                    const long long dt = clock64() - t0;
                    bool give_up = false;
                    if (dt > next_check) {  // once per millisecond at most
                        give_up = dt > SKB_WAIT_LIMIT_CYCLES || *(volatile int32_t*)(n_out + 1) != 0;
                        next_check = dt + (1ll << 21);
                    }
                    if (__any_sync(0xffffffffu, give_up)) {
                        if (lane == 0) skb_sweep_report(dbg, n_out, 5, c, look, reload);
                        val = 0;
                        found = true;  // give up: the count is wrong and the report says so
                        break;
                    }
                }
                excl += __reduce_add_sync(0xffffffffu, val);
                if (lane == 0) SKB_PROF_ADD(5, 1);
                if (found) break;
                look -= 32 * SKB_FUSED_ROWS;
            }
            if (lane == 0) __stcg(chunk_status + c, (SKB_ST_INCLUSIVE << 32) | (unsigned long long)(excl + chunk_total));
        }
        if (lane == 0 && c == nchunks - 1) *n_out = (int32_t)(excl + chunk_total);
        SKB_PROF_T(ps3);
        if (lane == 0) SKB_PROF_ADD(3, ps3 - ps2);
        // pass 2: rank structure and nodes; every lane walks its own 32 words (8 blocks of 256 voxels)
        if (chunk_total == 0) {
            if (c * 256 + 256 <= nwords / 4) {
#pragma unroll
                for (int r = 0; r < 8; ++r) pre256[c * 256 + r * 32 + lane] = excl;
            } else {
                for (int r = 0; r < 8; ++r)
                    if (c * 256 + r * 32 + lane < nwords / 4) pre256[c * 256 + r * 32 + lane] = excl;
            }
        } else {
            uint32_t id = excl + lane_incl - cnt;
            const int64_t blk0 = c * 256 + lane * 8;
            const bool full = (blk0 + 8) * 4 <= nwords;
            if (chunk_total <= SKB_FUSED_STAGE_CAP) {
                // compact the set bits (this lane knows the rank of each of its own) so that the emission
                // below runs one node per lane with coalesced stores and few divisions
                uint16_t* stg = node_stage[b];
                uint32_t k = lane_incl - cnt;
#pragma unroll 4
                for (int i = 0; i < 32; ++i) {
                    if ((i & 3) == 0 && (full || (blk0 + (i >> 2)) * 4 < nwords)) pre256[blk0 + (i >> 2)] = excl + k;
                    uint64_t w = cb[33 * lane + i];
                    const uint32_t woff = (uint32_t)((lane * 32 + i) * 64);
                    while (w) {
                        stg[k++] = (uint16_t)(woff + (uint32_t)(__ffsll((long long)w) - 1));
                        w &= w - 1;
                    }
                }
                __syncwarp();
                for (uint32_t i = lane; i < chunk_total; i += 32) {
                    if ((int64_t)(excl + i) >= node_cap) break;
                    const uint32_t o = stg[i];
                    const int j = (int)(o >> 14);  // tile of the chunk
                    skb_emit_one(o & 16383u, (c * 4 + j) * (int64_t)SKB_TILE_VOX, s_org[b][j][0],
                                 (g.ndim == 3 ? z_origin : 0), s_org[b][j][1], s_org[b][j][2], excl + i, g, img, dtype, lin,
                                 coords, values);
                }
            } else {  // dense chunk: every lane walks its own words
                const int j = lane >> 3;  // tile of the chunk
                const int64_t r0 = (c * 4 + j) * (int64_t)SKB_TILE_VOX;
                const int64_t z_add = (g.ndim == 3 ? z_origin : 0);
                const int32_t zt = s_org[b][j][0], y0 = s_org[b][j][1], x0 = s_org[b][j][2];
#pragma unroll 4
                for (int i = 0; i < 32; ++i) {
                    if ((i & 3) == 0 && (full || (blk0 + (i >> 2)) * 4 < nwords)) pre256[blk0 + (i >> 2)] = id;
                    uint64_t w = cb[33 * lane + i];
                    const uint32_t woff = (uint32_t)(((lane & 7) * 32 + i) * 64);
                    while (w) {
                        if ((int64_t)id < node_cap)
                            skb_emit_one(woff + (uint32_t)(__ffsll((long long)w) - 1), r0, zt, z_add, y0, x0, id, g, img, dtype,
                                         lin, coords, values);
                        ++id;
                        w &= w - 1;
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) { SKB_PROF_ADD(7, clock64() - ps3); SKB_PROF_ADD(11, 1); }
        if (lane == 0) skb_mbar_arrive(&free_bar[b]);
    }
}

template <int ESIZE>
static int skb_sweep_nodes_launch(const void* img, int64_t nvox, bool is_float, uint64_t* bits, int64_t ntiles,
                                  SkbGeom g, int dtype, int64_t z_origin, int64_t node_cap, int64_t* lin,
                                  int64_t* coords, double* values, uint32_t* pre256, void* scratch, int32_t* n_out,
                                  skb_stream_t s) {
    constexpr int STAGES = 3;
    const size_t smem = (size_t)STAGES * SKB_STAGE_BYTES;
    static uint8_t attr_set[SKB_MAX_DEV] = {};
    SKB_TRY(skb_optin_smem(skb_sweep_nodes_kernel<ESIZE, STAGES>, smem, attr_set));
    const int64_t nchunks = (nvox + SKB_CHUNK_VOX - 1) / SKB_CHUNK_VOX;
    int64_t grid = (int64_t)skb_sm_count() * 3;
    if (grid > nchunks) grid = nchunks;
    SKB_CUDA(cudaMemsetAsync(scratch, 0, (size_t)(nchunks + 1 + 128) * 8, s));
    SKB_CUDA(cudaMemsetAsync(n_out, 0, 8, s));
    skb_sweep_nodes_kernel<ESIZE, STAGES><<<(unsigned)grid, SKB_FUSED_THREADS, smem, s>>>(
        (const uint8_t*)img, nvox, is_float, bits, ntiles * SKB_TILE_WORDS, nchunks, g, dtype, z_origin, node_cap, lin,
        coords, values, pre256, (unsigned long long*)scratch, n_out);
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(cudaGetLastError(), "sweep_nodes launch");
}

// ------------------------------------------------------------------------------- fused round loops
// The junction MST and the two list rankings are loops of small kernels whose trip count depends on
// the data; launched one kernel per phase they cost a host synchronisation per round.  Here each
// loop is ONE cooperative launch (grid = what is co-resident) with grid-wide barriers between the
// phases and the convergence test done on the device.  The phase bodies are the same functors.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define SKB_HAVE_FUSED_LOOPS 1
#define SKB_COOP_THREADS 256

#define SKB_GRID_FOR(i, n) \
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

__global__ void __launch_bounds__(SKB_COOP_THREADS) skb_mst_coop_kernel(
    const int32_t* ea, const int32_t* eb, const uint64_t* ew, int64_t nej, int32_t* par, uint64_t* bestw,
    uint32_t* beste, int32_t* era, int32_t* erb, uint8_t* tree, int32_t* flags /* [3]: two flags + rounds */) {
    cg::grid_group grid = cg::this_grid();
    const bool first = blockIdx.x == 0 && threadIdx.x == 0;
    SkbMstInitItem init{ea, eb, par, bestw, beste};
    SKB_GRID_FOR(e, nej) {
        init(e);
        era[e] = ea[e] < 0 ? -2 : 0;  // padding rows of a gathered edge list are dead from the start
        tree[e] = 0;
    }
    if (first) flags[0] = flags[1] = 0;
    grid.sync();
    int round = 0;
    for (; round < 64; ++round) {
        int32_t* f = flags + (round & 1);
        SkbMstMinWItem p1{ea, eb, ew, par, era, erb, bestw, f};
        SKB_GRID_FOR(e, nej) p1(e);
        grid.sync();
        if (*(volatile int32_t*)f == 0) break;  // same value for every thread of the grid
        SkbMstMinEItem p2{ew, era, erb, bestw, beste};
        SKB_GRID_FOR(e, nej) p2(e);
        grid.sync();
        SkbMstHookItem p3{era, erb, beste, par, tree};
        SKB_GRID_FOR(e, nej) p3(e);
        grid.sync();
        SkbMstResetItem p4{era, erb, bestw, beste};
        SKB_GRID_FOR(e, nej) p4(e);
        if (first) flags[(round + 1) & 1] = 0;
        grid.sync();
    }
    if (first) flags[2] = round;
}

// Junction MST, table weights (every weight is one of the 26 neighbour distances: anything but height mode).  The order
// (weight, a, b) of an edge is then ONE 64-bit key: (rank of the weight among the table's values, a, direction code k27
// of b seen from a) -- for a fixed a ascending direction codes are ascending neighbour ids -- so the edges may be listed
// in ANY order (skb_junction_append_kernel appends them with one atomic per thread block), a round needs one atomic
// minimum per end instead of two phases, and the
// per-component minima are double-buffered by round parity (a round resets the entries the next round will use: the
// roots of round r+1 are among the roots round r saw), so there is no reset phase either: 2 grid barriers per round
// instead of 4.  After the last round the same launch clears the dropped junction-junction edges from `keep`
// (csr.py:1018-1019; keep = the raw masks, updated in place).
__device__ __forceinline__ unsigned long long skb_mst_key(const uint32_t* s_rank, int32_t a, uint8_t k) {
    return ((unsigned long long)s_rank[k] << 40) | ((unsigned long long)(uint32_t)a << 5) | (unsigned long long)k;
}

__global__ void __launch_bounds__(SKB_COOP_THREADS) skb_mst_keys_coop_kernel(
    const int32_t* ea, const int32_t* eb, const uint8_t* ek, const double* wtab, int64_t nej_arg, const int32_t* nej_ptr,
    int64_t nej_cap, int32_t* par,
    unsigned long long* best /* [2][n_ids] */, int64_t n_ids, int32_t* era, int32_t* erb, uint8_t* tree,
    int32_t* flags /* [3]: two flags + rounds */, uint32_t* keep, int32_t id_shift, int32_t n_local,
    const int32_t* jid, int32_t* eda, int32_t* edb, const int32_t* j_ptr) {
    // jid != NULL: par / best are indexed by DENSE ids of the junction voxels (jid[node], assigned by
    // skb_junction_append_kernel; *j_ptr of them, n_ids = the capacity of the arrays) instead of node ids.  On an image
    // with many junction clusters (30 M nodes, 2.65 M junction edges: a stack of 2-D images) the node-indexed arrays span
    // 600 MB and every find / atomic is a DRAM sector; the dense ones stay in L2.
    // more junction voxels than the arrays hold: nothing is done (the host reads the same count and goes the other way);
    // the same decision in every thread of the grid -- nobody writes the count during this launch -- before the first barrier
    if (jid && (int64_t)*j_ptr > n_ids) return;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t s_rank[27];
    // the number of edges: a launch argument, or (edges appended with an atomic counter, count still on its way to the
    // host) read from device memory and clamped to the capacity of the arrays
    int64_t nej = nej_arg;
    if (nej_ptr) {
        nej = *nej_ptr;
        if (nej > nej_cap) nej = nej_cap;
    }
    if (threadIdx.x < 27) {  // rank of every table weight among the table's distinct values (positive doubles: bit order)
        const unsigned long long me = (unsigned long long)__double_as_longlong(wtab[threadIdx.x]);
        uint32_t r = 0;
        for (int j = 0; j < 27; ++j) {
            const unsigned long long o = (unsigned long long)__double_as_longlong(wtab[j]);
            bool seen = false;
            for (int i = 0; i < j; ++i) seen |= (unsigned long long)__double_as_longlong(wtab[i]) == o;
            if (!seen && o < me) ++r;
        }
        s_rank[threadIdx.x] = r;
    }
    const bool first = blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long* best0 = best;
    unsigned long long* best1 = best + n_ids;
    SKB_GRID_FOR(e, nej) {
        const int32_t a = ea[e], b = eb[e];
        tree[e] = 0;
        if (a < 0) {  // padding row of a gathered edge list: dead from the start
            era[e] = -2;
            continue;
        }
        era[e] = 0;
        int32_t xa = a, xb = b;
        if (jid) {
            xa = jid[a];
            xb = jid[b];
            eda[e] = xa;
            edb[e] = xb;
        }
        par[xa] = xa;
        par[xb] = xb;
        best0[xa] = best0[xb] = best1[xa] = best1[xb] = ~0ull;
    }
    if (first) flags[0] = flags[1] = 0;
    __syncthreads();
    grid.sync();
    int round = 0;
    for (; round < 64; ++round) {
        unsigned long long* cur = (round & 1) ? best1 : best0;
        unsigned long long* nxt = (round & 1) ? best0 : best1;
        int32_t* f = flags + (round & 1);
        SKB_GRID_FOR(e, nej) {
            if (era[e] == -2) continue;  // already inside one component
            const int32_t ra = skb_find(par, jid ? eda[e] : ea[e]), rb = skb_find(par, jid ? edb[e] : eb[e]);
            if (ra == rb) {
                era[e] = -2;
                continue;
            }
            era[e] = ra;
            erb[e] = rb;
            const unsigned long long key = skb_mst_key(s_rank, ea[e], ek[e]);
            atomicMin(cur + ra, key);
            atomicMin(cur + rb, key);
            nxt[ra] = ~0ull;
            nxt[rb] = ~0ull;
            *f = 1;
        }
        grid.sync();
        if (*(volatile int32_t*)f == 0) break;  // same value for every thread of the grid
        if (first) flags[(round + 1) & 1] = 0;   // last read before the previous barrier, next written after the next one
        SKB_GRID_FOR(e, nej) {
            const int32_t ra = era[e];
            if (ra == -2) continue;
            const int32_t rb = erb[e];
            const unsigned long long key = skb_mst_key(s_rank, ea[e], ek[e]);
            const bool pa = __ldcg(cur + ra) == key, pb = __ldcg(cur + rb) == key;
            if (!(pa || pb)) continue;
            tree[e] = 1;
            // a component's root is written by exactly one edge (its chosen one); a mutual choice keeps the smaller
            // root so that the two hooks do not form a 2-cycle
            if (pa && !(pb && ra < rb)) skb_st_cg(par + ra, rb);
            if (pb && !(pa && rb < ra)) skb_st_cg(par + rb, ra);
        }
        grid.sync();
    }
    if (first) flags[2] = round;
    if (keep) {  // every tree flag was written before the barrier that ended the loop
        SkbMstApplyItem apply{ea, eb, ek, tree, id_shift, n_local, keep};
        SKB_GRID_FOR(e, nej) apply(e);
    }
}

template <typename Jump>
__device__ __forceinline__ void skb_rank_rounds(cg::grid_group& grid, Jump jump, skb_i4* b0, skb_i4* b1, SkbAux* a0,
                                                SkbAux* a1, int64_t n, int32_t* flags /* [4]: 3 flags + rounds */) {
    const bool first = blockIdx.x == 0 && threadIdx.x == 0;
    if (first) flags[0] = flags[1] = flags[2] = 0;
    grid.sync();
    int cur = 0, round = 0;
    for (; round < 64; ++round) {
        // three rotating flags: the one cleared during round r was last READ after the barrier of
        // round r-2, and every thread has passed that read before it can arrive at the barrier of r-1
        int32_t* f = flags + (round % 3);
        jump.src = cur ? b1 : b0;
        jump.dst = cur ? b0 : b1;
        jump.aux_src = cur ? a1 : a0;
        jump.aux_dst = cur ? a0 : a1;
        jump.flag = f;
        SKB_GRID_FOR(s, n) jump(s);
        if (first) flags[(round + 1) % 3] = 0;
        grid.sync();
        cur ^= 1;
        if (*(volatile int32_t*)f == 0) break;  // a round without progress: everything that can end has ended
    }
    if (cur == 1) {  // leave the result in buffer 0
        SKB_GRID_FOR(s, n) {
            skb_st_i4(b0 + s, skb_ld_i4(b1 + s));
            if (a0) a0[s] = a1[s];
        }
    }
    if (first) flags[3] = round + 1;
}

__global__ void __launch_bounds__(SKB_COOP_THREADS) skb_rank_coop_kernel(skb_i4* b0, skb_i4* b1, SkbAux* a0,
                                                                         SkbAux* a1, int64_t n, int32_t lo,
                                                                         int32_t hi, int32_t* flags, const int32_t* n_ptr,
                                                                         int guard) {
    if (guard && g_skb_poison) return;  // every thread of the grid reads the same value
    if (n_ptr && *n_ptr < n) n = *n_ptr;  // the number of starts in device memory
    // flags[3] == 0: the list-following pass (SkbRankWalkItem) finished every list, nothing to do.  Every
    // thread reads the same value: nobody writes flags[3] before the last grid barrier below.
    if (*(volatile int32_t*)(flags + 3) == 0) return;
    cg::grid_group grid = cg::this_grid();
    SkbJumpItem jump{b0, b1, a0, a1, lo, hi, flags};
    skb_rank_rounds(grid, jump, b0, b1, a0, a1, n, flags);
}

__global__ void __launch_bounds__(SKB_COOP_THREADS) skb_table_rank_coop_kernel(skb_i4* b0, skb_i4* b1, SkbAux* a0,
                                                                               SkbAux* a1, int64_t n,
                                                                               SkbSlabTable t, int32_t* flags, int guard) {
    if (guard && g_skb_poison) return;  // every thread of the grid reads the same value
    cg::grid_group grid = cg::this_grid();
    SkbTableJumpItem jump{b0, b1, a0, a1, t, flags};
    skb_rank_rounds(grid, jump, b0, b1, a0, a1, n, flags);
}

// a grid barrier costs more the more CTAs take part: "mst_blocks_per_sm" caps the Boruvka launch (0 = what fits)
static int g_skb_mst_blocks_per_sm = 0;  // measured: fewer CTAs lose (the rounds are bound by the finds, not the barriers)
static int g_skb_emit_blocks_per_sm = 16;  // measured at 2048^3: 5 -> 0.314, 8 -> 0.258, 16 -> 0.254 ms (tail balance)
static int g_skb_arena_guard = 0;  // "arena_guard": guard gaps between the runners' allocations (skb_run.inl)
static int g_skb_small_cycles = 1;  // "small_cycles": 0 = always the general machinery for junction-free cycles
static int g_skb_mst_keys = 1;  // "mst_keys": 0 = the two-phase rounds also for table weights (A/B)
static int64_t g_skb_dense_mst_min_nodes = 12000000;  // "dense_mst_min_nodes": dense junction ids for the Boruvka arrays above
                                                       // this many nodes (skb_run.inl; tests set 0 to force them)
// "long_path_exact": 1 = paths of any length are summed by one thread in the reference's order (bit-exact, O(longest
// path) time); 0 (default) = paths of more than SKB_LONG_PATH entries by one thread block in a fixed tree order
static int32_t g_skb_long_path = SKB_LONG_PATH;
template <typename K>
static int skb_coop_launch(K kernel, int64_t n_items, void** args, skb_stream_t s, int max_per_sm = 0) {
    int per_sm = 0;
    SKB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SKB_COOP_THREADS, 0));
    if (per_sm < 1) return skb_fail("cooperative kernel does not fit");
    if (max_per_sm > 0 && per_sm > max_per_sm) per_sm = max_per_sm;
    int64_t cap = (int64_t)per_sm * skb_sm_count();
    int64_t blocks = (n_items + SKB_COOP_THREADS - 1) / SKB_COOP_THREADS;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    SKB_CUDA(cudaLaunchCooperativeKernel((void*)kernel, dim3((unsigned)blocks), dim3(SKB_COOP_THREADS), args, 0, s));
    SKB_COUNT_LAUNCH(1);
    return 0;
}

// flag must hold 4 ints.  All three return 0: no host round trip, the result is in buffer 0 and
// the round count in flag[3] (flag[2] for the MST).
static int skb_mst_run(const int32_t* ea, const int32_t* eb, const uint64_t* ew, int64_t nej, int32_t* par,
                       uint64_t* bestw, uint32_t* beste, int32_t* era, int32_t* erb, uint8_t* tree, int32_t* flag,
                       skb_stream_t s) {
    void* args[] = {&ea, &eb, &ew, &nej, &par, &bestw, &beste, &era, &erb, &tree, &flag};
    return skb_coop_launch(skb_mst_coop_kernel, nej, args, s);
}

// nej_ptr != NULL: the edge count is read on the device (at most nej edges, the capacity of the arrays)
// dense ids of the junction voxels (single-image runner): see skb_mst_keys_coop_kernel
struct SkbDense {
    int32_t* jid;       // [n nodes], written for every node of raw degree >= 3 by skb_junction_append_kernel
    int32_t* jcounter;  // device: their number
    int32_t* eda;       // [edge capacity] dense ids of every edge's two ends
    int32_t* edb;
    int64_t j_cap;      // capacity of par / best
};
static int skb_mst_keys_run(const int32_t* ea, const int32_t* eb, const uint8_t* ek, const double* wtab, int64_t nej,
                            const int32_t* nej_ptr, int32_t* par, unsigned long long* best, int64_t n_ids, int32_t* era,
                            int32_t* erb, uint8_t* tree, int32_t* flag, uint32_t* keep, int32_t id_shift, int32_t n_local,
                            skb_stream_t s, const SkbDense* dense = nullptr) {
    int64_t nej_cap = nej;
    if (dense && !dense->jid) dense = nullptr;  // (junction voxels counted only: arrays indexed by node id)
    const int32_t* jid = dense ? dense->jid : nullptr;
    int32_t* eda = dense ? dense->eda : nullptr;
    int32_t* edb = dense ? dense->edb : nullptr;
    const int32_t* j_ptr = dense ? dense->jcounter : nullptr;
    if (dense) n_ids = dense->j_cap;
    void* args[] = {&ea, &eb, &ek, &wtab, &nej, &nej_ptr, &nej_cap, &par, &best, &n_ids, &era, &erb, &tree, &flag, &keep,
                    &id_shift, &n_local, &jid, &eda, &edb, &j_ptr};
    return skb_coop_launch(skb_mst_keys_coop_kernel, nej, args, s, g_skb_mst_blocks_per_sm);
}

// Junction-junction edges (csr.py:1004-1016) in ONE pass and in no particular order: every node of raw degree >= 3 tests
// its forward neighbours (one rank query each), a thread block sums its edge count, takes its share of the arrays with
// one atomic and writes.  Replaces count + scan + emit (two passes of rank queries) when the arrays can be sized by a
// capacity; *counter receives the number of edges (edges beyond `cap` are counted, not written).
__global__ void __launch_bounds__(256) skb_junction_append_kernel(SkbGeom g, const uint64_t* __restrict__ bits,
                                                                  const uint32_t* __restrict__ pre256,
                                                                  const int64_t* __restrict__ lin,
                                                                  const uint32_t* __restrict__ nb, int64_t n,
                                                                  int64_t a0, int64_t a1, int32_t* __restrict__ ea,
                                                                  int32_t* __restrict__ eb, uint8_t* __restrict__ ek,
                                                                  uint32_t cap, int32_t* counter, const int32_t* n_ptr,
                                                                  int32_t* __restrict__ jid, int32_t* jcounter) {
    // jid != NULL: every node of raw degree >= 3 also receives a dense id (jid[node]; *jcounter of them, in no
    // particular order: one atomic per thread block), by which the Boruvka launch indexes its arrays
    if (n_ptr) {  // the node count in device memory (at most n, the capacity of the node arrays)
        const int64_t nd = *n_ptr;
        if (nd < n) n = nd;
        if (a1 > n) a1 = n;
    }
    // Only one node in ~12 has raw degree >= 3, and such a node tests 1-3 forward neighbours, each a chain of three
    // dependent loads (block prefix, mask words, the neighbour's raw mask).  Done by the node's own lane, a warp waited for
    // its busiest lane while 29 others idled (ncu, round 2: 14 of 32 lanes active, 0.117 ms at 2048^3, the neighbour lists
    // in local memory).  Here the (node, direction) pairs of a warp are dealt out to its lanes, one pair per lane and round;
    // accepted neighbours are parked in shared memory under their owner, who writes them out after the block's one atomic.
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t s_base, s_jbase;
    __shared__ int32_t s_u[13][256];   // neighbour ids by (direction - 14, owner thread)
    __shared__ uint32_t s_acc[256];    // accepted directions of every owner
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t stride = (int64_t)gridDim.x * 256;
    const int64_t n_round = (n + 255) / 256 * 256;
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n_round; base += stride) {
        const int64_t i = base + tid;
        uint32_t fm = 0;   // forward directions (k27 >= 14) to test
        int64_t r = 0;
        uint32_t isj = 0;  // raw degree >= 3 (dense ids)
        if (i < n && i >= a0 && i < a1) {  // [a0, a1): the nodes whose forward edges this call emits (a Z-slab's own nodes)
            const uint32_t m = nb[i];
            if (__popc(m) >= 3) {
                isj = 1u;
                if (m >> 14) {
                    fm = m & ~((1u << 14) - 1u);
                    r = lin[i];
                }
            }
        }
        s_acc[tid] = 0u;
        const uint32_t nf = (uint32_t)__popc(fm);
        uint32_t pinc = nf;  // inclusive prefix of the pair counts over the warp
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, pinc, d);
            if (lane >= d) pinc += o;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, pinc, 31);
        __syncwarp();
        for (uint32_t t0 = 0; t0 < total; t0 += 32) {   // the same trip count in every lane of the warp
            const uint32_t t = t0 + lane;
            // owner = first lane whose inclusive prefix exceeds t
            int owner = 0;
#pragma unroll
            for (int step = 16; step; step >>= 1) {
                const uint32_t v = __shfl_sync(0xffffffffu, pinc, owner + step - 1);
                if (v <= t) owner += step;
            }
            owner &= 31;
            const uint32_t o_inc = __shfl_sync(0xffffffffu, pinc, owner);
            const uint32_t o_fm = __shfl_sync(0xffffffffu, fm, owner);
            const long long o_r = __shfl_sync(0xffffffffu, (long long)r, owner);
            if (t < total) {
                uint32_t skip = t - (o_inc - (uint32_t)__popc(o_fm));  // the pair's place among the owner's directions
                uint32_t w = o_fm;
                for (; skip; --skip) w &= w - 1u;
                const int k = __ffs((int)w) - 1;
                const int64_t oi = base + (int64_t)(warp * 32 + owner);
                const int32_t u = k == 14 ? (int32_t)oi + 1 : skb_rank(bits, pre256, skb_neighbor_lin(g, (int64_t)o_r, k));
                if (__popc(nb[u]) >= 3) {
                    s_u[k - 14][warp * 32 + owner] = u;
                    atomicOr(&s_acc[warp * 32 + owner], 1u << k);
                }
            }
        }
        __syncwarp();
        const uint32_t acc = s_acc[tid];
        const uint32_t cnt = (uint32_t)__popc(acc);
        // one scan for both counts: edges in the low 16 bits (a warp holds at most 416), junction voxels above
        uint32_t inc = cnt | (isj << 16);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0, tj = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const uint32_t c = warp_tot[w];
                warp_tot[w] = t | (tj << 16);
                t += c & 0xffffu;
                tj += c >> 16;
            }
            s_base = t ? (uint32_t)atomicAdd(counter, (int32_t)t) : 0u;
            s_jbase = (jcounter && tj) ? (uint32_t)atomicAdd(jcounter, (int32_t)tj) : 0u;  // (jid NULL: counted only)
        }
        __syncthreads();
        if (jid && isj) jid[i] = (int32_t)(s_jbase + (warp_tot[warp] >> 16) + (inc >> 16) - 1u);
        uint32_t o = s_base + (warp_tot[warp] & 0xffffu) + (inc & 0xffffu) - cnt;
        for (uint32_t am = acc; am; am &= am - 1u, ++o) {   // ascending directions = ascending neighbour ids, as before
            const int k = __ffs((int)am) - 1;
            if (o < cap) {
                ea[o] = (int32_t)i;
                eb[o] = s_u[k - 14][tid];
                ek[o] = (uint8_t)k;
            }
        }
        __syncthreads();
    }
}

static int g_skb_junction_append = 1;  // "junction_append": 0 = always count + scan + emit
// returns 0, or 1 when the one-pass form does not apply (the caller counts, scans and emits)
static int skb_junction_append(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape, const int64_t* lin,
                               const uint32_t* nb, int64_t n, int64_t a0, int64_t a1, int32_t* ea, int32_t* eb, uint8_t* ek,
                               int64_t cap, int32_t* counter, skb_stream_t s, const int32_t* n_ptr = nullptr,
                               const SkbDense* dense = nullptr);


static int skb_rank_run(skb_i4* b0, skb_i4* b1, SkbAux* a0, SkbAux* a1, int64_t n, int32_t lo, int32_t hi,
                        int32_t* flag, skb_stream_t s, const int32_t* n_ptr = nullptr) {
    int guard = g_skb_guard;
    void* args[] = {&b0, &b1, &a0, &a1, &n, &lo, &hi, &flag, &n_ptr, &guard};
    return skb_coop_launch(skb_rank_coop_kernel, n, args, s);
}

static int skb_table_rank_run(skb_i4* b0, skb_i4* b1, SkbAux* a0, SkbAux* a1, int64_t n, SkbSlabTable t,
                              int32_t* flag, skb_stream_t s) {
    int guard = g_skb_guard;
    void* args[] = {&b0, &b1, &a0, &a1, &n, &t, &flag, &guard};
    return skb_coop_launch(skb_table_rank_coop_kernel, n, args, s);
}

// ------------------------------------------------------------------------------- junction-free cycles, small case
// Junction-free cycles (csr.py:369-386) are rare and short in real skeletons: a few thousand voxels on rings of 4-20.
// The general machinery (sampled list ranking all over again: ~17 launches, several of them passes over all N nodes,
// 0.10-0.15 ms whatever the volume) is kept for large cases; up to SKB_SMALL_CYCLE_NODES uncovered chain voxels are
// handled by ONE thread block in shared memory: sort the voxels, link neighbours, label rings by their minimum
// (union by minimum), rank the ring minima (= path order), walk each ring once in shared memory for the positions,
// then write all entries in parallel.  `list` holds the uncovered degree-2 nodes in any order (skb_foreach_unmarked +
// an atomic counter).
#define SKB_SMALL_CYCLE_NODES 4096
#define SKB_SMALL_CYCLE_THREADS 1024

struct SkbCollectUncoveredItem {  // over nodes that are not covered: the chain voxels this rank owns go to the list
    const skb_i4* rec;
    const uint8_t* covered;
    int32_t* list;
    int32_t* counter;
    int32_t cap, own0, own1;
    const int32_t* n_ptr = nullptr;  // the node count in device memory: nodes beyond it do not exist
    __device__ void operator()(int64_t v) const {
        if (covered[v] || (int32_t)v < own0 || (int32_t)v >= own1) return;
        if (n_ptr && v >= (int64_t)*n_ptr) return;
        if ((skb_ld_i4(rec + 2 * v).w & SKB_REC_DEG_MASK) != 2) return;
        const int32_t k = atomicAdd(counter, 1);
        if (k < cap) list[k] = (int32_t)v;
    }
};

struct SkbSmallCycleSmem {
    int32_t u[SKB_SMALL_CYCLE_NODES];     // node ids, ascending after the sort
    int32_t par[SKB_SMALL_CYCLE_NODES];   // union by minimum, then the ring's minimum (local index) of every voxel
    int32_t cnt[SKB_SMALL_CYCLE_NODES];   // per ring minimum: voxels on the ring
    int32_t base[SKB_SMALL_CYCLE_NODES];  // per ring minimum: entry offset of the ring's path
    int32_t first[SKB_SMALL_CYCLE_NODES]; // per ring minimum: voxels on the rings before it (offset into `order`)
    double wv[SKB_SMALL_CYCLE_NODES];     // weight of the edge arriving at the voxel along the path
    double vv[SKB_SMALL_CYCLE_NODES];     // the voxel's value (1.0 for boolean images)
    uint16_t l0[SKB_SMALL_CYCLE_NODES], l1[SKB_SMALL_CYCLE_NODES];  // local indices of the two neighbours
    uint16_t pos[SKB_SMALL_CYCLE_NODES];  // position on the path (the minimum is entry 0 and the closing entry)
    uint16_t order[SKB_SMALL_CYCLE_NODES];  // voxels in path order, ring after ring
    uint8_t dir[SKB_SMALL_CYCLE_NODES];   // which neighbour precedes the voxel on the path (0: n0, 1: n1)
    uint32_t warp_a[32], warp_b[32];
    int32_t bad;
};

__device__ __forceinline__ int32_t skb_smem_find(int32_t* par, int32_t v) {
    int32_t p = *(volatile int32_t*)(par + v);
    while (p != v) {
        int32_t gp = *(volatile int32_t*)(par + p);
        if (gp != p) *(volatile int32_t*)(par + v) = gp;
        v = p;
        p = gp;
    }
    return v;
}

// what one launch writes: everything (one GPU), or -- in Z-slab mode, where the entry offsets of a rank's paths are only
// known after an exchange -- the path heads first (into arrays of their own, ids from pid_base) and the entries later
enum { SKB_CYC_ALL = 0, SKB_CYC_HEADS = 1, SKB_CYC_ENTRIES = 2 };

struct SkbSmallCycleArgs {
    const int32_t* list;      // uncovered chain voxels (local node ids), any order
    const int32_t* counter;   // their number as counted on the device (must equal m)
    int32_t m, mode;
    const skb_i4* rec;
    const double* values;     // node values or NULL
    int32_t pid_base;         // id of the first ring's path: SKB_CYC_ALL: the number of regular paths
    int32_t id_shift;         // global node id = local id + id_shift (0 on one GPU)
    const int64_t* lin;       // Z-slabs: raveled index of every node (for the branch table), else NULL
    int64_t lin_offset;
    int32_t* start_node;
    uint32_t* plen;
    int32_t* pmin;
    int32_t* pindptr;         // SKB_CYC_ALL: entry offsets, continued from pindptr[pid_base]; SKB_CYC_ENTRIES: the offsets to use
    int32_t* pidx;
    double* pdata;            // NULL: not written (boolean image whose paths.data was filled with ones)
    double* pw;
    uint32_t* is_min;
    SkbAux* head_aux;         // Z-slabs: per path sums and end node, else NULL
    int32_t* out;             // [0] rings, [1] entries, [2] error
    const int32_t* m_ptr = nullptr;         // the number of voxels in device memory (overrides m; more than the list holds:
                                            // error 2, nothing written)
    const int32_t* pid_base_ptr = nullptr;  // the number of regular paths in device memory (overrides pid_base)
    int guard = 0;                          // launched guarded (g_skb_poison)
};

__global__ void __launch_bounds__(SKB_SMALL_CYCLE_THREADS) skb_small_cycles_kernel(const SkbSmallCycleArgs A) {
    SKB_PDL_ENTRY();
    if (A.guard && g_skb_poison) return;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    SkbSmallCycleSmem& S = *(SkbSmallCycleSmem*)smem_raw;
    constexpr int PER = SKB_SMALL_CYCLE_NODES / SKB_SMALL_CYCLE_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int m = A.m;
    if (A.m_ptr) {
        const int32_t md = *A.m_ptr;
        if (md > SKB_SMALL_CYCLE_NODES || md < 0) {  // not a small case: the caller goes the general way
            if (tid == 0) {
                A.out[0] = 0;
                A.out[1] = 0;
                A.out[2] = 2;
            }
            return;
        }
        m = md;
    }
    const int32_t pid_base = A.pid_base_ptr ? *A.pid_base_ptr : A.pid_base;
    const skb_i4* __restrict__ rec = A.rec;
    int n2 = 32;
    while (n2 < m) n2 <<= 1;
    if (tid == 0) S.bad = (*A.counter != m) ? 1 : 0;
    if (m == 0) {  // (device-resident count) nothing to do
        if (tid == 0) {
            A.out[0] = 0;
            A.out[1] = 0;
            A.out[2] = S.bad;
        }
        return;
    }
    for (int i = tid; i < n2; i += SKB_SMALL_CYCLE_THREADS) S.u[i] = i < m ? A.list[i] : 0x7fffffff;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < n2; t += SKB_SMALL_CYCLE_THREADS) {
                const int x = t ^ j;
                if (x > t) {
                    const int32_t a = S.u[t], b = S.u[x];
                    if ((a > b) == ((t & k) == 0)) {
                        S.u[t] = b;
                        S.u[x] = a;
                    }
                }
            }
            __syncthreads();
        }
    // neighbours as local indices
    for (int i = tid; i < m; i += SKB_SMALL_CYCLE_THREADS) {
        const skb_i4 r = skb_ld_i4(rec + 2 * (int64_t)S.u[i]);
        int32_t loc[2];
        const int32_t nbr[2] = {r.x, r.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int lo = 0, hi = m;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (S.u[mid] < nbr[h]) lo = mid + 1;
                else hi = mid;
            }
            if (lo >= m || S.u[lo] != nbr[h]) {  // a neighbour that is not an uncovered chain voxel: not a ring
                S.bad = 1;
                lo = i;
            }
            loc[h] = lo;
        }
        S.l0[i] = (uint16_t)loc[0];
        S.l1[i] = (uint16_t)loc[1];
        S.par[i] = i;
        S.cnt[i] = 0;
    }
    __syncthreads();
    for (int i = tid; i < m; i += SKB_SMALL_CYCLE_THREADS) {  // union by minimum: the root is the ring's minimum
        const int32_t nbr[2] = {S.l0[i], S.l1[i]};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (nbr[h] >= i) continue;
            int32_t a = i, b = nbr[h];
            for (;;) {
                a = skb_smem_find(S.par, a);
                b = skb_smem_find(S.par, b);
                if (a == b) break;
                if (a < b) {
                    const int32_t t = a;
                    a = b;
                    b = t;
                }
                if (atomicCAS(S.par + a, a, b) == a) break;
            }
        }
    }
    __syncthreads();
    int32_t root_of[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = tid * PER + q;
        root_of[q] = i < m ? skb_smem_find(S.par, i) : -1;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = tid * PER + q;
        if (i < m) {
            S.par[i] = root_of[q];
            atomicAdd(S.cnt + root_of[q], 1);
        }
    }
    __syncthreads();
    // rank of every ring minimum among the minima (path order), and the entry offsets: one block scan of two sums
    uint32_t na = 0, nb = 0;  // rings, entries in this thread's voxels
    uint32_t ra[PER], rb[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = tid * PER + q;
        ra[q] = na;
        rb[q] = nb;
        if (i < m && root_of[q] == i) {
            na += 1u;
            nb += (uint32_t)S.cnt[i] + 1u;
        }
    }
    uint32_t ia = na, ib = nb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t oa = __shfl_up_sync(0xffffffffu, ia, d), ob = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) {
            ia += oa;
            ib += ob;
        }
    }
    if (lane == 31) {
        S.warp_a[warp] = ia;
        S.warp_b[warp] = ib;
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t wa = S.warp_a[lane], wb = S.warp_b[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t oa = __shfl_up_sync(0xffffffffu, wa, d), ob = __shfl_up_sync(0xffffffffu, wb, d);
            if (lane >= d) {
                wa += oa;
                wb += ob;
            }
        }
        S.warp_a[lane] = wa;
        S.warp_b[lane] = wb;
    }
    __syncthreads();
    const uint32_t ea = ia - na + (warp ? S.warp_a[warp - 1] : 0u), eb = ib - nb + (warp ? S.warp_b[warp - 1] : 0u);
    const uint32_t n_cycles = S.warp_a[31], n_entries = S.warp_b[31];
    const int32_t l_reg = A.mode == SKB_CYC_ALL ? A.pindptr[pid_base] : 0;  // entries of the regular paths
#pragma unroll
    for (int q = 0; q < PER; ++q) {
        const int i = tid * PER + q;
        if (i < m && root_of[q] == i) {
            const int32_t pid = pid_base + (int32_t)(ea + ra[q]);
            const int32_t gid = S.u[i] + A.id_shift;
            if (A.mode != SKB_CYC_ENTRIES) {
                A.start_node[pid] = gid;
                A.pmin[pid] = gid;
                A.plen[pid] = (uint32_t)S.cnt[i] + 1u;
                A.is_min[S.u[i]] = 1u;  // a junction-free cycle is a component of its own (csr.py:909)
            }
            if (A.mode == SKB_CYC_ALL) {
                A.pindptr[pid] = l_reg + (int32_t)(eb + rb[q]);
                S.base[i] = l_reg + (int32_t)(eb + rb[q]);
            } else if (A.mode == SKB_CYC_ENTRIES) {
                S.base[i] = A.pindptr[pid];
            }
            S.first[i] = (int32_t)((eb + rb[q]) - (ea + ra[q]));
            // the walk: from the minimum towards its smaller neighbour (csr.py:369-386), positions in shared memory only
            int32_t prev = i, cur = S.l0[i] < S.l1[i] ? S.l0[i] : S.l1[i];
            S.pos[i] = 0;
            S.order[S.first[i]] = (uint16_t)i;
            int32_t k = 1;
            while (cur != i && k <= m) {
                S.pos[cur] = (uint16_t)k;
                if (k < S.cnt[i]) S.order[S.first[i] + k] = (uint16_t)cur;
                const int32_t c0 = S.l0[cur], c1 = S.l1[cur];
                const bool from0 = c0 == prev;
                S.dir[cur] = from0 ? 0 : 1;
                prev = cur;
                cur = from0 ? c1 : c0;
                ++k;
            }
            S.dir[i] = S.l0[i] == prev ? 0 : 1;  // the closing entry arrives from the last voxel
            if (k != S.cnt[i]) S.bad = 1;
        }
    }
    if (tid == 0 && A.mode == SKB_CYC_ALL) A.pindptr[pid_base + (int32_t)n_cycles] = l_reg + (int32_t)n_entries;
    __syncthreads();
    for (int i = tid; i < m; i += SKB_SMALL_CYCLE_THREADS) {
        const int32_t r = S.par[i], node = S.u[i];
        const skb_d2 w = skb_ld_d2(rec + 2 * (int64_t)node + 1);
        const double w_in = S.dir[i] ? w.y : w.x;
        const double val = A.values ? A.values[node] : 1.0;
        S.wv[i] = w_in;
        S.vv[i] = val;
        if (A.mode != SKB_CYC_HEADS) {
            const int64_t at = (int64_t)S.base[r] + S.pos[i];
            A.pidx[at] = node + A.id_shift;
            if (A.pdata) A.pdata[at] = val;
            A.pw[at] = r == i ? 0.0 : w_in;
            if (r == i) {  // the ring closes on its minimum
                const int64_t end = (int64_t)S.base[r] + S.cnt[r];
                A.pidx[end] = node + A.id_shift;
                if (A.pdata) A.pdata[end] = val;
                A.pw[end] = w_in;
            }
        }
    }
    if (A.head_aux && A.mode != SKB_CYC_ENTRIES) {  // Z-slabs: the sums the statistics are made of, in path order
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int i = tid * PER + q;
            if (i < m && root_of[q] == i) {
                SkbAux a;
                a.sw = 0.0;
                a.sx = S.vv[i];
                a.sxx = S.vv[i] * S.vv[i];
                for (int32_t k = 1; k < S.cnt[i]; ++k) {
                    const int j = S.order[S.first[i] + k];
                    a.sw += S.wv[j];
                    a.sx += S.vv[j];
                    a.sxx += S.vv[j] * S.vv[j];
                }
                a.sw += S.wv[i];
                a.sx += S.vv[i];
                a.sxx += S.vv[i] * S.vv[i];
                a.end_lin = (A.lin ? A.lin[S.u[i]] : (int64_t)S.u[i]) + A.lin_offset;
                a.end_node = S.u[i] + A.id_shift;
                a.end_deg = 2;
                a.end_key = 0;
                a.pad = 0;
                a.end_value = S.vv[i];
                a.pad2 = 0.0;
                A.head_aux[pid_base + (int32_t)(ea + ra[q])] = a;
            }
        }
    }
    if (tid == 0) {
        A.out[0] = (int32_t)n_cycles;
        A.out[1] = (int32_t)n_entries;
        A.out[2] = S.bad;
    }
}

static int skb_small_cycles_launch(const SkbSmallCycleArgs& a_in, skb_stream_t s) {
    static uint8_t attr_set[SKB_MAX_DEV] = {};
    SKB_TRY(skb_optin_smem(skb_small_cycles_kernel, sizeof(SkbSmallCycleSmem), attr_set));
    SkbSmallCycleArgs a = a_in;
    a.guard = g_skb_guard;
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(skb_launch_pdl(skb_small_cycles_kernel, dim3(1), dim3(SKB_SMALL_CYCLE_THREADS),
                                         sizeof(SkbSmallCycleSmem), s, a),
                          "small cycles launch");
}

// One GPU: list i32[m], out i32[4] (device).  Returns 0, or 1 when the case is not a small one (the caller runs the
// general path).
static int skb_cycles_small(const skb_i4* rec, const uint8_t* covered, int64_t n, int64_t m, const double* values,
                            int64_t n_regular, int32_t* start_node, uint32_t* plen, int32_t* pmin, int32_t* pindptr,
                            int32_t* pidx, double* pdata, double* pw, uint32_t* is_min, int32_t* list, int32_t* out,
                            skb_stream_t s) {
    if (m <= 0 || m > SKB_SMALL_CYCLE_NODES || !g_skb_small_cycles) return 1;
    SKB_TRY(skb_fill_bytes(out, 0, 16, s));
    SKB_TRY(skb_foreach_unmarked(n, covered, s, SkbCollectUncoveredItem{rec, covered, list, out + 3, (int32_t)m, 0, 0x7fffffff}));
    SkbSmallCycleArgs a{list, out + 3, (int32_t)m, SKB_CYC_ALL, rec, values, (int32_t)n_regular, 0, nullptr, 0, start_node, plen,
                        pmin, pindptr, pidx, pdata, pw, is_min, nullptr, out};
    return skb_small_cycles_launch(a, s);
}

// Z-slabs, rings inside this rank's slab.  Heads: the owned uncovered chain voxels are collected (list i32[m]) and the rings'
// heads written to start_node / plen / pmin / head_aux[0 ...]; out i32[4] = rings, entries, error, voxels collected.
static int skb_cycles_small_heads(const skb_i4* rec, const uint8_t* covered, int64_t n, int64_t m, int64_t own0, int64_t own1,
                                  const double* values, const int64_t* lin, int64_t lin_offset, int64_t id_shift,
                                  int32_t* start_node, uint32_t* plen, int32_t* pmin, uint32_t* is_min, void* head_aux,
                                  int32_t* list, int32_t* out, skb_stream_t s) {
    if (m <= 0 || m > SKB_SMALL_CYCLE_NODES || !g_skb_small_cycles) return 1;
    SKB_TRY(skb_fill_bytes(out, 0, 16, s));
    SKB_TRY(skb_foreach_unmarked(n, covered, s,
                                 SkbCollectUncoveredItem{rec, covered, list, out + 3, (int32_t)m, (int32_t)own0, (int32_t)own1}));
    SkbSmallCycleArgs a{list, out + 3, (int32_t)m, SKB_CYC_HEADS, rec, values, 0, (int32_t)id_shift, lin, lin_offset, start_node,
                        plen, pmin, nullptr, nullptr, nullptr, nullptr, is_min, (SkbAux*)head_aux, out};
    return skb_small_cycles_launch(a, s);
}
// ... and, once the entry offsets are known (gptr[pid_base + j] for the j-th ring), the entries
static int skb_cycles_small_entries(const skb_i4* rec, int64_t m, const double* values, int64_t id_shift, int64_t pid_base,
                                    int32_t* gptr, int32_t* pidx, double* pdata, double* pw, const int32_t* list,
                                    int32_t* out, skb_stream_t s) {
    SkbSmallCycleArgs a{list, out + 3, (int32_t)m, SKB_CYC_ENTRIES, rec, values, (int32_t)pid_base, (int32_t)id_shift, nullptr, 0,
                        nullptr, nullptr, nullptr, gptr, pidx, pdata, pw, nullptr, nullptr, out};
    return skb_small_cycles_launch(a, s);
}

// ------------------------------------------------------------------------------- statistics of very long paths
// One thread block per path of more than SKB_LONG_PATH entries (the thread-per-path item skips those): every thread sums
// the entries j = t, t + 256, ... in order, the 256 partial sums are combined in a fixed tree, so the result does not
// depend on timing.  The order differs from the reference's left-to-right (distance) and pairwise (mean, stdev) sums:
// agreement is to the rounding error of those sums themselves (~n eps), see DESIGN.md.
__global__ void __launch_bounds__(256) skb_long_path_stats_kernel(const int32_t* __restrict__ pindptr,
                                                                  const double* __restrict__ pdata,
                                                                  const double* __restrict__ pw, int64_t n_paths,
                                                                  double* dist, double* mean, double* stdev,
                                                                  const int32_t* n_ptr, int32_t long_path) {
    SKB_PDL_ENTRY();
    if (n_ptr && *n_ptr < n_paths) n_paths = *n_ptr;  // the number of paths in device memory
    __shared__ double red[3][256];
    __shared__ int s_list[256];
    __shared__ int s_n;
    const int tid = threadIdx.x;
    // a block looks at 256 paths at a time (coalesced loads of their bounds): long paths are rare, so the usual pass over
    // a group ends at the vote
    for (int64_t base = (int64_t)blockIdx.x * 256; base < n_paths; base += (int64_t)gridDim.x * 256) {
        const int64_t q = base + tid;
        const bool is_long = q < n_paths && pindptr[q + 1] - pindptr[q] > long_path;
        if (!__syncthreads_or(is_long)) continue;
        if (tid == 0) s_n = 0;
        __syncthreads();
        if (is_long) s_list[atomicAdd(&s_n, 1)] = tid;  // (the order of the list does not enter any result)
        __syncthreads();
        const int n_long = s_n;
        for (int i = 0; i < n_long; ++i) {
            const int64_t p = base + s_list[i];
            const int32_t s = pindptr[p], e = pindptr[p + 1];
            double a = 0.0, b = 0.0, c = 0.0;
            for (int32_t j = s + tid; j < e; j += 256) {
                if (j > s) a += pw[j];
                if (pdata) {
                    const double x = pdata[j];
                    b += x;
                    c += x * x;
                }
            }
            red[0][tid] = a;
            red[1][tid] = b;
            red[2][tid] = c;
            __syncthreads();
            for (int d = 128; d > 0; d >>= 1) {
                if (tid < d) {
                    red[0][tid] += red[0][tid + d];
                    red[1][tid] += red[1][tid + d];
                    red[2][tid] += red[2][tid + d];
                }
                __syncthreads();
            }
            if (tid == 0) {
                const double n = (double)(e - s);
                if (dist) dist[p] = red[0][0];
                if (!pdata) {
                    if (mean) mean[p] = 1.0;
                    if (stdev) stdev[p] = 0.0;
                } else {
                    const double m = red[1][0] / n;
                    if (mean) mean[p] = m;
                    if (stdev) {
                        double var = red[2][0] / n - m * m;
                        if (!(var > 0.0)) var = (var == var) ? 0.0 : var;
                        stdev[p] = sqrt(var);
                    }
                }
            }
            __syncthreads();
        }
    }
}

static int skb_long_path_stats(const int32_t* pindptr, const double* pdata, const double* pw, int64_t n_paths, double* dist,
                               double* mean, double* stdev, skb_stream_t s, const int32_t* n_ptr = nullptr) {
    if (n_paths <= 0) return 0;
    int64_t blocks = (n_paths + 255) / 256;
    if (blocks > 4 * (int64_t)skb_sm_count()) blocks = 4 * (int64_t)skb_sm_count();
    if (g_skb_long_path == 0x7fffffff) return 0;  // "long_path_exact": every path was summed in the reference's order
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(skb_launch_pdl(skb_long_path_stats_kernel, dim3((unsigned)blocks), dim3(256), 0, s, pindptr, pdata, pw,
                                         n_paths, dist, mean, stdev, n_ptr, g_skb_long_path),
                          "long path stats launch");
}

#include "skb_api.inl"

static int skb_junction_append(const uint64_t* bits, const uint32_t* pre256, int ndim, const int64_t* shape, const int64_t* lin,
                               const uint32_t* nb, int64_t n, int64_t a0, int64_t a1, int32_t* ea, int32_t* eb, uint8_t* ek,
                               int64_t cap, int32_t* counter, skb_stream_t s, const int32_t* n_ptr, const SkbDense* dense) {
    if (!g_skb_junction_append || !g_skb_mst_keys || n <= 0 || cap <= 0) return 1;  // the keyed Boruvka needs no edge order
    SKB_TRY(skb_check_geom(ndim, shape));
    SKB_TRY(skb_fill_bytes(counter, 0, 4, s));
    if (dense && dense->jcounter) SKB_TRY(skb_fill_bytes(dense->jcounter, 0, 4, s));
    int64_t blocks = (n + 255) / 256;
    const int64_t max_blocks = (int64_t)skb_sm_count() * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    skb_junction_append_kernel<<<(unsigned)blocks, 256, 0, s>>>(skb_make_geom(ndim, shape), bits, pre256, lin, nb, n, a0, a1, ea, eb,
                                                                ek, cap > 0xffffffffLL ? 0xffffffffu : (uint32_t)cap, counter, n_ptr,
                                                                dense ? dense->jid : nullptr, dense ? dense->jcounter : nullptr);
    SKB_COUNT_LAUNCH(1);
    return skb_cuda_check(cudaGetLastError(), "junction append launch");
}

extern "C" {

const char* skb_last_error(void) { return g_skb_err; }

int64_t skb_scan_scratch_bytes(int64_t n) {
    int64_t nb = (n + SKB_SCAN_TILE - 1) / SKB_SCAN_TILE;
    return (nb + 2) * 8;
}

int64_t skb_tile_voxels(void) { return SKB_TILE_VOX; }

int skb_set_option(const char* name, int64_t value) {
    if (!name) return skb_fail("skb_set_option: null name");
    if (!strcmp(name, "splitter_bits")) {
        if (value < 0 || value > 12) return skb_fail("skb_set_option: splitter_bits must be in 0..12 (0 = chosen by the node count)");
        g_skb_split_bits = (int)value;
        return 0;
    }
    if (!strcmp(name, "emit_both_ways")) {
        g_skb_both_ways = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "mst_blocks_per_sm")) {
        if (value < 0 || value > 32) return skb_fail("skb_set_option: mst_blocks_per_sm must be in 0..32");
        g_skb_mst_blocks_per_sm = (int)value;
        return 0;
    }
    if (!strcmp(name, "emit_blocks_per_sm")) {
        if (value < 1 || value > 64) return skb_fail("skb_set_option: emit_blocks_per_sm must be in 1..64");
        g_skb_emit_blocks_per_sm = (int)value;
        return 0;
    }
    if (!strcmp(name, "junction_append")) {
        g_skb_junction_append = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "arena_guard")) {
        g_skb_arena_guard = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "small_cycles")) {
        g_skb_small_cycles = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "mst_keys")) {
        g_skb_mst_keys = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "dense_mst_min_nodes")) {
        if (value < 0) return skb_fail("skb_set_option: dense_mst_min_nodes must be >= 0");
        g_skb_dense_mst_min_nodes = value;
        return 0;
    }
    if (!strcmp(name, "pdl")) {
        g_skb_pdl = value ? 1 : 0;
        return 0;
    }
    if (!strcmp(name, "long_path_exact")) {
        g_skb_long_path = value ? 0x7fffffff : SKB_LONG_PATH;
        return 0;
    }
    return skb_fail("skb_set_option: unknown option");
}

int64_t skb_kernel_launches(void) { return (int64_t)__atomic_load_n(&g_skb_launches, __ATOMIC_RELAXED); }

// mask sweep (replaces csr.py:1070,122-123,131).  bits: ntiles*256 u64 words (the caller owns
// two zeroed pad words on either side); tile_counts: u32[ntiles].  variant 0 = vector loads,
// 1 = TMA bulk copies through shared memory.  The image must be 16-byte aligned.
int skb_pack_count(const void* image, int dtype, int64_t nvox, uint64_t* bits, uint32_t* tile_counts,
                   int64_t ntiles, int variant, void* stream) {
    if (dtype < 0 || dtype >= SKB_DT_COUNT) return skb_fail("unsupported image dtype");
    if (((uintptr_t)image & 15) != 0) return skb_fail("image pointer must be 16-byte aligned");
    if (ntiles != (nvox + SKB_TILE_VOX - 1) / SKB_TILE_VOX) return skb_fail("ntiles does not match nvox");
    if (nvox <= 0) return 0;
    bool is_float = dtype == SKB_DT_F32 || dtype == SKB_DT_F64;
    skb_stream_t s = (skb_stream_t)stream;
    switch (skb_dtype_size(dtype)) {
        case 1: return skb_pack_launch<1>(image, nvox, is_float, bits, tile_counts, ntiles, variant, s);
        case 2: return skb_pack_launch<2>(image, nvox, is_float, bits, tile_counts, ntiles, variant, s);
        case 4: return skb_pack_launch<4>(image, nvox, is_float, bits, tile_counts, ntiles, variant, s);
        default: return skb_pack_launch<8>(image, nvox, is_float, bits, tile_counts, ntiles, variant, s);
    }
}

// nodes in raster order (csr.py:131-132,1088,554-562).  tile_prefix: exclusive scan of the
// tile counts (ntiles+1).  values may be null (bool images).  pre256: u32[ntiles*64].  z_origin is
// added to the first coordinate of 3-D images (the volume may be one Z-slab of a larger one).
int skb_emit_nodes_cap(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                       const int64_t* shape, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                       int64_t* coords, double* values, uint32_t* pre256, int64_t node_cap, void* stream);
int skb_emit_nodes(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                   const int64_t* shape, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                   int64_t* coords, double* values, uint32_t* pre256, void* stream) {
    return skb_emit_nodes_cap(bits, tile_prefix, ntiles, ndim, shape, image, dtype, z_origin, lin, coords, values,
                              pre256, 0xffffffffLL, stream);
}

// the same with a capacity for lin / coords / values: nodes with id >= node_cap are not written
int skb_emit_nodes_cap(const uint64_t* bits, const uint32_t* tile_prefix, int64_t ntiles, int ndim,
                       const int64_t* shape, const void* image, int dtype, int64_t z_origin, int64_t* lin,
                       int64_t* coords, double* values, uint32_t* pre256, int64_t node_cap, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    if (ntiles <= 0) return 0;
    if (values && !image) return skb_fail("values requested without an image");
    if (((uintptr_t)bits & 15) != 0) return skb_fail("mask pointer must be 16-byte aligned");
    const uint32_t cap32 = node_cap < 0 ? 0u : node_cap > 0xffffffffLL ? 0xffffffffu : (uint32_t)node_cap;
    int64_t grid = (ntiles + 2 * SKB_EMIT_WARPS - 1) / (2 * SKB_EMIT_WARPS);  // a warp takes two tiles per iteration
    int64_t cap = (int64_t)skb_sm_count() * g_skb_emit_blocks_per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    SKB_COUNT_LAUNCH(1);
    const SkbGeom g = skb_make_geom(ndim, shape);
    double* no_values = nullptr;
    return skb_cuda_check(
        values ? skb_launch_pdl(skb_emit_nodes_kernel<true>, dim3((unsigned)grid), dim3(SKB_EMIT_THREADS), 0, (skb_stream_t)stream,
                                bits, tile_prefix, ntiles, g, image, dtype, z_origin, lin, coords, values, pre256, cap32)
               : skb_launch_pdl(skb_emit_nodes_kernel<false>, dim3((unsigned)grid), dim3(SKB_EMIT_THREADS), 0,
                                (skb_stream_t)stream, bits, tile_prefix, ntiles, g, image, dtype, z_origin, lin, coords,
                                no_values, pre256, cap32),
        "emit_nodes launch");
}

// Mask sweep, scan and node emission in one pass (csr.py:1070,122-123,131-132,1088,554-562): what
// skb_pack_count + skb_scan_u32 + skb_emit_nodes produce, except the tile counts.  lin / coords /
// values have room for node_cap nodes; *n_nodes (device) receives the true count, and when it
// exceeds node_cap the arrays hold the first node_cap nodes only (bits and pre256 are complete:
// call skb_tile_prefix + skb_emit_nodes with larger arrays).  scratch: skb_sweep_scratch_bytes(nvox).
int64_t skb_sweep_scratch_bytes(int64_t nvox) { return ((nvox + SKB_CHUNK_VOX - 1) / SKB_CHUNK_VOX + 2 + 128) * 8; }

int skb_sweep_nodes(const void* image, int dtype, int ndim, const int64_t* shape, int64_t z_origin, uint64_t* bits,
                    int64_t ntiles, uint32_t* pre256, int64_t node_cap, int64_t* lin, int64_t* coords, double* values,
                    void* scratch, int32_t* n_nodes, void* stream) {
    SKB_TRY(skb_check_geom(ndim, shape));
    SkbGeom g = skb_make_geom(ndim, shape);
    const int64_t nvox = g.nvox;
    if (dtype < 0 || dtype >= SKB_DT_COUNT) return skb_fail("unsupported image dtype");
    if (((uintptr_t)image & 15) != 0) return skb_fail("image pointer must be 16-byte aligned");
    if (((uintptr_t)bits & 7) != 0) return skb_fail("mask pointer must be 8-byte aligned");
    if (ntiles != (nvox + SKB_TILE_VOX - 1) / SKB_TILE_VOX) return skb_fail("ntiles does not match the shape");
    skb_stream_t s = (skb_stream_t)stream;
    if (nvox <= 0) return skb_cuda_check(cudaMemsetAsync(n_nodes, 0, 8, s), "memset");
    bool is_float = dtype == SKB_DT_F32 || dtype == SKB_DT_F64;
    switch (skb_dtype_size(dtype)) {
        case 1: return skb_sweep_nodes_launch<1>(image, nvox, is_float, bits, ntiles, g, dtype, z_origin, node_cap, lin, coords, values, pre256, scratch, n_nodes, s);
        case 2: return skb_sweep_nodes_launch<2>(image, nvox, is_float, bits, ntiles, g, dtype, z_origin, node_cap, lin, coords, values, pre256, scratch, n_nodes, s);
        case 4: return skb_sweep_nodes_launch<4>(image, nvox, is_float, bits, ntiles, g, dtype, z_origin, node_cap, lin, coords, values, pre256, scratch, n_nodes, s);
        default: return skb_sweep_nodes_launch<8>(image, nvox, is_float, bits, ntiles, g, dtype, z_origin, node_cap, lin, coords, values, pre256, scratch, n_nodes, s);
    }
}

}  // extern "C"

#include "skb_run.inl"
#include "skb_slab.inl"
