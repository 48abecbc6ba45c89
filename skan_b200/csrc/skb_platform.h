// skan_b200 -- platform layer shared by the sm_100a build (skb_cuda.cu) and the host
// emulation harness used by the CPU-only unit tests (tests/emul/skb_emul.cpp).
//
// The per-item device functions of the N-proportional stages (skb_items.cuh) and the C-ABI
// bodies that sequence them (skb_api.inl) are written once against this small layer:
//   SKB_HD                       __host__ __device__ inline
//   skb_popc / skb_popcll / skb_ffs
//   skb_atomic_*                 atomics (device) / plain read-modify-write (single host thread)
//   skb_ld_cg / skb_st_cg        L2-coherent loads/stores for data other CTAs update in-kernel
//   skb_foreach(n, stream, f)    one logical thread per item
// The emulation harness is test infrastructure for debugging the item logic without a GPU;
// the python package never loads it (skan_b200/_native.py only loads libskan_b200.so).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define SKB_HD __host__ __device__ __forceinline__
#define SKB_DEVICE_BUILD 1
#else
#define SKB_HD inline
#define SKB_DEVICE_BUILD 0
#endif

struct skb_i2 {
    int32_t x, y;
};

#if defined(__CUDA_ARCH__)
// ------------------------------------------------------------------ device side
SKB_HD int skb_popc(uint32_t v) { return __popc(v); }
SKB_HD int skb_popcll(uint64_t v) { return __popcll(v); }
SKB_HD int skb_ffs(uint32_t v) { return __ffs((int)v); }
SKB_HD int skb_ffsll(uint64_t v) { return __ffsll((long long)v); }
SKB_HD uint32_t skb_umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
SKB_HD uint64_t skb_atomic_min_u64(uint64_t* p, uint64_t v) {
    return (uint64_t)atomicMin((unsigned long long*)p, (unsigned long long)v);
}
SKB_HD uint32_t skb_atomic_min_u32(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
SKB_HD uint32_t skb_atomic_and_u32(uint32_t* p, uint32_t v) { return atomicAnd(p, v); }
SKB_HD int32_t skb_atomic_cas_i32(int32_t* p, int32_t cmp, int32_t v) { return atomicCAS(p, cmp, v); }
SKB_HD void skb_atomic_add_u64(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
// maximum of a value that few items raise: look before the atomic
SKB_HD void skb_atomic_max_u64_lazy(uint64_t* p, uint64_t v) {
    if (__ldcg((const unsigned long long*)p) < v) atomicMax((unsigned long long*)p, (unsigned long long)v);
}
SKB_HD int32_t skb_ld_cg(const int32_t* p) { return __ldcg(p); }
SKB_HD uint32_t skb_ld_cg(const uint32_t* p) { return __ldcg(p); }
SKB_HD uint64_t skb_ld_cg(const uint64_t* p) { return (uint64_t)__ldcg((const unsigned long long*)p); }
SKB_HD skb_i2 skb_ld_cg(const skb_i2* p) {
    long long v = __ldcg((const long long*)p);
    skb_i2 r;
    r.x = (int32_t)(v & 0xffffffffLL);
    r.y = (int32_t)(v >> 32);
    return r;
}
SKB_HD void skb_st_cg(skb_i2* p, skb_i2 v) {
    long long w = ((long long)(uint32_t)v.x) | (((long long)v.y) << 32);
    __stcg((long long*)p, w);
}
SKB_HD void skb_st_cg(int32_t* p, int32_t v) { __stcg(p, v); }
SKB_HD void skb_raise(int32_t* flag) { *flag = 1; }
SKB_HD double skb_sqrt(double v) { return sqrt(v); }
#else
// ------------------------------------------------------------------ host side
#include <math.h>
SKB_HD int skb_popc(uint32_t v) { return __builtin_popcount(v); }
SKB_HD int skb_popcll(uint64_t v) { return __builtin_popcountll(v); }
SKB_HD int skb_ffs(uint32_t v) { return __builtin_ffs((int)v); }
SKB_HD int skb_ffsll(uint64_t v) { return __builtin_ffsll((long long)v); }
SKB_HD uint32_t skb_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
SKB_HD uint64_t skb_atomic_min_u64(uint64_t* p, uint64_t v) {
    uint64_t o = *p;
    if (v < o) *p = v;
    return o;
}
SKB_HD uint32_t skb_atomic_min_u32(uint32_t* p, uint32_t v) {
    uint32_t o = *p;
    if (v < o) *p = v;
    return o;
}
SKB_HD uint32_t skb_atomic_and_u32(uint32_t* p, uint32_t v) {
    uint32_t o = *p;
    *p = o & v;
    return o;
}
SKB_HD int32_t skb_atomic_cas_i32(int32_t* p, int32_t cmp, int32_t v) {
    int32_t o = *p;
    if (o == cmp) *p = v;
    return o;
}
SKB_HD void skb_atomic_add_u64(unsigned long long* p, unsigned long long v) { *p += v; }
SKB_HD void skb_atomic_max_u64_lazy(uint64_t* p, uint64_t v) {
    if (*p < v) *p = v;
}
SKB_HD int32_t skb_ld_cg(const int32_t* p) { return *p; }
SKB_HD uint32_t skb_ld_cg(const uint32_t* p) { return *p; }
SKB_HD uint64_t skb_ld_cg(const uint64_t* p) { return *p; }
SKB_HD skb_i2 skb_ld_cg(const skb_i2* p) { return *p; }
SKB_HD void skb_st_cg(skb_i2* p, skb_i2 v) { *p = v; }
SKB_HD void skb_st_cg(int32_t* p, int32_t v) { *p = v; }
SKB_HD void skb_raise(int32_t* flag) { *flag = 1; }
SKB_HD double skb_sqrt(double v) { return sqrt(v); }
#endif

SKB_HD uint64_t skb_f64_bits(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    return b;
}
