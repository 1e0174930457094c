// Shared device/host helpers for libslb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/slb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libslb200 is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (thread-local message, no exceptions across the ABI)
// ---------------------------------------------------------------------------------------------
void slb_set_error(const char* fmt, ...);

#define SLB_REQUIRE(cond, code, ...)  \
    do {                              \
        if (!(cond)) {                \
            slb_set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

#define SLB_CUDA_OK(expr)                                                                    \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            slb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SLB_ECUDA;                                                                \
        }                                                                                    \
    } while (0)

// launch check: catches bad configurations without synchronising
#define SLB_LAUNCH_OK(name)                                                                 \
    do {                                                                                     \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            slb_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));          \
            return SLB_ECUDA;                                                                \
        }                                                                                    \
        slb_count_launch();                                                                  \
    } while (0)

void slb_count_launch();  // process-wide counter of kernels this library has launched (slb_launch_count)

// Optional live profiling (slb_profile_begin / slb_profile_end): while enabled, every single-kernel entry point brackets
// its launch with two CUDA events on the launch stream and notes its algorithmic work; nothing synchronises until
// slb_profile_summary. Costs one branch when disabled.
bool slb_profile_enabled();
void slb_profile_open(const char* name, void* stream, double flops, double bytes, void** token);
void slb_profile_close(void* token, void* stream);
struct SlbProfScope {
    void* token = nullptr;
    void* stream;
    SlbProfScope(const char* name, void* st, double flops, double bytes) : stream(st) {
        if (slb_profile_enabled()) slb_profile_open(name, st, flops, bytes, &token);
    }
    ~SlbProfScope() {
        if (token) slb_profile_close(token, stream);
    }
};
int slb_sm_count();  // cached per process (current device at first call)

static inline int64_t slb_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// bf16 helpers — bit-exact with c10::BFloat16 round_to_nearest_even (NaN -> 0x7FC0)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint16_t slb_f32_to_bf16_bits(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } cvt; cvt.f = f; u = cvt.u;
#endif
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0x7FC0;
    uint32_t bias = ((u >> 16) & 1u) + 0x7FFFu;
    return (uint16_t)((u + bias) >> 16);
}

__host__ __device__ __forceinline__ float slb_bf16_bits_to_f32(uint16_t b) {
    uint32_t u = ((uint32_t)b) << 16;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } cvt; cvt.u = u; return cvt.f;
#endif
}

// ---------------------------------------------------------------------------------------------
// top-k sort key: one u64 whose descending order is
//   (bf16 value desc, +0 == -0, NaN greatest)  then  (real ids ascending)  then  (placeholders id=-1 last)
// layout: [63:48] order-preserving value key | [47:1] (2^47-1 - id), 0 for id=-1 | [0] sign bit of a zero value
// ---------------------------------------------------------------------------------------------
#define SLB_ID_MASK ((1ull << 47) - 1ull)

__host__ __device__ __forceinline__ uint64_t slb_topk_key(uint16_t bits, int64_t id) {
    uint32_t vkey;
    uint64_t zsign = 0;
    if ((bits & 0x7FFFu) > 0x7F80u) {
        vkey = 0xFFFFu;  // NaN sorts first (ATen TopKImpl.h:57)
    } else if ((bits & 0x7FFFu) == 0) {
        vkey = 0x8000u;
        zsign = bits >> 15;
    } else if (bits & 0x8000u) {
        vkey = (~(uint32_t)bits) & 0xFFFFu;
    } else {
        vkey = (uint32_t)bits | 0x8000u;
    }
    uint64_t idf = (id < 0) ? 0ull : ((SLB_ID_MASK - (uint64_t)id) & SLB_ID_MASK);
    return ((uint64_t)vkey << 48) | (idf << 1) | zsign;
}

__host__ __device__ __forceinline__ void slb_topk_unkey(uint64_t key, uint16_t* bits, int64_t* id) {
    uint32_t vkey = (uint32_t)(key >> 48);
    uint16_t b;
    if (vkey == 0xFFFFu) {
        b = 0x7FC0;
    } else if (vkey == 0x8000u) {
        b = (uint16_t)((key & 1ull) << 15);
    } else if (vkey & 0x8000u) {
        b = (uint16_t)(vkey & 0x7FFFu);
    } else {
        b = (uint16_t)((~vkey) & 0xFFFFu);
    }
    uint64_t idf = (key >> 1) & SLB_ID_MASK;
    *bits = b;
    *id = (idf == 0) ? -1 : (int64_t)(SLB_ID_MASK - idf);
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers: shared-memory addresses, mbarrier, bulk async copy (TMA 1-D), fences
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t slb_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void slb_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(slb_smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void slb_fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// make generic-proxy smem writes visible to the async proxy (TMA / tcgen05 reads)
__device__ __forceinline__ void slb_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void slb_mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(slb_smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void slb_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(slb_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool slb_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(slb_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void slb_mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!slb_mbar_try_wait(bar, parity)) {
    }
}

// 1-D bulk async copy global -> shared (UBLKCP), completion counted in bytes on `bar`.
// dst/src 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void slb_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     slb_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(slb_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ float slb_warp_sum_butterfly(float v) {
    v = v + __shfl_xor_sync(0xffffffffu, v, 16);
    v = v + __shfl_xor_sync(0xffffffffu, v, 8);
    v = v + __shfl_xor_sync(0xffffffffu, v, 4);
    v = v + __shfl_xor_sync(0xffffffffu, v, 2);
    v = v + __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
#endif  // __CUDACC__
