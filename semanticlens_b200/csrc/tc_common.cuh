// tcgen05 / TMEM / TMA (tensor-map) PTX wrappers for sm_100a. Hand-written: no CUTLASS/CuTe in product code.
#pragma once
#include "slb_common.cuh"

#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

// ---------------------------------------------------------------------------------------------
// host: tensor-map encoder without linking libcuda
// ---------------------------------------------------------------------------------------------
typedef CUresult (*slb_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                       const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
slb_tmap_encode_fn slb_get_tmap_encoder();  // nullptr (+ error string) if the driver does not provide it

// 3-D map over 16-bit planes [planes][rows][cols] (cols contiguous); box = {64 cols, box_rows, planes}; 128B swizzle.
// (cached per host thread by its arguments; slb_encode_plane_map always calls the driver)
int slb_make_plane_map(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int planes, int box_rows);
int slb_encode_plane_map(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int planes, int box_rows);
// output map of the GEMM's TMA-store epilogue: planes == 0 -> fp32 (M, N); planes == 2 -> 16-bit planes (2, M, N)
int slb_make_store_map(CUtensorMap* out, const void* base, int64_t M, int64_t N, int planes);

// im2col-mode map over ONE channels-last 16-bit plane (B, H, W, C) for a ksize x ksize / stride / pad convolution: a load
// brings `pixels` consecutive output pixels x 64 channels of one filter tap (128B swizzle, zero fill outside the image).
int slb_make_im2col_map(CUtensorMap* out, const void* base, int64_t B, int64_t H, int64_t W, int64_t C, int ksize, int stride,
                        int pad, int pixels);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// device: TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void slb_prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<uint64_t>(p)));
}
__device__ __forceinline__ void slb_prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void slb_tma_load_3d(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2,
                                                uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(slb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2),
        "r"(slb_smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void slb_tma_load_4d(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(slb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(slb_smem_u32(bar))
        : "memory");
}

// im2col-mode load: {c, w, h, n} = channel offset and the coordinates of the FIRST output pixel's window corner in the
// input (w = column * stride - pad, ...); {off_w, off_h} = the filter tap. The TMA unit steps through the output pixels.
__device__ __forceinline__ void slb_tma_load_im2col_4d(void* smem_dst, const CUtensorMap* m, int c, int w, int h, int n,
                                                       uint16_t off_w, uint16_t off_h, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
        ::"r"(slb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c), "r"(w), "r"(h), "r"(n), "r"(slb_smem_u32(bar)),
        "h"(off_w), "h"(off_h)
        : "memory");
}

// ---------------------------------------------------------------------------------------------
// device: tcgen05 (cta_group::1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void slb_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void slb_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void slb_tmem_alloc(uint32_t* smem_slot) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slb_smem_u32(smem_slot)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void slb_tmem_dealloc(uint32_t taddr) {  // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle, rows of 64 16-bit elements (128 B), 8-row atoms of 1024 B.
// Bit layout (PTX "shared memory matrix descriptor", sm_100): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version = 1, [49,52) base offset = 0, [61,64) swizzle mode (2 = 128B).
__device__ __forceinline__ uint64_t slb_umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;              // LBO: unused for K-major swizzled tiles whose K extent is one atom
    d |= (uint64_t)(1024 >> 4) << 32;    // SBO: 8 rows * 128 B between row groups
    d |= (uint64_t)1 << 46;              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;              // SWIZZLE_128B
    return d;
}

// The same for rows of 32 16-bit elements (64 B) under the 64-byte swizzle: 8-row atoms of 512 B. Used by the implicit
// convolution over 32-channel maps, whose k-block is one filter tap = 32 channels (CLIP ModifiedResNet stem).
__device__ __forceinline__ uint64_t slb_umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;     // SBO: 8 rows * 64 B between row groups
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;              // SWIZZLE_64B
    return d;
}

// kind::f16 instruction descriptor: D fp32, A/B both `fmt` (0 = fp16, 1 = bf16), both K-major, shape M x N x 16.
__device__ __forceinline__ uint32_t slb_umma_idesc_f16(int fmt, int M, int N) {
    uint32_t d = 0;
    d |= 1u << 4;                        // c_format = F32
    d |= (uint32_t)fmt << 7;             // a_format
    d |= (uint32_t)fmt << 10;            // b_format
    d |= (uint32_t)(N >> 3) << 17;       // n_dim
    d |= (uint32_t)(M >> 4) << 24;       // m_dim
    return d;
}

__device__ __forceinline__ void slb_umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// arrive on `bar` once every tcgen05 op issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void slb_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(slb_smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// device: thread-block clusters and CTA pairs (cta_group::2)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t slb_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void slb_cluster_sync() {  // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the location `smem_addr` (a shared::cta address of this CTA) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t slb_mapa(uint32_t smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void slb_mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// TMA tile load issued by either CTA of a pair into ITS OWN shared memory; the bytes are counted on the mbarrier at
// shared::cluster address `bar_cluster` (the leader CTA's "full" barrier).
__device__ __forceinline__ void slb_tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, int c0, int c1, int c2,
                                                     uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(slb_smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster)
        : "memory");
}

template <int COLS>
__device__ __forceinline__ void slb_tmem_alloc_pair(uint32_t* smem_slot) {  // the same warp of BOTH CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slb_smem_u32(smem_slot)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void slb_tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

// one MMA across the CTA pair: D (256 x N, rows split over the two CTAs' TMEM) += A (256 x 16) * B (N x 16)^T, issued by
// one thread of the leader CTA; descriptors are shared::cta offsets valid in both CTAs
__device__ __forceinline__ void slb_umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in every CTA of `cta_mask` once all prior MMAs completed
__device__ __forceinline__ void slb_umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     slb_smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp receives row (lane base + i), columns [col, col+32)
__device__ __forceinline__ void slb_tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void slb_tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void slb_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// 16-bit plane formats: 0 = IEEE fp16, 1 = bf16  (same codes as the instruction descriptor)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t slb_to_plane(float v, int fmt) {
    if (fmt == 0) return __half_as_ushort(__float2half_rn(v));
    return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float slb_from_plane(uint16_t b, int fmt) {
    if (fmt == 0) return __half2float(__ushort_as_half(b));
    return __bfloat162float(__ushort_as_bfloat16(b));
}
// s * x ~= hi + lo with hi = rn16(s * x), lo = rn16(s * x - hi); the caller applies the tensor's power-of-two scale s
// (see slb200.h). 22 (fp16) / 16 (bf16) significant bits in total. fp16 planes saturate at +-65504.
__device__ __forceinline__ void slb_split2(float v, int fmt, uint16_t& hi, uint16_t& lo) {
    if (fmt == 0 && v == v) v = fminf(fmaxf(v, -65504.0f), 65504.0f);  // saturate, but let NaN through (fminf / fmaxf drop it)
    hi = slb_to_plane(v, fmt);
    lo = slb_to_plane(v - slb_from_plane(hi, fmt), fmt);
}
// planes written by a kernel (activations) carry SLB_ACT_PLANE_SCALE
__device__ __forceinline__ void slb_split2_act(float v, int fmt, uint16_t& hi, uint16_t& lo) {
    slb_split2(v * SLB_ACT_PLANE_SCALE, fmt, hi, lo);
}
// Two values at once, fp16 planes: hi = the scaled value with its low 13 mantissa bits cleared (exact in fp16 above the
// subnormal range), lo = the exact remainder rounded to fp16 — two packed conversions per pair instead of four scalar ones
// and two conversions back. hi + lo carries the same 22 bits as slb_split2 (hi is truncated instead of rounded, so the
// planes differ in the last place of hi while their sum agrees to 2^-22); saturates like slb_split2, NaN goes through.
__device__ __forceinline__ void slb_split_pair_act_f16(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    x0 *= SLB_ACT_PLANE_SCALE;
    x1 *= SLB_ACT_PLANE_SCALE;
    if (x0 == x0) x0 = fminf(fmaxf(x0, -65504.0f), 65504.0f);
    if (x1 == x1) x1 = fminf(fmaxf(x1, -65504.0f), 65504.0f);
    const float h0 = __uint_as_float(__float_as_uint(x0) & 0xFFFFE000u);
    const float h1 = __uint_as_float(__float_as_uint(x1) & 0xFFFFE000u);
    const __half2 hp = __floats2half2_rn(h0, h1);
    const __half2 lp = __floats2half2_rn(x0 - h0, x1 - h1);
    hi = *reinterpret_cast<const uint32_t*>(&hp);
    lo = *reinterpret_cast<const uint32_t*>(&lp);
}
#endif  // __CUDACC__
