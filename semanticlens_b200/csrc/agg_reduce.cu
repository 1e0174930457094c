// K1 — aggregate hooked activation maps: (B,C,inner) -> (B,C)  or  (B,inner,C) -> (B,C).
//
// Replaces the reference aggregators (semanticlens/component_visualization/aggregators.py:38-244), which
// clone the map, reduce it with ATen and copy the result to the host. Here the map is read from HBM exactly
// once, staged through shared memory by 1-D bulk async copies (TMA, UBLKCP) behind an mbarrier ring, and
// reduced in fp32 in a FIXED order that depends on the row length only — so results are bit-identical for
// any batch size, grid size, alignment or code path (staged vs direct).
//
// Canonical reduction order (NCHW rows of L elements):
//   element e goes to accumulator (ws, lane) = ((e / 32) % 8, e % 32), each accumulator folds its elements in
//   increasing e starting from the identity (+0.0f / -inf); the 8 ws-accumulators of a lane are combined as
//   ((a0+a1)+(a2+a3))+((a4+a5)+(a6+a7)); the 32 lanes by an xor butterfly 16,8,4,2,1.
// Canonical order (BTF): the T tokens of a chain (b, f) are cut into blocks of 64; each block is folded sequentially from
//   the identity, the block partials are folded sequentially in block order (T <= 64: one plain sequential fold).
// mean = sum / (float)L (IEEE division), then rounded once to the input dtype (no-op for fp32).
#include "slb_common.cuh"

#include <stdlib.h>
#include <algorithm>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kStages = 3;
constexpr int kStageBytes = 32768;

// ------------------------------------------------------------------------------------------------
// element access / operators
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ float round_to_input(float v);
template <>
__device__ __forceinline__ float round_to_input<float>(float v) { return v; }
template <>
__device__ __forceinline__ float round_to_input<__half>(float v) { return __half2float(__float2half_rn(v)); }
template <>
__device__ __forceinline__ float round_to_input<__nv_bfloat16>(float v) {
    return slb_bf16_bits_to_f32(slb_f32_to_bf16_bits(v));
}

__device__ __forceinline__ float nanmax(float a, float b) {
    // torch.amax propagates NaN; fmaxf does not
    return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
}

template <int OP>
struct Agg {
    static constexpr bool kIsMax = (OP == SLB_AGG_MAX || OP == SLB_AGG_ABSMAX);
    static constexpr bool kAbs = (OP == SLB_AGG_ABSMEAN || OP == SLB_AGG_ABSMAX);
    __device__ static __forceinline__ float identity() { return kIsMax ? -INFINITY : 0.0f; }
    __device__ static __forceinline__ float pre(float v) { return kAbs ? fabsf(v) : v; }
    __device__ static __forceinline__ float fold(float a, float b) { return kIsMax ? nanmax(a, b) : (a + b); }
    __device__ static __forceinline__ float finish(float a, int64_t n) {
        return kIsMax ? a : __fdiv_rn(a, (float)n);
    }
};

template <int OP>
__device__ __forceinline__ float tree8(const float (&a)[8]) {
    using A = Agg<OP>;
    return A::fold(A::fold(A::fold(a[0], a[1]), A::fold(a[2], a[3])),
                   A::fold(A::fold(a[4], a[5]), A::fold(a[6], a[7])));
}

template <int OP>
__device__ __forceinline__ float butterfly(float v) {
    using A = Agg<OP>;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = A::fold(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// one warp folds one row that lives at `p` (shared or global memory), canonical order
template <typename T, int OP>
__device__ __forceinline__ float warp_row(const T* __restrict__ p, int L, int lane) {
    using A = Agg<OP>;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = A::identity();
    for (int base = 0; base < L; base += 256) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int e = base + 32 * j + lane;
            v[j] = (e < L) ? A::pre(to_f32<T>(p[e])) : A::identity();
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int e = base + 32 * j + lane;
            if (e < L) acc[j] = A::fold(acc[j], v[j]);
        }
    }
    return butterfly<OP>(tree8<OP>(acc));
}

// ------------------------------------------------------------------------------------------------
// NCHW, staged: persistent CTAs, 3-stage ring of 32 KB tiles filled by cp.async.bulk
//   mode SMALL: a tile is R whole rows, one warp folds one row at a time
//   mode LARGE: a tile is one <=32 KB chunk of one row, the 8 warps of the CTA share the row
// ------------------------------------------------------------------------------------------------
struct RowsParams {
    const void* x;
    float* out;
    int64_t n_rows;
    int L;
    int rows_per_tile;    // SMALL
    int64_t n_tiles;      // SMALL
    int chunk_elems;      // LARGE (multiple of 256)
    int chunks_per_row;   // LARGE
};

struct __align__(16) RowsSmem {
    uint64_t full[kStages];
    float partial[2][kWarps][32];
};

template <typename T, int OP, bool LARGE>
__global__ void __launch_bounds__(kThreads, 2) agg_rows_staged_kernel(RowsParams p) {
    using A = Agg<OP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;  // kStages * kStageBytes
    RowsSmem* ss = reinterpret_cast<RowsSmem*>(smem_raw + kStages * kStageBytes);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const T* x = static_cast<const T*>(p.x);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) slb_mbar_init(&ss->full[s], 1);
        slb_fence_mbar_init();
    }
    __syncthreads();

    // item = tile (SMALL) or (row, chunk) (LARGE); items of this CTA: it = blockIdx.x + i * gridDim.x (SMALL),
    // or rows r = blockIdx.x + j * gridDim.x, chunks in order (LARGE)
    int64_t n_items;
    if (LARGE) {
        int64_t my_rows = (p.n_rows > blockIdx.x) ? (p.n_rows - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        n_items = my_rows * p.chunks_per_row;
    } else {
        n_items = (p.n_tiles > blockIdx.x) ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    }

    auto issue = [&](int64_t i) {
        // called by thread 0 only
        int s = (int)(i % kStages);
        unsigned char* dst = stage_base + (size_t)s * kStageBytes;
        const T* src;
        int64_t elems;
        if (LARGE) {
            int64_t r = blockIdx.x + (i / p.chunks_per_row) * (int64_t)gridDim.x;
            int c = (int)(i % p.chunks_per_row);
            src = x + r * (int64_t)p.L + (int64_t)c * p.chunk_elems;
            elems = min((int64_t)p.chunk_elems, (int64_t)p.L - (int64_t)c * p.chunk_elems);
        } else {
            int64_t t = blockIdx.x + i * (int64_t)gridDim.x;
            int64_t r0 = t * p.rows_per_tile;
            int64_t nr = min((int64_t)p.rows_per_tile, p.n_rows - r0);
            src = x + r0 * (int64_t)p.L;
            elems = nr * p.L;
        }
        uint32_t bytes = (uint32_t)(elems * sizeof(T));
        uint32_t bulk = bytes & ~15u;
        // tail (< 16 bytes, only possible on the very last tile): plain copies by this thread, published by the
        // release semantics of the arrive below
        for (uint32_t o = bulk; o < bytes; o += sizeof(T))
            *reinterpret_cast<T*>(dst + o) = *reinterpret_cast<const T*>(reinterpret_cast<const unsigned char*>(src) + o);
        if (bulk) {
            slb_mbar_arrive_expect_tx(&ss->full[s], bulk);
            slb_bulk_g2s(dst, src, bulk, &ss->full[s]);
        } else {
            slb_mbar_arrive(&ss->full[s]);
        }
    };

    if (tid == 0) {
        for (int64_t i = 0; i < kStages && i < n_items; ++i) issue(i);
    }

    float acc = A::identity();  // LARGE: accumulator (warp, lane) of the current row
    int row_parity = 0;

    for (int64_t i = 0; i < n_items; ++i) {
        const int s = (int)(i % kStages);
        const uint32_t parity = (uint32_t)((i / kStages) & 1);
        slb_mbar_wait(&ss->full[s], parity);
        const T* tile = reinterpret_cast<const T*>(stage_base + (size_t)s * kStageBytes);

        if (LARGE) {
            const int64_t r = blockIdx.x + (i / p.chunks_per_row) * (int64_t)gridDim.x;
            const int c = (int)(i % p.chunks_per_row);
            const int elems = min(p.chunk_elems, p.L - c * p.chunk_elems);
            // chunk starts at a multiple of 256 within the row => local index keeps (ws, lane)
            for (int e0 = 32 * warp; e0 < elems; e0 += 256 * 4) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int e = e0 + 256 * j + lane;
                    v[j] = (e < elems) ? A::pre(to_f32<T>(tile[e])) : A::identity();
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int e = e0 + 256 * j + lane;
                    if (e < elems) acc = A::fold(acc, v[j]);
                }
            }
            const bool row_end = (c == p.chunks_per_row - 1);
            if (row_end) {
                ss->partial[row_parity][warp][lane] = acc;
                acc = A::identity();
            }
            __syncthreads();  // stage s is free again; partials (if any) are visible
            if (tid == 0 && i + kStages < n_items) issue(i + kStages);
            if (row_end) {
                if (warp == 0) {
                    float a[8];
#pragma unroll
                    for (int w = 0; w < 8; ++w) a[w] = ss->partial[row_parity][w][lane];
                    float rsum = butterfly<OP>(tree8<OP>(a));
                    if (lane == 0) p.out[r] = round_to_input<T>(A::finish(rsum, p.L));
                }
                row_parity ^= 1;
            }
        } else {
            const int64_t t = blockIdx.x + i * (int64_t)gridDim.x;
            const int64_t r0 = t * p.rows_per_tile;
            const int nr = (int)min((int64_t)p.rows_per_tile, p.n_rows - r0);
            for (int rr = warp; rr < nr; rr += kWarps) {
                float rsum = warp_row<T, OP>(tile + (size_t)rr * p.L, p.L, lane);
                if (lane == 0) p.out[r0 + rr] = round_to_input<T>(A::finish(rsum, p.L));
            }
            __syncthreads();
            if (tid == 0 && i + kStages < n_items) issue(i + kStages);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// NCHW, staged, SHORT rows (L <= 256, fp32): G = 8 or 16 lanes fold one row, so a warp folds 32/G rows at a time
// instead of idling most of its lanes (a 7x7 map is 49 elements). Bit-identical to the canonical order: lane g of a
// group owns the accumulator slots s = g, g + G, ... of the 32-slot layout; a slot's tree8 over its <= NWS elements
// (missing ones are the identity, folded at compile time), then the butterfly levels 16, 8 (, ...) >= G are local
// adds between the lane's own slots and the levels < G are shuffles inside the group.
// ------------------------------------------------------------------------------------------------
template <int OP, int G, int NWS>
__global__ void __launch_bounds__(kThreads, 2) agg_rows_subwarp_kernel(RowsParams p) {
    using A = Agg<OP>;
    constexpr int SPL = 32 / G;  // slots per lane
    constexpr int RPW = 32 / G;  // rows per warp per round
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;
    RowsSmem* ss = reinterpret_cast<RowsSmem*>(smem_raw + kStages * kStageBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane % G, sub = lane / G;
    const float* x = static_cast<const float*>(p.x);
    const int L = p.L;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) slb_mbar_init(&ss->full[s], 1);
        slb_fence_mbar_init();
    }
    __syncthreads();
    const int64_t n_items = (p.n_tiles > blockIdx.x) ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    auto issue = [&](int64_t i) {  // thread 0 only; tiles are whole rows and 16-byte multiples except the very last
        const int s = (int)(i % kStages);
        unsigned char* dst = stage_base + (size_t)s * kStageBytes;
        const int64_t t = blockIdx.x + i * (int64_t)gridDim.x;
        const int64_t r0 = t * p.rows_per_tile;
        const int64_t nr = min((int64_t)p.rows_per_tile, p.n_rows - r0);
        const float* src = x + r0 * (int64_t)L;
        const uint32_t bytes = (uint32_t)(nr * L * sizeof(float));
        const uint32_t bulk = bytes & ~15u;
        for (uint32_t o = bulk; o < bytes; o += 4)
            *reinterpret_cast<float*>(dst + o) = *reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(src) + o);
        if (bulk) {
            slb_mbar_arrive_expect_tx(&ss->full[s], bulk);
            slb_bulk_g2s(dst, src, bulk, &ss->full[s]);
        } else {
            slb_mbar_arrive(&ss->full[s]);
        }
    };
    if (tid == 0)
        for (int64_t i = 0; i < kStages && i < n_items; ++i) issue(i);

    for (int64_t i = 0; i < n_items; ++i) {
        const int s = (int)(i % kStages);
        slb_mbar_wait(&ss->full[s], (uint32_t)((i / kStages) & 1));
        const float* tile = reinterpret_cast<const float*>(stage_base + (size_t)s * kStageBytes);
        const int64_t r0 = (blockIdx.x + i * (int64_t)gridDim.x) * p.rows_per_tile;
        const int nr = (int)min((int64_t)p.rows_per_tile, p.n_rows - r0);
        for (int base = warp * RPW; base < nr; base += kWarps * RPW) {  // warp-uniform trip count (shuffles below)
            const int rr = base + sub;
            const bool valid = rr < nr;
            const float* row = tile + (size_t)(valid ? rr : base) * L;
            float slot[SPL];
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                const int sl = g + G * q;
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a[j] = A::identity();
                    if (j < NWS) {
                        const int e = 32 * j + sl;
                        if (e < L) a[j] = A::fold(A::identity(), A::pre(row[e]));
                    }
                }
                slot[q] = tree8<OP>(a);
            }
            // butterfly levels 16 .. G: between this lane's own slots (slot q <-> slot q + off / G)
#pragma unroll
            for (int off = 16; off >= G; off >>= 1) {
                const int d = off / G;
#pragma unroll
                for (int q = 0; q < SPL; ++q)
                    if ((q & d) == 0 && q + d < SPL) slot[q] = A::fold(slot[q], slot[q + d]);
            }
            float v = slot[0];
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) v = A::fold(v, __shfl_xor_sync(0xffffffffu, v, off));
            if (valid && g == 0) p.out[r0 + rr] = A::finish(v, L);
        }
        __syncthreads();
        if (tid == 0 && i + kStages < n_items) issue(i + kStages);
    }
}

// ------------------------------------------------------------------------------------------------
// NCHW, staged, MID rows (one row = 1/RPT of a stage, RPT = 2..4): like LARGE, the 8 warps share a row (warp w folds
// accumulator column ws = w), but a tile carries RPT whole rows, so the per-tile costs (barrier wait, __syncthreads,
// refill) are paid once per RPT rows and the bulk copies are RPT times larger.
// ------------------------------------------------------------------------------------------------
struct __align__(16) GroupSmem {
    uint64_t full[kStages];
    float partial[2][4][kWarps][32];
};

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads, 2) agg_rows_group_kernel(RowsParams p) {
    using A = Agg<OP>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* stage_base = smem_raw;
    GroupSmem* ss = reinterpret_cast<GroupSmem*>(smem_raw + kStages * kStageBytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const T* x = static_cast<const T*>(p.x);
    const int L = p.L, RPT = p.rows_per_tile;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) slb_mbar_init(&ss->full[s], 1);
        slb_fence_mbar_init();
    }
    __syncthreads();
    const int64_t n_items = (p.n_tiles > blockIdx.x) ? (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t i) {  // thread 0; row_bytes % 16 == 0, so every tile is a 16-byte multiple
        const int s = (int)(i % kStages);
        const int64_t r0 = (blockIdx.x + i * (int64_t)gridDim.x) * RPT;
        const int64_t nr = min((int64_t)RPT, p.n_rows - r0);
        const uint32_t bytes = (uint32_t)(nr * L * sizeof(T));
        slb_mbar_arrive_expect_tx(&ss->full[s], bytes);
        slb_bulk_g2s(stage_base + (size_t)s * kStageBytes, x + r0 * (int64_t)L, bytes, &ss->full[s]);
    };
    if (tid == 0)
        for (int64_t i = 0; i < kStages && i < n_items; ++i) issue(i);

    int parity = 0;
    for (int64_t i = 0; i < n_items; ++i) {
        const int s = (int)(i % kStages);
        slb_mbar_wait(&ss->full[s], (uint32_t)((i / kStages) & 1));
        const T* tile = reinterpret_cast<const T*>(stage_base + (size_t)s * kStageBytes);
        const int64_t r0 = (blockIdx.x + i * (int64_t)gridDim.x) * RPT;
        const int nr = (int)min((int64_t)RPT, p.n_rows - r0);
        for (int rr = 0; rr < nr; ++rr) {
            const T* row = tile + (size_t)rr * L;
            float acc = A::identity();
            for (int e0 = 32 * warp; e0 < L; e0 += 256 * 4) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = e0 + 256 * j + lane;
                    v[j] = (e < L) ? A::pre(to_f32<T>(row[e])) : A::identity();
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = e0 + 256 * j + lane;
                    if (e < L) acc = A::fold(acc, v[j]);
                }
            }
            ss->partial[parity][rr][warp][lane] = acc;
        }
        __syncthreads();  // the stage is free again; the partials are visible
        if (tid == 0 && i + kStages < n_items) issue(i + kStages);
        if (warp < nr) {
            float a[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) a[w] = ss->partial[parity][warp][w][lane];
            const float rsum = butterfly<OP>(tree8<OP>(a));
            if (lane == 0) p.out[r0 + warp] = round_to_input<T>(A::finish(rsum, L));
        }
        parity ^= 1;
    }
}

// ------------------------------------------------------------------------------------------------
// NCHW, direct from global (any alignment / row length). Same canonical order.
// ------------------------------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(kThreads) agg_rows_direct_warp_kernel(const T* __restrict__ x, float* __restrict__ out,
                                                                        int64_t n_rows, int L) {
    using A = Agg<OP>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < n_rows; r += (int64_t)gridDim.x * kWarps) {
        float rsum = warp_row<T, OP>(x + r * (int64_t)L, L, lane);
        if (lane == 0) out[r] = round_to_input<T>(A::finish(rsum, L));
    }
}

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads) agg_rows_direct_cta_kernel(const T* __restrict__ x, float* __restrict__ out,
                                                                       int64_t n_rows, int L) {
    using A = Agg<OP>;
    __shared__ float partial[kWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int64_t r = blockIdx.x; r < n_rows; r += gridDim.x) {
        const T* p = x + r * (int64_t)L;
        float acc = A::identity();
        for (int e0 = 32 * warp; e0 < L; e0 += 256 * 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int e = e0 + 256 * j + lane;
                v[j] = (e < L) ? A::pre(to_f32<T>(p[e])) : A::identity();
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int e = e0 + 256 * j + lane;
                if (e < L) acc = A::fold(acc, v[j]);
            }
        }
        partial[warp][lane] = acc;
        __syncthreads();
        if (warp == 0) {
            float a[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) a[w] = partial[w][lane];
            float rsum = butterfly<OP>(tree8<OP>(a));
            if (lane == 0) out[r] = round_to_input<T>(A::finish(rsum, L));
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// BTF: x (B, T, F) -> out (B, F) — token maps of transformers and the channels-last maps (B, H*W, C) of the accelerated
// probed forward. A chain (b, f) walks T rows that lie F elements apart, so the parallelism of a strictly sequential
// fold is B*F chains — 8192 for a 64-channel 112x112 map, far too few loads in flight for HBM. The canonical order is
// therefore two-level (shape-only, so still invariant to batch size / grid / code path):
//   the tokens are cut into blocks of 64; a block is folded sequentially from the identity, the block partials are
//   folded sequentially in block order (T <= 64: one plain sequential fold).
// CTA = (image, group of 32*VEC features), 4 warps; lane = VEC adjacent features (one 4..16-byte load per row, a warp reads
// 128..512 contiguous bytes of every row); warp w owns blocks w, w+4, ... and keeps 4 KB of rows
// in flight in registers (8..32 independent loads per lane, five CTAs per SM) before folding them in order; the partials meet in
// shared memory and warp 0 folds them in block order. No staging ring: every byte is read once, straight into registers.
// ------------------------------------------------------------------------------------------------
constexpr int kBtfThreads = 128;
constexpr int kBtfWarps = 4;
constexpr int kBtfBlock = 64;          // tokens per canonical block
constexpr int kBtfChunkFloats = 8192;  // partials staged per in-order combine (32 KB)

struct BtfParams {
    const void* x;
    float* out;
    int64_t B;
    int T;
    int64_t F;
    int groups;  // ceil(F / (32 * VEC)) feature groups per image
};

// raw per-lane load of BYTES bytes kept as 32-bit words (so that the rows in flight provably stay in registers)
template <int BYTES>
struct BtfRaw;
template <>
struct BtfRaw<2> {
    uint32_t w[1];
    __device__ __forceinline__ void load(const void* q) { w[0] = *static_cast<const uint16_t*>(q); }
};
template <>
struct BtfRaw<4> {
    uint32_t w[1];
    __device__ __forceinline__ void load(const void* q) { w[0] = *static_cast<const uint32_t*>(q); }
};
template <>
struct BtfRaw<8> {
    uint32_t w[2];
    __device__ __forceinline__ void load(const void* q) {
        const uint2 t = *static_cast<const uint2*>(q);
        w[0] = t.x; w[1] = t.y;
    }
};
template <>
struct BtfRaw<16> {
    uint32_t w[4];
    __device__ __forceinline__ void load(const void* q) {
        const uint4 t = *static_cast<const uint4*>(q);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
    }
};

template <typename T>
__device__ __forceinline__ float btf_elem(const uint32_t* w, int e);
template <>
__device__ __forceinline__ float btf_elem<float>(const uint32_t* w, int e) { return __uint_as_float(w[e]); }
template <>
__device__ __forceinline__ float btf_elem<__half>(const uint32_t* w, int e) {
    return __half2float(__ushort_as_half((unsigned short)(w[e >> 1] >> (16 * (e & 1)))));
}
template <>
__device__ __forceinline__ float btf_elem<__nv_bfloat16>(const uint32_t* w, int e) {
    return __uint_as_float((w[e >> 1] >> (16 * (e & 1))) << 16);
}

template <typename T, int OP, int VEC>
__global__ void __launch_bounds__(kBtfThreads, 5) agg_btf_kernel(BtfParams p) {
    using A = Agg<OP>;
    constexpr int kBytes = (int)sizeof(T) * VEC;
    using Raw = BtfRaw<kBytes>;
    constexpr int kRows = kBytes <= 4 ? 32 : 128 / kBytes;  // rows in flight per lane: 32 / 16 / 8 (32 registers, 4 KB per warp)
    constexpr int kChunkBlocks = kBtfChunkFloats / (32 * VEC);
    extern __shared__ __align__(16) float btf_part[];  // [min(n_blocks, kChunkBlocks)][32 * VEC]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x / p.groups;
    const int grp = (int)(blockIdx.x % p.groups);
    const int64_t f = (int64_t)grp * (32 * VEC) + (int64_t)lane * VEC;
    const bool active = f < p.F;  // F % VEC == 0: a lane's features are all inside or all outside
    const T* src = static_cast<const T*>(p.x) + (b * p.T) * p.F + (active ? f : 0);
    const int n_blocks = (p.T + kBtfBlock - 1) / kBtfBlock;

    float total[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) total[e] = A::identity();

    for (int c0 = 0; c0 < n_blocks; c0 += kChunkBlocks) {
        const int nb = min(kChunkBlocks, n_blocks - c0);
        for (int j = warp; j < nb; j += kBtfWarps) {
            const int t0 = (c0 + j) * kBtfBlock;
            float acc[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] = A::identity();
            if (active) {
                const T* q = src + (int64_t)t0 * p.F;
                if (t0 + kBtfBlock <= p.T) {
#pragma unroll 1
                    for (int h = 0; h < kBtfBlock / kRows; ++h) {
                        Raw v[kRows];
#pragma unroll
                        for (int u = 0; u < kRows; ++u) v[u].load(q + (int64_t)(h * kRows + u) * p.F);
#pragma unroll
                        for (int u = 0; u < kRows; ++u)
#pragma unroll
                            for (int e = 0; e < VEC; ++e) acc[e] = A::fold(acc[e], A::pre(btf_elem<T>(v[u].w, e)));
                    }
                } else {  // the last, partial block: the bound is uniform over the warp
                    const int nt = p.T - t0;
#pragma unroll 1
                    for (int h = 0; h * kRows < nt; ++h) {
                        Raw v[kRows];
#pragma unroll
                        for (int u = 0; u < kRows; ++u)
                            if (h * kRows + u < nt) v[u].load(q + (int64_t)(h * kRows + u) * p.F);
#pragma unroll
                        for (int u = 0; u < kRows; ++u)
                            if (h * kRows + u < nt) {
#pragma unroll
                                for (int e = 0; e < VEC; ++e) acc[e] = A::fold(acc[e], A::pre(btf_elem<T>(v[u].w, e)));
                            }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) btf_part[(size_t)j * (32 * VEC) + lane * VEC + e] = acc[e];
        }
        __syncthreads();
        if (warp == 0) {
            for (int j = 0; j < nb; ++j) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) total[e] = A::fold(total[e], btf_part[(size_t)j * (32 * VEC) + lane * VEC + e]);
            }
        }
        __syncthreads();
    }
    if (warp == 0 && active) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) p.out[b * p.F + f + e] = round_to_input<T>(A::finish(total[e], p.T));
    }
}

template <typename T>
__global__ void token_select_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t B, int64_t T_,
                                    int64_t F, int64_t pos) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * F) return;
    int64_t b = i / F, f = i % F;
    out[i] = to_f32<T>(x[(b * T_ + pos) * F + f]);
}

// ------------------------------------------------------------------------------------------------
// host dispatch
// ------------------------------------------------------------------------------------------------
int g_force_direct = -1;  // -1: read env once

bool force_direct() {
    if (g_force_direct < 0) {
        const char* e = getenv("SLB_AGG_DIRECT");
        g_force_direct = (e && e[0] == '1') ? 1 : 0;
    }
    return g_force_direct == 1;
}

template <typename T, int OP>
int launch_rows(const void* x, float* out, int64_t n_rows, int64_t L, cudaStream_t st) {
    const int sms = slb_sm_count();
    const size_t esz = sizeof(T);
    const bool aligned = ((uintptr_t)x % 16) == 0;
    // rows per tile must keep every tile start 16-byte aligned
    int64_t row_bytes = L * (int64_t)esz;
    int64_t g = 16;
    {  // gcd(16, row_bytes)
        int64_t a = 16, b = row_bytes % 16;
        while (b) { int64_t t = a % b; a = b; b = t; }
        g = a;
    }
    const int64_t align_r = 16 / g;
    int64_t R = (kStageBytes / row_bytes) / align_r * align_r;
    const size_t smem = (size_t)kStages * kStageBytes + sizeof(RowsSmem);

    if constexpr (sizeof(T) == 4) {
        // short rows: sub-warp groups (fp32 maps; L <= 256: one 256-element block of the canonical order)
        if (aligned && !force_direct() && R >= kWarps && L <= 256) {
            RowsParams p{};
            p.x = x; p.out = out; p.n_rows = n_rows; p.L = (int)L;
            p.rows_per_tile = (int)R;
            p.n_tiles = slb_ceil_div(n_rows, R);
            const int grid = (int)std::min<int64_t>(p.n_tiles, (int64_t)sms * 2);
            const int nws = (int)slb_ceil_div(L, 32);
#define SLB_SUBWARP(G_, NWS_)                                                                                         \
    {                                                                                                                 \
        auto kern = agg_rows_subwarp_kernel<OP, G_, NWS_>;                                                            \
        SLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        kern<<<grid, kThreads, smem, st>>>(p);                                                                        \
    }
            if (nws == 1) SLB_SUBWARP(8, 1)
            else if (nws == 2) SLB_SUBWARP(8, 2)
            else if (nws == 3) SLB_SUBWARP(16, 3)
            else if (nws == 4) SLB_SUBWARP(16, 4)
            else if (nws == 5) SLB_SUBWARP(16, 5)
            else if (nws == 6) SLB_SUBWARP(16, 6)
            else if (nws == 7) SLB_SUBWARP(16, 7)
            else SLB_SUBWARP(16, 8)
#undef SLB_SUBWARP
            SLB_LAUNCH_OK("agg_rows_subwarp");
            return SLB_OK;
        }
    }
    {
        // mid rows: 2..4 whole rows per stage, all 8 warps on each row
        const int64_t rpt = std::min<int64_t>(4, kStageBytes / row_bytes);
        if (aligned && !force_direct() && (row_bytes % 16) == 0 && L >= 1024 && rpt >= 2) {
            RowsParams p{};
            p.x = x; p.out = out; p.n_rows = n_rows; p.L = (int)L;
            p.rows_per_tile = (int)rpt;
            p.n_tiles = slb_ceil_div(n_rows, rpt);
            auto kern = agg_rows_group_kernel<T, OP>;
            const size_t gsmem = (size_t)kStages * kStageBytes + sizeof(GroupSmem);
            SLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
            const int grid = (int)std::min<int64_t>(p.n_tiles, (int64_t)sms * 2);
            kern<<<grid, kThreads, gsmem, st>>>(p);
            SLB_LAUNCH_OK("agg_rows_group");
            return SLB_OK;
        }
    }
    if (aligned && !force_direct() && R >= kWarps) {
        RowsParams p{};
        p.x = x; p.out = out; p.n_rows = n_rows; p.L = (int)L;
        p.rows_per_tile = (int)R;
        p.n_tiles = slb_ceil_div(n_rows, R);
        auto kern = agg_rows_staged_kernel<T, OP, false>;
        SLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (int)std::min<int64_t>(p.n_tiles, (int64_t)sms * 2);
        kern<<<grid, kThreads, smem, st>>>(p);
        SLB_LAUNCH_OK("agg_rows_staged<small>");
        return SLB_OK;
    }
    if (aligned && !force_direct() && (row_bytes % 16) == 0 && L >= 1024) {
        RowsParams p{};
        p.x = x; p.out = out; p.n_rows = n_rows; p.L = (int)L;
        p.chunk_elems = (int)(kStageBytes / esz);  // multiple of 256
        p.chunks_per_row = (int)slb_ceil_div(L, p.chunk_elems);
        auto kern = agg_rows_staged_kernel<T, OP, true>;
        SLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (int)std::min<int64_t>(n_rows, (int64_t)sms * 2);
        kern<<<grid, kThreads, smem, st>>>(p);
        SLB_LAUNCH_OK("agg_rows_staged<large>");
        return SLB_OK;
    }
    if (L >= 2048) {
        int grid = (int)std::min<int64_t>(n_rows, (int64_t)sms * 8);
        agg_rows_direct_cta_kernel<T, OP><<<grid, kThreads, 0, st>>>(static_cast<const T*>(x), out, n_rows, (int)L);
        SLB_LAUNCH_OK("agg_rows_direct_cta");
    } else {
        int grid = (int)std::min<int64_t>(slb_ceil_div(n_rows, kWarps), (int64_t)sms * 8);
        agg_rows_direct_warp_kernel<T, OP><<<grid, kThreads, 0, st>>>(static_cast<const T*>(x), out, n_rows, (int)L);
        SLB_LAUNCH_OK("agg_rows_direct_warp");
    }
    return SLB_OK;
}

template <typename T, int OP, int VEC>
int launch_btf_vec(const BtfParams& p0, cudaStream_t st) {
    BtfParams p = p0;
    p.groups = (int)slb_ceil_div(p.F, (int64_t)32 * VEC);
    const int64_t grid = p.B * p.groups;
    SLB_REQUIRE(grid <= 0x7FFFFFFF, SLB_EUNSUPPORTED, "agg_btf: grid too large");
    const int n_blocks = (p.T + kBtfBlock - 1) / kBtfBlock;
    const size_t smem = (size_t)std::min(n_blocks, kBtfChunkFloats / (32 * VEC)) * (32 * VEC) * sizeof(float);
    agg_btf_kernel<T, OP, VEC><<<(int)grid, kBtfThreads, smem, st>>>(p);
    SLB_LAUNCH_OK("agg_btf");
    return SLB_OK;
}

template <typename T, int OP>
int launch_btf(const void* x, float* out, int64_t B, int64_t T_, int64_t F, cudaStream_t st) {
    BtfParams p{};
    p.x = x; p.out = out; p.B = B; p.T = (int)T_; p.F = F;
    // widest per-lane vector (16 bytes at most) that divides F, keeps every row start aligned and still leaves at least
    // two CTAs per SM; narrower vectors mean more, smaller CTAs
    const int64_t want = 2 * (int64_t)slb_sm_count();
    constexpr int kMaxVec = 16 / (int)sizeof(T);
    int vec = 1;
    for (int v = kMaxVec; v > 1; v >>= 1) {
        const bool fits = (F % v) == 0 && ((uintptr_t)x % (v * sizeof(T))) == 0;
        if (fits && B * slb_ceil_div(F, (int64_t)32 * v) >= want) { vec = v; break; }
    }
    switch (vec) {
        case 8:
            if constexpr (kMaxVec >= 8) return launch_btf_vec<T, OP, 8>(p, st);
        case 4: return launch_btf_vec<T, OP, 4>(p, st);
        case 2: return launch_btf_vec<T, OP, 2>(p, st);
        default: return launch_btf_vec<T, OP, 1>(p, st);
    }
}

template <typename T>
int dispatch_op(const void* x, int layout, int64_t B, int64_t C, int64_t inner, int op, int64_t token_pos, float* out,
                cudaStream_t st) {
    if (op == SLB_AGG_TOKEN) {
        SLB_REQUIRE(layout == SLB_LAYOUT_BTF, SLB_EINVAL, "SLB_AGG_TOKEN needs the (B, T, F) layout");
        int64_t pos = token_pos < 0 ? token_pos + inner : token_pos;
        SLB_REQUIRE(pos >= 0 && pos < inner, SLB_EINVAL, "token position %lld out of range for %lld tokens",
                    (long long)token_pos, (long long)inner);
        int64_t n = B * C;
        token_select_kernel<T><<<(unsigned)slb_ceil_div(n, 256), 256, 0, st>>>(static_cast<const T*>(x), out, B, inner,
                                                                                C, pos);
        SLB_LAUNCH_OK("token_select");
        return SLB_OK;
    }
#define SLB_CASE(OPV)                                                             \
    case OPV:                                                                     \
        return layout == SLB_LAYOUT_NCHW ? launch_rows<T, OPV>(x, out, B * C, inner, st) \
                                         : launch_btf<T, OPV>(x, out, B, inner, C, st);
    switch (op) {
        SLB_CASE(SLB_AGG_MEAN)
        SLB_CASE(SLB_AGG_MAX)
        SLB_CASE(SLB_AGG_ABSMEAN)
        SLB_CASE(SLB_AGG_ABSMAX)
        default:
            break;
    }
#undef SLB_CASE
    slb_set_error("unknown aggregation op %d", op);
    return SLB_EINVAL;
}

}  // namespace

extern "C" int slb_agg_reduce(const void* x, int dtype, int layout, int64_t B, int64_t C, int64_t inner, int agg_op,
                              int64_t token_pos, float* out, void* stream) {
    SLB_REQUIRE(B >= 0 && C >= 0 && inner >= 0, SLB_EINVAL, "slb_agg_reduce: negative size");
    SLB_REQUIRE(layout == SLB_LAYOUT_NCHW || layout == SLB_LAYOUT_BTF, SLB_EINVAL, "slb_agg_reduce: bad layout %d",
                layout);
    if (B == 0 || C == 0) return SLB_OK;
    SLB_REQUIRE(x != nullptr && out != nullptr, SLB_EINVAL, "slb_agg_reduce: null pointer");
    SLB_REQUIRE(inner >= 1, SLB_EINVAL, "slb_agg_reduce: empty reduction axis");
    SLB_REQUIRE(inner < (1ll << 31) - 4096, SLB_EUNSUPPORTED, "slb_agg_reduce: reduction axis too long");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const double esz = dtype == SLB_DT_F32 ? 4.0 : 2.0;
    SlbProfScope prof("K1 agg_reduce", stream, 0.0, (double)B * (double)C * (double)inner * esz);
    switch (dtype) {
        case SLB_DT_F32:
            return dispatch_op<float>(x, layout, B, C, inner, agg_op, token_pos, out, st);
        case SLB_DT_F16:
            return dispatch_op<__half>(x, layout, B, C, inner, agg_op, token_pos, out, st);
        case SLB_DT_BF16:
            return dispatch_op<__nv_bfloat16>(x, layout, B, C, inner, agg_op, token_pos, out, st);
        default:
            slb_set_error("slb_agg_reduce: bad dtype %d", dtype);
            return SLB_EINVAL;
    }
}
