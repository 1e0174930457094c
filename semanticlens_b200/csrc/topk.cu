// K2 — streaming per-latent top-k.
//
// Replaces ActMax.update (semanticlens/component_visualization/activation_caching.py:112-141): the reference
// transposes the (B, C) aggregate, casts to bf16, concatenates it with the (C, k) state, runs torch.topk on the
// CPU and gathers the ids. Here one warp owns one latent: it encodes (bf16 value, id) pairs as order-preserving
// u64 keys, drops every candidate that cannot beat the current k-th key, and bitonic-sorts the survivors together
// with the state in shared memory. The order is canonical — (value desc, id asc, placeholders last) — so the state
// is invariant to batch size and to how images are sharded across ranks.
#include "slb_common.cuh"

#include <algorithm>

namespace {

constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t* keys, int P, int lane) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = lane; i < (P >> 1); i += 32) {
                int lo = 2 * i - (i & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = keys[lo], b = keys[hi];
                if ((a < b) == desc) {
                    keys[lo] = b;
                    keys[hi] = a;
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ int next_pow2(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

template <typename CT>
__device__ __forceinline__ uint16_t cand_bits(CT v);
template <>
__device__ __forceinline__ uint16_t cand_bits<float>(float v) { return slb_f32_to_bf16_bits(v); }
template <>
__device__ __forceinline__ uint16_t cand_bits<uint16_t>(uint16_t v) {
    return ((v & 0x7FFFu) > 0x7F80u) ? (uint16_t)0x7FC0 : v;
}

// cand (B, C) row-major. One warp per latent c; `P` u64 slots of dynamic smem per warp (>= next_pow2(k + B)).
template <typename CT>
__global__ void __launch_bounds__(kWarpsPerCta * 32) topk_update_kernel(const CT* __restrict__ cand, int64_t B, int64_t C,
                                                                        const int64_t* __restrict__ ids, int64_t id_base,
                                                                        uint16_t* __restrict__ svals,
                                                                        int64_t* __restrict__ sids, int k, int P,
                                                                        int warps_per_cta) {
    extern __shared__ __align__(16) uint64_t keys_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * warps_per_cta + warp;
    if (warp >= warps_per_cta || c >= C) return;
    uint64_t* keys = keys_all + (size_t)warp * P;

    // state -> keys; threshold = smallest state key (robust to states written in a different tie order)
    uint64_t thr = ~0ull;
    for (int i = lane; i < k; i += 32) {
        uint64_t key = slb_topk_key(svals[c * k + i], sids[c * k + i]);
        keys[i] = key;
        thr = min(thr, key);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) thr = min(thr, __shfl_xor_sync(0xffffffffu, thr, o));

    int n = k;
    for (int64_t b0 = 0; b0 < B; b0 += 32) {
        int64_t b = b0 + lane;
        uint64_t key = 0;
        if (b < B) {
            uint16_t bits = cand_bits<CT>(cand[b * C + c]);
            int64_t id = ids ? ids[b] : (id_base + b);
            key = slb_topk_key(bits, id);
        }
        bool keep = (b < B) && (key > thr);
        unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) keys[n + __popc(m & ((1u << lane) - 1u))] = key;
        n += __popc(m);
    }
    if (n == k) return;  // nothing beats the current k-th: state unchanged (the steady-state case)

    const int Pn = next_pow2(n);
    for (int i = n + lane; i < Pn; i += 32) keys[i] = 0ull;  // below every real key
    __syncwarp();
    warp_bitonic_sort_desc(keys, Pn, lane);
    for (int i = lane; i < k; i += 32) {
        uint16_t bits;
        int64_t id;
        slb_topk_unkey(keys[i], &bits, &id);
        svals[c * k + i] = bits;
        sids[c * k + i] = id;
    }
}

// vals/ids (R, C, k) -> out (C, k)
__global__ void __launch_bounds__(kWarpsPerCta * 32) topk_merge_lists_kernel(const uint16_t* __restrict__ vals,
                                                                             const int64_t* __restrict__ ids, int R,
                                                                             int64_t C, int k, uint16_t* __restrict__ ovals,
                                                                             int64_t* __restrict__ oids, int P,
                                                                             int warps_per_cta) {
    extern __shared__ __align__(16) uint64_t keys_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * warps_per_cta + warp;
    if (warp >= warps_per_cta || c >= C) return;
    uint64_t* keys = keys_all + (size_t)warp * P;
    const int n = R * k;
    for (int i = lane; i < P; i += 32) {
        uint64_t key = 0ull;
        if (i < n) {
            int r = i / k, j = i % k;
            int64_t off = ((int64_t)r * C + c) * k + j;
            key = slb_topk_key(vals[off], ids[off]);
        }
        keys[i] = key;
    }
    __syncwarp();
    warp_bitonic_sort_desc(keys, P, lane);
    for (int i = lane; i < k; i += 32) {
        uint16_t bits;
        int64_t id;
        slb_topk_unkey(keys[i], &bits, &id);
        ovals[c * k + i] = bits;
        oids[c * k + i] = id;
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ table, int64_t N, int64_t D, const int64_t* __restrict__ idx,
                                   int64_t n_idx, float* __restrict__ out, int* __restrict__ err) {
    // one warp per output row; float4 when D % 4 == 0
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_idx) return;
    int64_t i = idx[r];
    if (i < 0) i += N;
    if (i < 0 || i >= N) {
        if (err && lane == 0) *err = 1;
        return;
    }
    const float* src = table + i * D;
    float* dst = out + r * D;
    if ((D & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int64_t j = lane; j < (D >> 2); j += 32) d4[j] = s4[j];
    } else {
        for (int64_t j = lane; j < D; j += 32) dst[j] = src[j];
    }
}

int pick_warps(int P, size_t* smem_out) {
    // keep dynamic smem per CTA <= 64 KB
    int w = kWarpsPerCta;
    while (w > 1 && (size_t)w * P * 8 > 65536) w >>= 1;
    *smem_out = (size_t)w * P * 8;
    return w;
}

int host_next_pow2(int64_t n) {
    int p = 32;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace

extern "C" int slb_topk_update(const void* cand, int cand_dtype, int64_t B, int64_t C, const int64_t* ids, int64_t id_base,
                               uint16_t* state_vals, int64_t* state_ids, int64_t k, void* stream) {
    SLB_REQUIRE(B >= 0 && C >= 0 && k >= 0, SLB_EINVAL, "slb_topk_update: negative size");
    if (B == 0 || C == 0 || k == 0) return SLB_OK;  // k = 0 is legal (reference tests/…/test_activation_based.py:142-161)
    SLB_REQUIRE(cand && state_vals && state_ids, SLB_EINVAL, "slb_topk_update: null pointer");
    SLB_REQUIRE(cand_dtype == SLB_DT_F32 || cand_dtype == SLB_DT_BF16, SLB_EINVAL,
                "slb_topk_update: candidates must be fp32 or bf16");
    SLB_REQUIRE(k + B <= 8192, SLB_EUNSUPPORTED, "slb_topk_update: k + B = %lld exceeds 8192 (split the batch)",
                (long long)(k + B));
    SLB_REQUIRE(ids != nullptr || (id_base >= 0 && id_base + B < (int64_t)SLB_ID_MASK), SLB_EINVAL,
                "slb_topk_update: sample ids must lie in [0, 2^47-1)");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SlbProfScope prof("K2 topk_update", stream, 0.0,
                      (double)B * (double)C * (cand_dtype == SLB_DT_F32 ? 4.0 : 2.0) + 2.0 * (double)C * (double)k * 10.0);
    const int P = host_next_pow2(k + B);
    size_t smem;
    const int w = pick_warps(P, &smem);
    const unsigned grid = (unsigned)slb_ceil_div(C, w);
    if (cand_dtype == SLB_DT_F32) {
        auto kern = topk_update_kernel<float>;
        SLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        kern<<<grid, w * 32, smem, st>>>(static_cast<const float*>(cand), B, C, ids, id_base, state_vals, state_ids,
                                         (int)k, P, w);
    } else {
        auto kern = topk_update_kernel<uint16_t>;
        SLB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        kern<<<grid, w * 32, smem, st>>>(static_cast<const uint16_t*>(cand), B, C, ids, id_base, state_vals, state_ids,
                                         (int)k, P, w);
    }
    SLB_LAUNCH_OK("topk_update");
    return SLB_OK;
}

extern "C" int slb_topk_merge_lists(const uint16_t* vals, const int64_t* ids, int64_t R, int64_t C, int64_t k,
                                    uint16_t* out_vals, int64_t* out_ids, void* stream) {
    SLB_REQUIRE(R >= 1 && C >= 0 && k >= 0, SLB_EINVAL, "slb_topk_merge_lists: bad size");
    if (C == 0 || k == 0) return SLB_OK;
    SLB_REQUIRE(vals && ids && out_vals && out_ids, SLB_EINVAL, "slb_topk_merge_lists: null pointer");
    SLB_REQUIRE((const void*)vals != (const void*)out_vals && (const void*)ids != (const void*)out_ids, SLB_EINVAL,
                "slb_topk_merge_lists: outputs may not alias inputs");
    SLB_REQUIRE(R * k <= 8192, SLB_EUNSUPPORTED, "slb_topk_merge_lists: R*k = %lld exceeds 8192", (long long)(R * k));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SlbProfScope prof("K2 topk_merge_lists", stream, 0.0, (double)(R + 1) * (double)C * (double)k * 10.0);
    const int P = host_next_pow2(R * k);
    size_t smem;
    const int w = pick_warps(P, &smem);
    SLB_CUDA_OK(cudaFuncSetAttribute(topk_merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    topk_merge_lists_kernel<<<(unsigned)slb_ceil_div(C, w), w * 32, smem, st>>>(vals, ids, (int)R, C, (int)k, out_vals,
                                                                               out_ids, P, w);
    SLB_LAUNCH_OK("topk_merge_lists");
    return SLB_OK;
}

extern "C" int slb_agg_topk_update(const void* x, int dtype, int layout, int64_t B, int64_t C, int64_t inner, int agg_op,
                                   int64_t token_pos, int64_t id_base, uint16_t* state_vals, int64_t* state_ids,
                                   int64_t k, void* scratch, size_t scratch_bytes, void* stream) {
    SLB_REQUIRE(B >= 0 && C >= 0, SLB_EINVAL, "slb_agg_topk_update: negative size");
    if (B == 0 || C == 0) return SLB_OK;
    SLB_REQUIRE(scratch != nullptr && scratch_bytes >= (size_t)(B * C) * sizeof(float), SLB_EWORKSPACE,
                "slb_agg_topk_update: scratch needs %lld bytes", (long long)(B * C * 4));
    int rc = slb_agg_reduce(x, dtype, layout, B, C, inner, agg_op, token_pos, static_cast<float*>(scratch), stream);
    if (rc != SLB_OK) return rc;
    return slb_topk_update(scratch, SLB_DT_F32, B, C, nullptr, id_base, state_vals, state_ids, k, stream);
}

extern "C" int slb_gather_rows(const float* table, int64_t N, int64_t D, const int64_t* idx, int64_t n_idx, float* out,
                               void* stream) {
    SLB_REQUIRE(N >= 0 && D >= 0 && n_idx >= 0, SLB_EINVAL, "slb_gather_rows: negative size");
    if (n_idx == 0 || D == 0) return SLB_OK;
    SLB_REQUIRE(table && idx && out && N > 0, SLB_EINVAL, "slb_gather_rows: null pointer or empty table");
    SLB_REQUIRE(((uintptr_t)table % 16) == 0 && ((uintptr_t)out % 16) == 0, SLB_EINVAL,
                "slb_gather_rows: buffers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SlbProfScope prof("K5 gather_rows", stream, 0.0, (double)n_idx * ((double)D * 8.0 + 8.0));
    const int threads = 256;
    gather_rows_kernel<<<(unsigned)slb_ceil_div(n_idx * 32, threads), threads, 0, st>>>(table, N, D, idx, n_idx, out,
                                                                                       nullptr);
    SLB_LAUNCH_OK("gather_rows");
    return SLB_OK;
}
