// K6/K7 — the score kernels behind semanticlens/scores.py (reference), all HBM-bound streaming passes except the
// cosine matmul, which is one slb_gemm_split on the tensor cores.
//
//   slb_normalize_split_rows  F.normalize(x, dim=-1) (scores.py:120-121) fused with the split-plane conversion the
//                             tensor-core GEMM consumes: one read of x, no normalised fp32 copy
//   slb_cosine_gemm           similarity_score's `x_.matmul(y_.T)` (scores.py:124-125): normalise+split both operands,
//                             then tcgen05 GEMM, fp32 out
//   slb_cosine_rows           the equal-shape branch, F.cosine_similarity(x, y, dim=-1) (scores.py:127)
//   slb_clarity               clarity_score (scores.py:45-46): ((|mean_k normalize(V)|^2 - 1/k) / (k-1)) * k in ONE pass
//                             over V (the reference materialises the normalised copy: 2x the traffic)
#include "tc_common.cuh"

#include <algorithm>

namespace {

constexpr int kWarpsPerCta = 8;

__device__ __forceinline__ float sumsq4(const float4& v) { return (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w); }

// F.normalize: x / max(||x||_2, eps). torch's clamp_min keeps a NaN norm (the whole row becomes NaN); fmaxf would drop it.
__device__ __forceinline__ float inv_norm(float ss, float eps) {
    const float nrm = sqrtf(ss);
    return 1.0f / (nrm != nrm ? nrm : fmaxf(nrm, eps));
}

// ------------------------------------------------------------------------------------------------
// one warp per row: inverse norm, then planes of x * inv (zero padded from D up to Kpad columns)
// ------------------------------------------------------------------------------------------------
template <int NCH>  // row chunks of 128 floats kept in registers (D <= 128 * NCH)
__global__ void __launch_bounds__(kWarpsPerCta * 32)
normalize_split_kernel(const float* __restrict__ x, int64_t rows, int D, int Kpad, float eps, int fmt, float plane_scale,
                       uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, float* __restrict__ inv_out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + r * (int64_t)D);
    const int d4 = D >> 2;
    float4 v[NCH];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int i = c * 32 + lane;
        v[c] = i < d4 ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        ss += sumsq4(v[c]);
    }
    const float inv = inv_norm(slb_warp_sum_butterfly(ss), eps);
    if (inv_out && lane == 0) inv_out[r] = inv;
    if (!hi) return;
    const float sc = inv * plane_scale;
    const int k4 = Kpad >> 2;
    uint2* h2 = reinterpret_cast<uint2*>(hi + r * (int64_t)Kpad);
    uint2* l2 = reinterpret_cast<uint2*>(lo + r * (int64_t)Kpad);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int i = c * 32 + lane;
        if (i < k4) {
            uint16_t h[4], l[4];
            slb_split2(v[c].x * sc, fmt, h[0], l[0]);
            slb_split2(v[c].y * sc, fmt, h[1], l[1]);
            slb_split2(v[c].z * sc, fmt, h[2], l[2]);
            slb_split2(v[c].w * sc, fmt, h[3], l[3]);
            h2[i] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
            l2[i] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// F.cosine_similarity(x, y, dim=-1, eps): sum_d (x/max(|x|,eps)) * (y/max(|y|,eps)); one warp per row pair
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32)
cosine_rows_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t rows, int D, float eps,
                   float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* xr = x + r * (int64_t)D;
    const float* yr = y + r * (int64_t)D;
    float sx = 0.f, sy = 0.f;
    for (int i = lane; i < D; i += 32) {
        const float a = xr[i], b = yr[i];
        sx = fmaf(a, a, sx);
        sy = fmaf(b, b, sy);
    }
    const float ix = inv_norm(slb_warp_sum_butterfly(sx), eps), iy = inv_norm(slb_warp_sum_butterfly(sy), eps);
    float dot = 0.f;
    for (int i = lane; i < D; i += 32) dot = fmaf(xr[i] * ix, yr[i] * iy, dot);
    dot = slb_warp_sum_butterfly(dot);
    if (lane == 0) out[r] = dot;
}

// ------------------------------------------------------------------------------------------------
// K7 clarity: one CTA per neuron, 8 warps stride over its k rows; a lane keeps its slice of the running sum of
// normalised rows in registers. The k rows (k*D*4 bytes) are read exactly once, 512 B per warp-load instruction.
// ------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
clarity_kernel(const float* __restrict__ V, int64_t n_neurons, int k, int D, float eps, float* __restrict__ out) {
    extern __shared__ float4 part[];  // [kWarpsPerCta][NCH * 32]
    __shared__ double red[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d4 = D >> 2;
    for (int64_t n = blockIdx.x; n < n_neurons; n += gridDim.x) {
        const float4* base = reinterpret_cast<const float4*>(V + n * (int64_t)k * D);
        float4 acc[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = warp; r < k; r += kWarpsPerCta) {
            const float4* xr = base + (int64_t)r * d4;
            float4 v[NCH];
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int i = c * 32 + lane;
                v[c] = i < d4 ? __ldcs(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                ss += sumsq4(v[c]);
            }
            const float inv = inv_norm(slb_warp_sum_butterfly(ss), eps);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                acc[c].x = fmaf(v[c].x, inv, acc[c].x);
                acc[c].y = fmaf(v[c].y, inv, acc[c].y);
                acc[c].z = fmaf(v[c].z, inv, acc[c].z);
                acc[c].w = fmaf(v[c].w, inv, acc[c].w);
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) part[warp * NCH * 32 + c * 32 + lane] = acc[c];
        __syncthreads();
        // |sum over warps|^2, in double from here on (the affine map below cancels to ~0 for unclear neurons)
        double q = 0.0;
        for (int i = threadIdx.x; i < NCH * 32; i += blockDim.x) {
            float4 s = part[i];
#pragma unroll
            for (int w = 1; w < kWarpsPerCta; ++w) {
                const float4 t = part[w * NCH * 32 + i];
                s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
            }
            // mean over k first (the reference squares the fp32 mean)
            const float ik = 1.0f / (float)k;
            const float mx = s.x * ik, my = s.y * ik, mz = s.z * ik, mw = s.w * ik;
            q += (double)(mx * mx) + (double)(my * my) + (double)(mz * mz) + (double)(mw * mw);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        if (lane == 0) red[warp] = q;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kWarpsPerCta; ++w) t += red[w];
            // fp32 like the reference: (m2 - 1/k) / (k - 1) * k
            const float m2 = (float)t;
            out[n] = (m2 - 1.0f / (float)k) / (float)(k - 1) * (float)k;
        }
        __syncthreads();
    }
}

// max over a row of S with the diagonal element lowered by 2 (redundancy_score); one warp per row
__global__ void __launch_bounds__(kWarpsPerCta * 32)
rowmax_offdiag_kernel(const float* __restrict__ S, int64_t rows, int64_t cols, int64_t row0, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (r >= rows) return;
    const float* sr = S + r * cols;
    const int64_t dg = row0 + r;
    float mx = -INFINITY;
    bool nan = false;
    for (int64_t j = lane; j < cols; j += 32) {
        float v = sr[j];
        if (j == dg) v -= 2.0f;
        nan |= (v != v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nan = __any_sync(0xffffffffu, nan);
    if (lane == 0) out[r] = nan ? __int_as_float(0x7FC00000) : mx;  // torch.max propagates NaN
}

template <int NCH>
int launch_normalize(const float* x, int64_t rows, int D, int Kpad, float eps, int fmt, float plane_scale, uint16_t* hi,
                     uint16_t* lo, float* inv, cudaStream_t st) {
    const unsigned grid = (unsigned)slb_ceil_div(rows, kWarpsPerCta);
    normalize_split_kernel<NCH><<<grid, kWarpsPerCta * 32, 0, st>>>(x, rows, D, Kpad, eps, fmt, plane_scale, hi, lo, inv);
    SLB_LAUNCH_OK("normalize_split");
    return SLB_OK;
}

// mean of n floats in a fixed order (one CTA, float64 accumulation): the last step of redundancy_score
__global__ void __launch_bounds__(256) mean_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    __shared__ double part[256];
    double acc = 0.0;
    bool nan = false;
    for (int64_t i = threadIdx.x; i < n; i += 256) {
        const float v = x[i];
        nan |= v != v;
        acc += (double)v;
    }
    part[threadIdx.x] = nan ? __longlong_as_double(0x7FF8000000000000ll) : acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)(part[0] / (double)n);
}

__global__ void fill_kernel(float* __restrict__ x, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = v;
}

template <int NCH>
int launch_clarity(const float* V, int64_t C, int k, int D, float eps, float* out, cudaStream_t st) {
    const unsigned grid = (unsigned)std::min<int64_t>(C, (int64_t)slb_sm_count() * 8);
    const size_t smem = sizeof(float4) * kWarpsPerCta * NCH * 32;
    if (smem > 48 * 1024)
        SLB_CUDA_OK(cudaFuncSetAttribute(clarity_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    clarity_kernel<NCH><<<grid, kWarpsPerCta * 32, smem, st>>>(V, C, k, D, eps, out);
    SLB_LAUNCH_OK("clarity");
    return SLB_OK;
}

int64_t pad64(int64_t d) { return (d + 63) / 64 * 64; }

}  // namespace

// planes as two explicit pointers: the lo plane need not follow the hi plane directly (padded operands)
static int normalize_split_rows_2p(const float* x, int64_t rows, int64_t D, float eps, int plane_fmt, float plane_scale,
                                   uint16_t* hi, uint16_t* lo, float* inv_norms, void* stream) {
    SLB_REQUIRE(plane_scale > 0.0f, SLB_EINVAL, "slb_normalize_split_rows: plane_scale must be positive");
    SLB_REQUIRE(rows >= 0 && D > 0, SLB_EINVAL, "slb_normalize_split_rows: bad size");
    if (rows == 0) return SLB_OK;
    SLB_REQUIRE(x && (hi || inv_norms), SLB_EINVAL, "slb_normalize_split_rows: null pointer");
    SLB_REQUIRE(plane_fmt == SLB_PLANE_F16 || plane_fmt == SLB_PLANE_BF16, SLB_EINVAL, "slb_normalize_split_rows: bad format");
    SLB_REQUIRE(D % 4 == 0 && D <= 2048 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)hi % 8) == 0 && ((uintptr_t)lo % 8) == 0,
                SLB_EUNSUPPORTED, "slb_normalize_split_rows: D must be a multiple of 4 and <= 2048 (got %lld), x 16-byte aligned",
                (long long)D);
    const int Kpad = (int)pad64(D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SlbProfScope prof("K6 normalize_split_rows", stream, 0.0, (double)rows * (4.0 * (double)D + (hi ? 4.0 * Kpad : 0.0)));
    const int nch = (int)slb_ceil_div(Kpad, 128);
    if (nch <= 2) return launch_normalize<2>(x, rows, (int)D, Kpad, eps, plane_fmt, plane_scale, hi, lo, inv_norms, st);
    if (nch <= 4) return launch_normalize<4>(x, rows, (int)D, Kpad, eps, plane_fmt, plane_scale, hi, lo, inv_norms, st);
    if (nch <= 8) return launch_normalize<8>(x, rows, (int)D, Kpad, eps, plane_fmt, plane_scale, hi, lo, inv_norms, st);
    return launch_normalize<16>(x, rows, (int)D, Kpad, eps, plane_fmt, plane_scale, hi, lo, inv_norms, st);
}

extern "C" int slb_normalize_split_rows(const float* x, int64_t rows, int64_t D, float eps, int plane_fmt, float plane_scale,
                                        uint16_t* planes, float* inv_norms, void* stream) {
    return normalize_split_rows_2p(x, rows, D, eps, plane_fmt, plane_scale, planes,
                                   planes ? planes + rows * pad64(D) : nullptr, inv_norms, stream);
}

// K9 redundancy_score (scores.py:51-81): mean_i max_{j != i} cos(x_i, x_j). The rows are normalised into split planes
// once; ONE tensor-core GEMM of the planes against themselves keeps only the row maxima (diagonal lowered by 2) in its
// epilogue — the n x n cosine matrix (17 GB at n = 65 536) is never written — and a fixed-order mean closes.
int slb_gemm_rowmax_offdiag(const uint16_t* planes, int64_t n, int64_t n_pad, int64_t K, float alpha, float* rowmax, void* stream);

extern "C" size_t slb_redundancy_workspace_bytes(int64_t n, int64_t D) {
    if (n <= 0 || D <= 0) return 0;
    const size_t n_pad = (size_t)((n + 7) / 8 * 8);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return up((size_t)2 * n_pad * (size_t)pad64(D) * 2) + up((size_t)n * 4);
}

extern "C" int slb_redundancy(const float* cones, int64_t n, int64_t D, float* out, void* workspace, size_t workspace_bytes,
                              void* stream) {
    SLB_REQUIRE(n > 0 && D > 0, SLB_EINVAL, "slb_redundancy: bad size");
    SLB_REQUIRE(cones && out && workspace, SLB_EINVAL, "slb_redundancy: null pointer");
    SLB_REQUIRE(((uintptr_t)workspace % 256) == 0, SLB_EINVAL, "slb_redundancy: workspace must be 256-byte aligned");
    const size_t need = slb_redundancy_workspace_bytes(n, D);
    SLB_REQUIRE(workspace_bytes >= need, SLB_EWORKSPACE, "slb_redundancy: workspace needs %zu bytes, got %zu", need, workspace_bytes);
    const int64_t n_pad = (n + 7) / 8 * 8, Kpad = pad64(D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint16_t* hi = static_cast<uint16_t*>(workspace);
    uint16_t* lo = hi + n_pad * Kpad;
    float* rowmax = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + (((size_t)2 * n_pad * Kpad * 2 + 255) & ~(size_t)255));
    if (n_pad != n) {  // zero rows: cos = 0 against everything, and excluded from the maxima by n_valid
        SLB_CUDA_OK(cudaMemsetAsync(hi + n * Kpad, 0, (size_t)(n_pad - n) * Kpad * 2, st));
        SLB_CUDA_OK(cudaMemsetAsync(lo + n * Kpad, 0, (size_t)(n_pad - n) * Kpad * 2, st));
    }
    const float sc = 1024.0f;
    int rc = normalize_split_rows_2p(cones, n, D, 1e-12f, SLB_PLANE_F16, sc, hi, lo, nullptr, stream);
    if (rc != SLB_OK) return rc;
    fill_kernel<<<(unsigned)slb_ceil_div(n, 256), 256, 0, st>>>(rowmax, n, -INFINITY);
    SLB_LAUNCH_OK("fill");
    rc = slb_gemm_rowmax_offdiag(hi, n, n_pad, Kpad, 1.0f / (sc * sc), rowmax, stream);
    if (rc != SLB_OK) return rc;
    mean_kernel<<<1, 256, 0, st>>>(rowmax, n, out);
    SLB_LAUNCH_OK("mean");
    return SLB_OK;
}

extern "C" size_t slb_cosine_gemm_workspace_bytes(int64_t M, int64_t N, int64_t D) {
    if (M < 0 || N < 0 || D <= 0) return 0;
    const size_t Kpad = (size_t)pad64(D);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    return up((size_t)2 * M * Kpad * 2) + up((size_t)2 * N * Kpad * 2);
}

extern "C" int slb_cosine_gemm(const float* x, int64_t M, const float* y, int64_t N, int64_t D, float* out,
                               void* workspace, size_t workspace_bytes, void* stream) {
    SLB_REQUIRE(M >= 0 && N >= 0 && D > 0, SLB_EINVAL, "slb_cosine_gemm: bad size");
    if (M == 0 || N == 0) return SLB_OK;
    SLB_REQUIRE(x && y && out && workspace, SLB_EINVAL, "slb_cosine_gemm: null pointer");
    SLB_REQUIRE(N % 8 == 0, SLB_EUNSUPPORTED, "slb_cosine_gemm: N must be a multiple of 8 (got %lld); pad y with zero rows",
                (long long)N);
    SLB_REQUIRE(((uintptr_t)workspace % 256) == 0, SLB_EINVAL, "slb_cosine_gemm: workspace must be 256-byte aligned");
    const size_t need = slb_cosine_gemm_workspace_bytes(M, N, D);
    SLB_REQUIRE(workspace_bytes >= need, SLB_EWORKSPACE, "slb_cosine_gemm: workspace needs %zu bytes, got %zu", need,
                workspace_bytes);
    const int64_t Kpad = pad64(D);
    uint16_t* px = static_cast<uint16_t*>(workspace);
    uint16_t* py = reinterpret_cast<uint16_t*>(static_cast<unsigned char*>(workspace) +
                                               (((size_t)2 * M * Kpad * 2 + 255) & ~(size_t)255));
    // F.normalize default eps = 1e-12 (scores.py:120-121)
    // unit rows: elements ~ 1/sqrt(D), scaled by 2^10 into the fp16 normal range on both sides
    const float sc = 1024.0f;
    int rc = slb_normalize_split_rows(x, M, D, 1e-12f, SLB_PLANE_F16, sc, px, nullptr, stream);
    if (rc != SLB_OK) return rc;
    rc = slb_normalize_split_rows(y, N, D, 1e-12f, SLB_PLANE_F16, sc, py, nullptr, stream);
    if (rc != SLB_OK) return rc;
    return slb_gemm_split(px, py, SLB_PLANE_F16, M, N, Kpad, 1.0f / (sc * sc), nullptr, nullptr, nullptr, nullptr,
                          SLB_EPI_NONE, 3, out, nullptr, stream);
}

extern "C" int slb_cosine_rows(const float* x, const float* y, int64_t rows, int64_t D, float eps, float* out,
                               void* stream) {
    SLB_REQUIRE(rows >= 0 && D > 0 && D < (1ll << 31), SLB_EINVAL, "slb_cosine_rows: bad size");
    if (rows == 0) return SLB_OK;
    SLB_REQUIRE(x && y && out, SLB_EINVAL, "slb_cosine_rows: null pointer");
    cosine_rows_kernel<<<(unsigned)slb_ceil_div(rows, kWarpsPerCta), kWarpsPerCta * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        x, y, rows, (int)D, eps, out);
    SLB_LAUNCH_OK("cosine_rows");
    return SLB_OK;
}

extern "C" int slb_clarity(const float* V, int64_t C, int64_t k, int64_t D, float* out, void* stream) {
    SLB_REQUIRE(C >= 0 && k > 0 && D > 0, SLB_EINVAL, "slb_clarity: bad size");
    if (C == 0) return SLB_OK;
    SLB_REQUIRE(V && out, SLB_EINVAL, "slb_clarity: null pointer");
    SLB_REQUIRE(D % 4 == 0 && D <= 2048 && k < (1ll << 31) && ((uintptr_t)V % 16) == 0, SLB_EUNSUPPORTED,
                "slb_clarity: D must be a multiple of 4 and <= 2048 (got %lld), V 16-byte aligned", (long long)D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SlbProfScope prof("K7 clarity", stream, 0.0, 4.0 * (double)C * (double)k * (double)D);
    const int nch = (int)slb_ceil_div(D, 128);
    const float eps = 1e-12f;  // F.normalize default (scores.py:45)
    if (nch <= 2) return launch_clarity<2>(V, C, (int)k, (int)D, eps, out, st);
    if (nch <= 4) return launch_clarity<4>(V, C, (int)k, (int)D, eps, out, st);
    if (nch <= 6) return launch_clarity<6>(V, C, (int)k, (int)D, eps, out, st);
    if (nch <= 8) return launch_clarity<8>(V, C, (int)k, (int)D, eps, out, st);
    return launch_clarity<16>(V, C, (int)k, (int)D, eps, out, st);
}

extern "C" int slb_rowmax_offdiag(const float* S, int64_t rows, int64_t cols, int64_t row0, float* out, void* stream) {
    SLB_REQUIRE(rows >= 0 && cols > 0 && row0 >= 0, SLB_EINVAL, "slb_rowmax_offdiag: bad size");
    if (rows == 0) return SLB_OK;
    SLB_REQUIRE(S && out, SLB_EINVAL, "slb_rowmax_offdiag: null pointer");
    rowmax_offdiag_kernel<<<(unsigned)slb_ceil_div(rows, kWarpsPerCta), kWarpsPerCta * 32, 0, static_cast<cudaStream_t>(stream)>>>(
        S, rows, cols, row0, out);
    SLB_LAUNCH_OK("rowmax_offdiag");
    return SLB_OK;
}
