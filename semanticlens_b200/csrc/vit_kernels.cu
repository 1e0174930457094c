// K3 + the non-GEMM parts of the ViT image tower (preprocess, patchify, token assembly, LayerNorm, attention).
//
// Replaces, on the reference's embed path (foundation_models/clip.py:103-163 -> open_clip VisionTransformer):
//   * `ToTensor` + `Normalize` of the eval transform, applied per PIL image on the host (clip.py:157-160)
//   * the im2col view of conv1 (kernel = stride = patch), class-token concat + positional embedding
//   * nn.LayerNorm (ln_pre, ln_1, ln_2, ln_post) and nn.MultiheadAttention's softmax(QK^T)V
// Everything here is HBM/L2-bound fp32 SIMT work; the dense contractions live in gemm_tc.cu. LayerNorm and attention
// write their result directly as split planes, the A-operand format of the next GEMM.
#include "tc_common.cuh"

#include <stdlib.h>
#include <algorithm>

int slb_attention_mma_dh64(const float* q, int64_t q_bs, int64_t q_rs, const float* k, const float* v, int64_t kv_bs,
                           int64_t kv_rs, int64_t B, int64_t Tq, int64_t Tk, int64_t H, float scale, int plane_fmt,
                           float* out_f32, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st);

namespace {

// ------------------------------------------------------------------------------------------------
// K3: u8 planar image -> normalised fp32   out = (u8/255 - mean[c]) / std[c]   (torch's op order)
// ------------------------------------------------------------------------------------------------
// Only 256 x Cc different results exist, so the host computes them (the same two IEEE divisions torch performs) and the
// kernel is a table look-up: the two divisions per pixel were half of the old kernel's time (58 us per 256 images = 0.52 of
// the HBM peak). The table is replicated 16 times in shared memory — entry (c, v) of lane l sits in bank 16 (v & 1) + (l & 15) —
// so a warp's 32 data-dependent look-ups conflict at most two-way (one shared copy: ~3.5-way on random bytes, 47 us; 32
// copies leave room for only two CTAs per SM, 53 us). Persistent CTAs (four per SM) fill their Cc x 16 KB copy once; thread = 4
// pixels per step: a warp reads 128 contiguous bytes and writes 512.
struct U8Lut {
    float v[3 * 256];
};
constexpr int kU8Threads = 512;

__global__ void __launch_bounds__(kU8Threads) u8_norm_kernel(const uint8_t* __restrict__ img, int64_t planes, int words, int slices, int Cc,
                                                             const __grid_constant__ U8Lut lut_in, float* __restrict__ out) {
    extern __shared__ float lut[];  // [Cc][256][16]
    for (int t = threadIdx.x; t < Cc * 256 * 16; t += kU8Threads) lut[t] = lut_in.v[t >> 4];
    __syncthreads();
    const int lane = threadIdx.x & 15;
    const int per = (words + slices - 1) / slices;  // 4-pixel words of a plane handled per job
    for (int64_t job = blockIdx.x; job < planes * slices; job += gridDim.x) {
        const int64_t plane = job / slices;
        const int sl = (int)(job - plane * slices);
        const float* tab = lut + (int)(plane % Cc) * (256 * 16) + lane;
        const uint32_t* src = reinterpret_cast<const uint32_t*>(img) + plane * words;
        float4* dst = reinterpret_cast<float4*>(out) + plane * words;
        const int end = min(words, (sl + 1) * per);
#pragma unroll 4
        for (int i = sl * per + threadIdx.x; i < end; i += kU8Threads) {
            const uint32_t w = __ldg(src + i);
            dst[i] = make_float4(tab[(w & 0xFF) << 4], tab[((w >> 8) & 0xFF) << 4], tab[((w >> 16) & 0xFF) << 4], tab[(w >> 24) << 4]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// patchify: (B, 3, S, S) fp32 -> planes [2][B*g*g][Kpad], row = (b, gy, gx), col = (c, py, px); cols >= 3*P*P are 0
// (Kpad = 3*P*P rounded up to 64: patch 14 gives 588 -> 640). One thread = 2 adjacent columns (P is even).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ img, int64_t B, int S, int P, int Kpad,
                                                       int fmt, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int g = S / P;
    const int Kc = 3 * P * P;
    const int K2 = Kpad / 2;
    const int64_t n2 = B * g * g * (int64_t)K2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % K2) * 2;
        const int64_t row = i / K2;
        uint32_t h2 = 0, l2 = 0;
        if (col < Kc) {
            const int c = col / (P * P), py = (col / P) % P, px = col % P;
            const int gx = (int)(row % g), gy = (int)((row / g) % g);
            const int64_t b = row / (g * g);
            const float2 v = *reinterpret_cast<const float2*>(img + ((b * 3 + c) * S + (gy * P + py)) * (int64_t)S + gx * P + px);
            uint16_t h0, l0, h1, l1;
            slb_split2_act(v.x, fmt, h0, l0);
            slb_split2_act(v.y, fmt, h1, l1);
            h2 = (uint32_t)h0 | ((uint32_t)h1 << 16);
            l2 = (uint32_t)l0 | ((uint32_t)l1 << 16);
        }
        reinterpret_cast<uint32_t*>(hi)[i] = h2;
        reinterpret_cast<uint32_t*>(lo)[i] = l2;
    }
}

// The same for patch sizes that are multiples of 8 (32, 16, 8): one thread = 8 adjacent columns = 8 pixels of one image row
// (two 16-byte loads, one 16-byte store per plane) instead of 2 — the 2-column version is LSU-bound at 1.6 TB/s.
__global__ void __launch_bounds__(256) patchify8_kernel(const float* __restrict__ img, int64_t B, int S, int P, int Kpad,
                                                        int fmt, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int g = S / P;
    const int Kc = 3 * P * P;
    const int K8 = Kpad / 8;
    const int64_t n8 = B * g * g * (int64_t)K8;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % K8) * 8;
        const int64_t row = i / K8;
        uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
        if (col < Kc) {  // Kc is a multiple of 8 here, so the 8 columns are valid together
            const int c = col / (P * P), py = (col / P) % P, px = col % P;
            const int gx = (int)(row % g), gy = (int)((row / g) % g);
            const int64_t b = row / (g * g);
            const float4* src = reinterpret_cast<const float4*>(img + ((b * 3 + c) * S + (gy * P + py)) * (int64_t)S + gx * P + px);
            const float4 v0 = __ldg(src), v1 = __ldg(src + 1);
            const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint16_t hh, ll;
                slb_split2_act(v[j], fmt, hh, ll);
                h[j >> 1] |= (uint32_t)hh << ((j & 1) * 16);
                l[j >> 1] |= (uint32_t)ll << ((j & 1) * 16);
            }
        }
        reinterpret_cast<uint4*>(hi)[i] = make_uint4(h[0], h[1], h[2], h[3]);
        reinterpret_cast<uint4*>(lo)[i] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// x[b, t, :] = (t < has_cls ? cls : patch[b, t - has_cls, :]) + pos[t, :]
__global__ void __launch_bounds__(256) assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                                       const float* __restrict__ pos, int64_t B, int T, int W4, int has_cls,
                                                       float* __restrict__ out) {
    const int64_t n = B * T * (int64_t)W4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W4);
        const int t = (int)((i / W4) % T);
        const int64_t b = i / ((int64_t)W4 * T);
        float4 a = (has_cls && t == 0) ? reinterpret_cast<const float4*>(cls)[w]
                                       : reinterpret_cast<const float4*>(patch)[(b * (T - has_cls) + (t - has_cls)) * W4 + w];
        const float4 p = pos ? reinterpret_cast<const float4*>(pos)[(int64_t)t * W4 + w] : make_float4(0, 0, 0, 0);
        a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        reinterpret_cast<float4*>(out)[i] = a;
    }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, two-pass fp32 (mean, then centred variance), biased variance, eps inside the sqrt
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t rows, int cols,
                                                        int64_t row_stride, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int fmt,
                                                        float* __restrict__ out_f32, uint16_t* __restrict__ out_hi,
                                                        uint16_t* __restrict__ out_lo) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + r * row_stride);
    const int c4 = cols >> 2;
    float s = 0.f;
    for (int i = lane; i < c4; i += 32) {
        float4 v = xr[i];
        s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = slb_warp_sum_butterfly(s) / (float)cols;
    float q = 0.f;
    for (int i = lane; i < c4; i += 32) {
        float4 v = xr[i];
        float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float var = slb_warp_sum_butterfly(q) / (float)cols;
    const float rstd = 1.0f / sqrtf(var + eps);
    for (int i = lane; i < c4; i += 32) {
        float4 v = xr[i];
        const float4 g = reinterpret_cast<const float4*>(gamma)[i];
        const float4 bt = beta ? reinterpret_cast<const float4*>(beta)[i] : make_float4(0, 0, 0, 0);
        float y[4] = {(v.x - mean) * rstd * g.x + bt.x, (v.y - mean) * rstd * g.y + bt.y, (v.z - mean) * rstd * g.z + bt.z,
                      (v.w - mean) * rstd * g.w + bt.w};
        if (out_f32) reinterpret_cast<float4*>(out_f32 + r * cols)[i] = make_float4(y[0], y[1], y[2], y[3]);
        if (out_hi) {
            uint16_t h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) slb_split2_act(y[k], fmt, h[k], l[k]);
            reinterpret_cast<uint2*>(out_hi + r * cols)[i] =
                make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
            reinterpret_cast<uint2*>(out_lo + r * cols)[i] =
                make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
        }
    }
}

// Register-resident variant for cols <= 128 * NCH: the row is read from global memory ONCE (the generic kernel above
// re-reads it for the variance and the normalisation). Same accumulation order as the generic kernel, so the results
// are bit-identical.
template <int NCH>
__global__ void __launch_bounds__(256) layernorm_reg_kernel(const float* __restrict__ x, int64_t rows, int cols,
                                                            int64_t row_stride, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float eps, int fmt,
                                                            float* __restrict__ out_f32, uint16_t* __restrict__ out_hi,
                                                            uint16_t* __restrict__ out_lo) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + r * row_stride);
    const int c4 = cols >> 2;
    float4 v[NCH];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int i = c * 32 + lane;
        if (i < c4) {
            v[c] = xr[i];
            s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
        }
    }
    const float mean = slb_warp_sum_butterfly(s) / (float)cols;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int i = c * 32 + lane;
        if (i < c4) {
            const float a = v[c].x - mean, b = v[c].y - mean, cc = v[c].z - mean, d = v[c].w - mean;
            q += (a * a + b * b) + (cc * cc + d * d);
        }
    }
    const float var = slb_warp_sum_butterfly(q) / (float)cols;
    const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int i = c * 32 + lane;
        if (i < c4) {
            const float4 g = reinterpret_cast<const float4*>(gamma)[i];
            const float4 bt = beta ? reinterpret_cast<const float4*>(beta)[i] : make_float4(0, 0, 0, 0);
            float y[4] = {(v[c].x - mean) * rstd * g.x + bt.x, (v[c].y - mean) * rstd * g.y + bt.y,
                          (v[c].z - mean) * rstd * g.z + bt.z, (v[c].w - mean) * rstd * g.w + bt.w};
            if (out_f32) reinterpret_cast<float4*>(out_f32 + r * cols)[i] = make_float4(y[0], y[1], y[2], y[3]);
            if (out_hi) {
                uint16_t h[4], l[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) slb_split2_act(y[k], fmt, h[k], l[k]);
                reinterpret_cast<uint2*>(out_hi + r * cols)[i] =
                    make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
                reinterpret_cast<uint2*>(out_lo + r * cols)[i] =
                    make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// attention for short sequences: one CTA per (batch, head); K and V of the head live in shared memory (fp32),
// one warp per query row: scores across lanes, softmax by warp shuffles, PV with lanes across head_dim.
// ------------------------------------------------------------------------------------------------
constexpr int kAttnWarps = 8;
constexpr int kAttnMaxChunks = 10;  // Tk <= 320

struct AttnParams {
    const float* q; int64_t q_bs, q_rs;
    const float* k; const float* v; int64_t kv_bs, kv_rs;
    int64_t B; int Tq, Tk, H, dh;
    float scale;
    float* out_f32; uint16_t* out_hi; uint16_t* out_lo; int fmt;
};

__global__ void __launch_bounds__(kAttnWarps * 32) attention_small_kernel(AttnParams p) {
    extern __shared__ __align__(16) float sm[];
    const int dh = p.dh, Tk = p.Tk, ldk = dh + 1;
    float* sK = sm;                         // [Tk][dh+1]
    float* sV = sK + (((size_t)Tk * ldk + 3) & ~(size_t)3);  // [Tk][dh], 16-byte aligned
    float* sQ = sV + (size_t)Tk * dh;       // [warps][dh]
    float* sP = sQ + kAttnWarps * dh;       // [warps][Tk]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x / p.H;
    const int h = (int)(blockIdx.x % p.H);

    const float* kb = p.k + b * p.kv_bs + (int64_t)h * dh;
    const float* vb = p.v + b * p.kv_bs + (int64_t)h * dh;
    const int dh4 = dh >> 2;
    for (int i = threadIdx.x; i < Tk * dh4; i += blockDim.x) {
        const int j = i / dh4, d4 = i % dh4;
        const float4 kk = *reinterpret_cast<const float4*>(kb + (int64_t)j * p.kv_rs + 4 * d4);
        const float4 vv = *reinterpret_cast<const float4*>(vb + (int64_t)j * p.kv_rs + 4 * d4);
        float* kd = sK + (size_t)j * ldk + 4 * d4;
        kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
        *reinterpret_cast<float4*>(sV + (size_t)j * dh + 4 * d4) = vv;
    }
    __syncthreads();

    float* myQ = sQ + warp * dh;
    float* myP = sP + (size_t)warp * Tk;
    const int nch = (Tk + 31) >> 5;
    const int Wd = p.H * dh;
    for (int qi = warp; qi < p.Tq; qi += kAttnWarps) {
        const float* qr = p.q + b * p.q_bs + (int64_t)qi * p.q_rs + (int64_t)h * dh;
        for (int d = lane; d < dh; d += 32) myQ[d] = qr[d] * p.scale;
        __syncwarp();
        float s[kAttnMaxChunks];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kAttnMaxChunks; ++c) {
            s[c] = -INFINITY;
            if (c < nch) {
                const int j = c * 32 + lane;
                if (j < Tk) {
                    const float* kr = sK + (size_t)j * ldk;
                    float a0 = 0.f, a1 = 0.f;
                    for (int d = 0; d < dh; d += 2) {
                        a0 = fmaf(myQ[d], kr[d], a0);
                        a1 = fmaf(myQ[d + 1], kr[d + 1], a1);
                    }
                    s[c] = a0 + a1;
                }
                mx = fmaxf(mx, s[c]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int c = 0; c < kAttnMaxChunks; ++c) {
            if (c < nch) {
                const int j = c * 32 + lane;
                const float e = (j < Tk) ? expf(s[c] - mx) : 0.f;
                s[c] = e;
                sum += e;
            }
        }
        sum = slb_warp_sum_butterfly(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int c = 0; c < kAttnMaxChunks; ++c) {
            if (c < nch) {
                const int j = c * 32 + lane;
                if (j < Tk) myP[j] = s[c] * inv;
            }
        }
        __syncwarp();
        const int64_t orow = b * p.Tq + qi;
        for (int d = lane; d < dh; d += 32) {
            float o0 = 0.f, o1 = 0.f;
            int j = 0;
            for (; j + 1 < Tk; j += 2) {
                o0 = fmaf(myP[j], sV[(size_t)j * dh + d], o0);
                o1 = fmaf(myP[j + 1], sV[(size_t)(j + 1) * dh + d], o1);
            }
            if (j < Tk) o0 = fmaf(myP[j], sV[(size_t)j * dh + d], o0);
            const float o = o0 + o1;
            const int64_t idx = orow * Wd + (int64_t)h * dh + d;
            if (p.out_f32) p.out_f32[idx] = o;
            if (p.out_hi) {
                uint16_t hh, ll;
                slb_split2_act(o, p.fmt, hh, ll);
                p.out_hi[idx] = hh;
                p.out_lo[idx] = ll;
            }
        }
        __syncwarp();
    }
}

int grid_for(int64_t n, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(slb_ceil_div(n, threads), (int64_t)slb_sm_count() * 16));
}

}  // namespace

extern "C" int slb_u8_to_f32_norm(const uint8_t* img, int64_t B, int64_t Cc, int64_t n_pix, const float* mean3,
                                  const float* std3, float* out, void* stream) {
    SLB_REQUIRE(B >= 0 && n_pix >= 0, SLB_EINVAL, "slb_u8_to_f32_norm: negative size");
    SLB_REQUIRE(Cc >= 1 && Cc <= 3, SLB_EUNSUPPORTED, "slb_u8_to_f32_norm: 1..3 channels supported (got %lld)", (long long)Cc);
    if (B == 0 || n_pix == 0) return SLB_OK;
    SLB_REQUIRE(img && out && mean3 && std3, SLB_EINVAL, "slb_u8_to_f32_norm: null pointer (mean3/std3 are HOST arrays)");
    SLB_REQUIRE(n_pix % 16 == 0 && ((uintptr_t)img % 16) == 0 && ((uintptr_t)out % 16) == 0, SLB_EUNSUPPORTED,
                "slb_u8_to_f32_norm: planes must be multiples of 16 pixels and 16-byte aligned");
    SLB_REQUIRE(n_pix / 4 < (1ll << 31), SLB_EUNSUPPORTED, "slb_u8_to_f32_norm: plane too large");
    SlbProfScope prof("K3 u8_to_f32_norm", stream, 0.0, 5.0 * (double)B * (double)Cc * (double)n_pix);
    U8Lut lut;
    for (int c = 0; c < (int)Cc; ++c) {
        // volatile: keeps the compiler from folding the two divisions into anything but two IEEE divisions
        volatile float m = mean3[c], sd = std3[c];
        for (int v = 0; v < 256; ++v) {
            volatile float t = (float)v / 255.0f;
            volatile float u = t - m;
            lut.v[c * 256 + v] = u / sd;
        }
    }
    for (int i = (int)Cc * 256; i < 3 * 256; ++i) lut.v[i] = 0.f;
    const int words = (int)(n_pix / 4);
    // jobs of ~2048 words (4 per thread) over four persistent CTAs per SM
    const int slices = (int)std::max<int64_t>(1, slb_ceil_div(words, 2048));
    const int64_t jobs = B * Cc * slices;
    const int grid = (int)std::min<int64_t>(jobs, (int64_t)slb_sm_count() * 4);
    const size_t smem = (size_t)Cc * 256 * 16 * sizeof(float);
    SLB_CUDA_OK(cudaFuncSetAttribute(u8_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 256 * 16 * (int)sizeof(float)));
    u8_norm_kernel<<<grid, kU8Threads, smem, static_cast<cudaStream_t>(stream)>>>(img, B * Cc, words, slices, (int)Cc, lut, out);
    SLB_LAUNCH_OK("u8_norm");
    return SLB_OK;
}

extern "C" int slb_patchify(const float* img, int64_t B, int64_t S, int64_t P, int plane_fmt, uint16_t* out_planes,
                            void* stream) {
    SLB_REQUIRE(B >= 0 && S > 0 && P > 0, SLB_EINVAL, "slb_patchify: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(img && out_planes, SLB_EINVAL, "slb_patchify: null pointer");
    SLB_REQUIRE(S % P == 0 && P % 2 == 0, SLB_EUNSUPPORTED, "slb_patchify: need S %% P == 0 and an even patch size");
    SLB_REQUIRE(((uintptr_t)img % 8) == 0 && ((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_patchify: misaligned");
    const int64_t g = S / P, Kpad = slb_patch_k(P), n = B * g * g * Kpad;
    SlbProfScope prof("K4 patchify", stream, 0.0, 12.0 * (double)B * (double)S * (double)S + 4.0 * (double)n);
    if (P % 8 == 0 && S % 4 == 0 && ((uintptr_t)img % 16) == 0)
        patchify8_kernel<<<grid_for(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, B, (int)S, (int)P, (int)Kpad,
                                                                                             plane_fmt, out_planes, out_planes + n);
    else
        patchify_kernel<<<grid_for(n / 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, B, (int)S, (int)P, (int)Kpad,
                                                                                            plane_fmt, out_planes, out_planes + n);
    SLB_LAUNCH_OK("patchify");
    return SLB_OK;
}

extern "C" int64_t slb_patch_k(int64_t P) { return (3 * P * P + 63) / 64 * 64; }

extern "C" int slb_assemble_tokens(const float* patch, const float* cls, const float* pos, int64_t B, int64_t T, int64_t W,
                                   int has_cls, float* out, void* stream) {
    SLB_REQUIRE(B >= 0 && T > 0 && W > 0, SLB_EINVAL, "slb_assemble_tokens: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(patch && out && (!has_cls || cls), SLB_EINVAL, "slb_assemble_tokens: null pointer");
    SLB_REQUIRE(W % 4 == 0, SLB_EUNSUPPORTED, "slb_assemble_tokens: width must be a multiple of 4");
    SlbProfScope prof("K4 assemble_tokens", stream, 0.0, 8.0 * (double)B * (double)T * (double)W);
    assemble_kernel<<<grid_for(B * T * W / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        patch, cls, pos, B, (int)T, (int)(W / 4), has_cls ? 1 : 0, out);
    SLB_LAUNCH_OK("assemble_tokens");
    return SLB_OK;
}

extern "C" int slb_layernorm(const float* x, int64_t rows, int64_t cols, int64_t row_stride, const float* gamma,
                             const float* beta, float eps, int plane_fmt, float* out_f32, uint16_t* out_planes,
                             void* stream) {
    SLB_REQUIRE(rows >= 0 && cols > 0, SLB_EINVAL, "slb_layernorm: bad size");
    if (rows == 0) return SLB_OK;
    SLB_REQUIRE(x && gamma && (out_f32 || out_planes), SLB_EINVAL, "slb_layernorm: null pointer");
    SLB_REQUIRE(cols % 4 == 0 && row_stride % 4 == 0 && row_stride >= cols, SLB_EUNSUPPORTED,
                "slb_layernorm: cols and row_stride must be multiples of 4");
    const int threads = 256;
    SlbProfScope prof("K4 layernorm", stream, 0.0, (double)rows * (double)cols * (4.0 + (out_f32 ? 4.0 : 0.0) + (out_planes ? 4.0 : 0.0)));
    const unsigned grid = (unsigned)slb_ceil_div(rows * 32, threads);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint16_t* lo = out_planes ? out_planes + rows * cols : nullptr;
    // in-place use (ln_pre: out_f32 == x) is safe in both kernels: a warp owns its row
    if (cols <= 768)
        layernorm_reg_kernel<6><<<grid, threads, 0, st>>>(x, rows, (int)cols, row_stride, gamma, beta, eps, plane_fmt, out_f32,
                                                         out_planes, lo);
    else if (cols <= 1024)
        layernorm_reg_kernel<8><<<grid, threads, 0, st>>>(x, rows, (int)cols, row_stride, gamma, beta, eps, plane_fmt, out_f32,
                                                         out_planes, lo);
    else
        layernorm_kernel<<<grid, threads, 0, st>>>(x, rows, (int)cols, row_stride, gamma, beta, eps, plane_fmt, out_f32,
                                                   out_planes, lo);
    SLB_LAUNCH_OK("layernorm");
    return SLB_OK;
}

extern "C" int slb_attention_small(const float* q, int64_t q_batch_stride, int64_t q_row_stride, const float* k,
                                   const float* v, int64_t kv_batch_stride, int64_t kv_row_stride, int64_t B, int64_t Tq,
                                   int64_t Tk, int64_t H, int64_t dh, float scale, int plane_fmt, float* out_f32,
                                   uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && Tq > 0 && Tk > 0 && H > 0 && dh > 0, SLB_EINVAL, "slb_attention_small: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(q && k && v && (out_f32 || out_planes), SLB_EINVAL, "slb_attention_small: null pointer");
    SLB_REQUIRE(dh % 4 == 0 && dh <= 128 && kv_row_stride % 4 == 0 && kv_batch_stride % 4 == 0 &&
                    ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0,
                SLB_EUNSUPPORTED, "slb_attention_small: head_dim must be a multiple of 4 (<= 128), K/V 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // QK^T and PV, three plane products each on the tensor-core path
    SlbProfScope prof("K4 attention", stream, 4.0 * (double)B * (double)H * (double)Tq * (double)Tk * (double)dh * 3.0,
                      4.0 * (double)B * (double)H * (double)dh * ((double)Tq * 2.0 + (double)Tk * 2.0));
    // head_dim 64 (every CLIP / SigLIP tower): tensor-core path (attention_mma.cu); anything else: the SIMT kernel below
    static const bool force_simt = [] { const char* e = getenv("SLB_ATTN_SIMT"); return e && e[0] == '1'; }();
    if (dh == 64 && !force_simt && q_row_stride % 2 == 0 && q_batch_stride % 2 == 0 && ((uintptr_t)q % 8) == 0 &&
        (Tk + 15) / 16 * 16 <= 400) {
        uint16_t* lo = out_planes ? out_planes + B * Tq * H * dh : nullptr;
        return slb_attention_mma_dh64(q, q_batch_stride, q_row_stride, k, v, kv_batch_stride, kv_row_stride, B, Tq, Tk, H,
                                      scale, plane_fmt, out_f32, out_planes, lo, st);
    }
    SLB_REQUIRE(Tk <= 32 * kAttnMaxChunks, SLB_EUNSUPPORTED, "slb_attention_small: at most %d keys (got %lld)",
                32 * kAttnMaxChunks, (long long)Tk);
    AttnParams p{};
    p.q = q; p.q_bs = q_batch_stride; p.q_rs = q_row_stride;
    p.k = k; p.v = v; p.kv_bs = kv_batch_stride; p.kv_rs = kv_row_stride;
    p.B = B; p.Tq = (int)Tq; p.Tk = (int)Tk; p.H = (int)H; p.dh = (int)dh;
    p.scale = scale;
    p.out_f32 = out_f32; p.out_hi = out_planes; p.out_lo = out_planes ? out_planes + B * Tq * H * dh : nullptr;
    p.fmt = plane_fmt;
    const size_t smem = sizeof(float) * ((((size_t)Tk * (dh + 1) + 3) & ~(size_t)3) + (size_t)Tk * dh + (size_t)kAttnWarps * dh + (size_t)kAttnWarps * Tk);
    SLB_REQUIRE(smem <= 227 * 1024, SLB_EUNSUPPORTED, "slb_attention_small: K/V of one head do not fit shared memory");
    SLB_CUDA_OK(cudaFuncSetAttribute(attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SLB_REQUIRE(B * H <= 0x7FFFFFFF, SLB_EUNSUPPORTED, "slb_attention_small: grid too large");
    attention_small_kernel<<<(unsigned)(B * H), kAttnWarps * 32, smem, st>>>(p);
    SLB_LAUNCH_OK("attention_small");
    return SLB_OK;
}
