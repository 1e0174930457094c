// The CLIP ModifiedResNet image tower ("RN50" / "RN101") as ONE C-ABI call: (B,3,S,S) fp32 -> (B, out_dim) features.
//
// Replaces `self.model.encode_image(img)` of the reference (foundation_models/clip.py:103-118) for open_clip's
// ModifiedResNet (the embed model of BASELINE.json configs[0]): 3-convolution stem + AvgPool2d(2),
// bottlenecks whose stride is an average pool (anti-aliasing), AttentionPool2d head.
//
// Data layout: activations are channels-last split planes [2, B*H*W, C] at SLB_ACT_PLANE_SCALE. That makes every 1x1
// convolution (two thirds of the convolutions, 60 % of the flops) exactly slb_gemm_split over the activation planes —
// TMA-fed tcgen05 tiles, no data movement — and a 3x3 convolution of the bottlenecks an IMPLICIT GEMM (slb_conv_gemm: the
// producer warp fetches (filter tap, 64-channel chunk) k-blocks with TMA im2col-mode loads; the 9x im2col matrix of
// round 1 — 97 MB per image — is gone). Only the stem's 32-channel 3x3 convolutions still go through im2col planes.
// Eval-mode BatchNorm, ReLU and the shortcut add are GEMM epilogues (col_scale / bias / SLB_EPI_RELU / SLB_EPI_ADD_RELU);
// the block output is written both as planes (operand of the next block) and fp32 (its shortcut).
// No allocation: caller-supplied workspace (slb_rn_workspace_bytes). Nothing synchronises.
#include "tc_common.cuh"

#include <stdlib.h>
#include <algorithm>

namespace {

constexpr float kAlpha = 1.0f / (SLB_ACT_PLANE_SCALE * SLB_WEIGHT_PLANE_SCALE);
// 50+ convolutions in sequence with no normalisation between them: the truncation bias of the tensor-core accumulate adds
// up coherently, so the cross terms get their own accumulator (slb200.h, SLB_PASSES_SPLIT_ACC). SLB_RN_FAST=1 trades that
// for the faster single-accumulator tiles (measured 5e-5 instead of ~2e-5 of the largest feature on RN50).
int rn_passes() {
    static const int v = [] {
        const char* e = getenv("SLB_RN_FAST");
        return (e && e[0] == '1') ? 3 : SLB_PASSES_SPLIT_ACC;
    }();
    return v;
}

bool rn_f32_stream() {
    static const bool v = [] {
        const char* e = getenv("SLB_RN_F32_STREAM");
        return e && atoi(e) != 0;
    }();
    return v;
}

bool rn_explicit_im2col() {
    static const bool v = [] {
        const char* e = getenv("SLB_RN_IM2COL");
        return e && e[0] == '1';
    }();
    return v;
}

int grid_for(int64_t n, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(slb_ceil_div(n, threads), (int64_t)slb_sm_count() * 16));
}

// ---- stem im2col: one thread = one 16-byte chunk (8 columns) of one output row, both planes --------------------------
__global__ void __launch_bounds__(256) im2col_stem_kernel(const float* __restrict__ img, int64_t n_chunks, int S, int fmt,
                                                          uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int Ho = S >> 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int kc = (int)(i & 7);
        const int64_t m = i >> 3;
        const int x = (int)(m % Ho);
        const int y = (int)((m / Ho) % Ho);
        const int64_t b = m / ((int64_t)Ho * Ho);
        uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
        if (kc < 4) {
            const float* base = img + b * 3 * (int64_t)S * S;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = kc * 8 + j;
                float v = 0.f;
                if (k < 27) {
                    const int tap = k / 3, c = k - tap * 3;
                    const int yy = 2 * y + tap / 3 - 1, xx = 2 * x + tap % 3 - 1;
                    if (yy >= 0 && yy < S && xx >= 0 && xx < S) v = __ldg(base + ((int64_t)c * S + yy) * S + xx);
                }
                uint16_t hh, ll;
                slb_split2_act(v, fmt, hh, ll);
                h[j >> 1] |= (uint32_t)hh << ((j & 1) * 16);
                l[j >> 1] |= (uint32_t)ll << ((j & 1) * 16);
            }
        }
        *reinterpret_cast<uint4*>(hi + m * 64 + kc * 8) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + m * 64 + kc * 8) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ---- the stem's first convolution computed directly (fp32 FMAs): 3 -> 8 G channels, 3x3, stride 2, pad 1 ---------------------
// 27 products per output are not worth a tensor-core pass: as a GEMM this layer cost an im2col kernel (0.29 ms per 128
// images: 411 MB of zero-padded K = 64 rows) plus a GEMM that is pure epilogue (0.33 ms), against ~0.1 ms of HBM time for
// reading the images and writing the planes once. Thread = (output pixel, 8 adjacent output channels); the G threads of a
// pixel share its 27 inputs (same addresses: one transaction), a warp stores 32/G pixels x 16 G bytes = 512 contiguous bytes
// per plane. Weights are the 22-bit values the planes hold (hi + lo, descaled), accumulation is fp32 in tap order (ky, kx, c):
// batch-invariant and at least as accurate as the split GEMM. BatchNorm + ReLU + plane split as in the GEMM epilogue.
template <int G>
__global__ void __launch_bounds__(256) stem_conv_kernel(const float* __restrict__ img, int64_t n_work, int S, int fmt,
                                                        const uint16_t* __restrict__ w_hi, const uint16_t* __restrict__ w_lo,
                                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                                        uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    constexpr int CO = 8 * G;
    __shared__ float4 wsm[27][CO / 4];
    __shared__ float sc[CO], sh[CO];
    for (int t = threadIdx.x; t < 27 * CO; t += 256) {
        const int k = t / CO, n = t - k * CO;
        reinterpret_cast<float*>(wsm)[t] =
            (slb_from_plane(w_hi[n * 64 + k], fmt) + slb_from_plane(w_lo[n * 64 + k], fmt)) * (1.0f / SLB_WEIGHT_PLANE_SCALE);
    }
    if (threadIdx.x < CO) {
        sc[threadIdx.x] = scale[threadIdx.x];
        sh[threadIdx.x] = shift[threadIdx.x];
    }
    __syncthreads();
    // thread = (4 horizontally adjacent output pixels, 8 adjacent output channels): a weight fetched from shared memory feeds
    // four FMAs (one pixel per thread made the kernel shared-memory bound: 54 16-byte reads per 216 FMAs), and the 9 input
    // columns of a row serve all four pixels
    const int Ho = S >> 1, Q = (Ho + 3) >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_work; i += (int64_t)gridDim.x * blockDim.x) {
        const int g = (int)(i % G);
        const int64_t q = i / G;
        const int xq = (int)(q % Q);
        const int y = (int)((q / Q) % Ho);
        const int64_t b = q / ((int64_t)Q * Ho);
        const int x0 = xq * 4;
        const float* base = img + b * 3 * (int64_t)S * S;
        float acc[4][8];
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = 2 * y + ky - 1;
            const bool row_in = yy >= 0 && yy < S;
            float v[3][9];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float* row = base + ((int64_t)c * S + (row_in ? yy : 0)) * S;
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int xx = 2 * x0 + t - 1;
                    v[c][t] = (row_in && xx >= 0 && xx < S) ? __ldg(row + xx) : 0.f;
                }
            }
            // same accumulation order per output as the one-pixel form: (ky, kx, c)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int k = (ky * 3 + kx) * 3 + c;
                    const float4 w0 = wsm[k][g * 2], w1 = wsm[k][g * 2 + 1];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float a = v[c][2 * p + kx];
                        acc[p][0] = fmaf(a, w0.x, acc[p][0]); acc[p][1] = fmaf(a, w0.y, acc[p][1]);
                        acc[p][2] = fmaf(a, w0.z, acc[p][2]); acc[p][3] = fmaf(a, w0.w, acc[p][3]);
                        acc[p][4] = fmaf(a, w1.x, acc[p][4]); acc[p][5] = fmaf(a, w1.y, acc[p][5]);
                        acc[p][6] = fmaf(a, w1.z, acc[p][6]); acc[p][7] = fmaf(a, w1.w, acc[p][7]);
                    }
                }
        }
        const int64_t m0 = (b * Ho + y) * (int64_t)Ho + x0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            if (x0 + p >= Ho) break;
            uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float z = fmaxf(fmaf(acc[p][j], sc[g * 8 + j], sh[g * 8 + j]), 0.0f);
                uint16_t hh, ll;
                slb_split2_act(z, fmt, hh, ll);
                h[j >> 1] |= (uint32_t)hh << ((j & 1) * 16);
                l[j >> 1] |= (uint32_t)ll << ((j & 1) * 16);
            }
            *reinterpret_cast<uint4*>(hi + (m0 + p) * CO + g * 8) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(lo + (m0 + p) * CO + g * 8) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

// ---- 3x3 / stride 1 / pad 1 im2col over channels-last planes: pure 16-byte moves ------------------------------------
// grid.y = plane. One warp owns an output pixel (row of the im2col matrix) at a time: the pixel is decoded once, the
// lanes sweep the row's 16-byte chunks (tap-major, channels contiguous), so stores are fully coalesced and loads are
// contiguous within a tap.
template <bool kPow2>
__global__ void __launch_bounds__(256) im2col3x3_kernel(const uint16_t* __restrict__ in, int M, int H, int W, int C8, int c8_shift,
                                                        int K8, uint16_t* __restrict__ out) {
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)blockIdx.y * M * C8;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)blockIdx.y * M * K8;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
        const int x = m % W;
        const int y = (m / W) % H;
        const uint4* row_src = src + (int64_t)m * C8;
        uint4* row_dst = dst + (int64_t)m * K8;
        for (int kc = lane; kc < K8; kc += 32) {
            const int tap = kPow2 ? (kc >> c8_shift) : (kc / C8);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (tap < 9) {
                const int c8 = kc - tap * C8;
                const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                if ((unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W)
                    v = __ldg(row_src + (dy * W + dx) * C8 + c8);
            }
            row_dst[kc] = v;
        }
    }
}

// ---- AvgPool2d(2) over channels-last planes ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) avgpool2_kernel(const uint16_t* __restrict__ in, int64_t M_in, int64_t n_chunks, int Ho,
                                                       int Wo, int C8, int fmt, uint16_t* __restrict__ out, int64_t M_out) {
    const uint4* sh = reinterpret_cast<const uint4*>(in);
    const uint4* sl = sh + M_in * C8;
    uint4* dh = reinterpret_cast<uint4*>(out);
    uint4* dl = dh + M_out * C8;
    const int W = 2 * Wo;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const int64_t mo = i / C8;
        const int xo = (int)(mo % Wo);
        const int yo = (int)((mo / Wo) % Ho);
        const int64_t b = mo / ((int64_t)Wo * Ho);
        const int64_t m00 = (b * 2 * Ho + 2 * yo) * W + 2 * xo;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int64_t m = m00 + (t >> 1) * W + (t & 1);
            const uint4 a = __ldg(sh + m * C8 + c8), c = __ldg(sl + m * C8 + c8);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint16_t hb = (uint16_t)(aw[j >> 1] >> ((j & 1) * 16)), lb = (uint16_t)(cw[j >> 1] >> ((j & 1) * 16));
                acc[j] += slb_from_plane(hb, fmt) + slb_from_plane(lb, fmt);
            }
        }
        uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint16_t hh, ll;
            slb_split2(acc[j] * 0.25f, fmt, hh, ll);  // the sum is already at the activation scale
            h[j >> 1] |= (uint32_t)hh << ((j & 1) * 16);
            l[j >> 1] |= (uint32_t)ll << ((j & 1) * 16);
        }
        dh[i] = make_uint4(h[0], h[1], h[2], h[3]);
        dl[i] = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// ---- attention-pool tokens: thread = (image, channel); positions are walked with coalesced loads ------------------------
__global__ void __launch_bounds__(256) pool_tokens_kernel(const float* __restrict__ x, const float* __restrict__ pos, int64_t B,
                                                          int HW, int C, int fmt, uint16_t* __restrict__ tok, uint16_t* __restrict__ qry) {
    const int64_t n = B * C;
    const int T = HW + 1;
    uint16_t* tok_lo = tok + B * T * (int64_t)C;
    uint16_t* qry_lo = qry + n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int64_t b = i / C;
        const float* xb = x + b * HW * (int64_t)C + c;
        float sum = 0.f;
        for (int t = 0; t < HW; ++t) {
            const float v = xb[(int64_t)t * C];
            sum += v;
            uint16_t hh, ll;
            slb_split2_act(v + __ldg(pos + (int64_t)(t + 1) * C + c), fmt, hh, ll);
            const int64_t o = (b * T + t + 1) * C + c;
            tok[o] = hh;
            tok_lo[o] = ll;
        }
        uint16_t hh, ll;
        slb_split2_act(sum / (float)HW + __ldg(pos + c), fmt, hh, ll);
        tok[b * T * C + c] = hh;
        tok_lo[b * T * C + c] = ll;
        qry[i] = hh;
        qry_lo[i] = ll;
    }
}

// ---- workspace ---------------------------------------------------------------------------------------------------------
struct RnLayout {
    size_t xa, xb, t1, t2, t3, xp, f, col, head, total;
};

size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

int n_convs_of(const SlbRnWeights* w) {
    int n = 3;
    for (int i = 0; i < 4; ++i) n += 3 * w->blocks[i] + 1;
    return n;
}

bool rn_layout(const SlbRnWeights* w, int64_t B, RnLayout* L) {
    if (!w || w->width <= 0 || w->width % 64 || w->image_size <= 0 || w->image_size % 32) return false;
    for (int i = 0; i < 4; ++i)
        if (w->blocks[i] < 1) return false;
    const int64_t wd = w->width, S = w->image_size;
    int64_t M = B * (S / 2) * (S / 2);
    // element counts (per plane); byte sizes are 4x for planes (2 planes x 2 bytes) and for fp32
    int64_t x_max = M * wd, t_max = M * (wd / 2), t3_max = 0, xp_max = 0, f_max = 0;
    int64_t col_max = std::max<int64_t>(M * 64, M * slb_conv_k(wd / 2, 3));
    M /= 4;  // after the stem's pool
    int64_t inpl = wd;
    for (int li = 0; li < 4; ++li) {
        const int64_t pl = wd << li;
        for (int bi = 0; bi < w->blocks[li]; ++bi) {
            const int stride = (bi == 0 && li > 0) ? 2 : 1;
            x_max = std::max(x_max, M * inpl);
            t_max = std::max(t_max, M * pl);
            col_max = std::max(col_max, M * slb_conv_k(pl, 3));
            if (stride == 2) {
                xp_max = std::max(xp_max, M / 4 * inpl);
                t3_max = std::max(t3_max, M / 4 * pl);
                M /= 4;
            }
            x_max = std::max(x_max, M * 4 * pl);
            f_max = std::max(f_max, M * 4 * pl);
            inpl = 4 * pl;
        }
    }
    const int64_t E = 32 * wd, T = (S / 32) * (S / 32) + 1;
    col_max = std::max(col_max, B * T * 2 * E);  // the k | v projections (fp32) reuse the im2col region
    x_max = std::max(x_max, B * T * E);          // the attention-pool tokens reuse an activation buffer
    size_t o = 0;
    L->xa = o;   o += align_up((size_t)x_max * 4);
    L->xb = o;   o += align_up((size_t)x_max * 4);
    L->t1 = o;   o += align_up((size_t)t_max * 4);
    L->t2 = o;   o += align_up((size_t)t_max * 4);
    L->t3 = o;   o += align_up((size_t)t3_max * 4);
    L->xp = o;   o += align_up((size_t)xp_max * 4);
    L->f = o;    o += align_up((size_t)f_max * 4);
    L->col = o;  o += align_up((size_t)col_max * 4);
    L->head = o; o += 3 * align_up((size_t)B * E * 4);  // query planes | projected query fp32 | pooled planes
    L->total = o;
    return true;
}

}  // namespace

extern "C" int64_t slb_conv_k(int64_t cin, int64_t ksize) { return (cin * ksize * ksize + 63) / 64 * 64; }

extern "C" int slb_im2col_stem(const float* img, int64_t B, int64_t S, int plane_fmt, uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && S > 0, SLB_EINVAL, "slb_im2col_stem: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(img && out_planes, SLB_EINVAL, "slb_im2col_stem: null pointer");
    SLB_REQUIRE(S % 2 == 0, SLB_EUNSUPPORTED, "slb_im2col_stem: the image size must be even");
    SLB_REQUIRE(((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_im2col_stem: misaligned output");
    const int64_t M = B * (S / 2) * (S / 2);
    SlbProfScope prof("K4 im2col", stream, 0.0, 12.0 * (double)B * (double)S * (double)S + 4.0 * 64.0 * (double)M);
    im2col_stem_kernel<<<grid_for(M * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, M * 8, (int)S, plane_fmt, out_planes,
                                                                                           out_planes + M * 64);
    SLB_LAUNCH_OK("im2col_stem");
    return SLB_OK;
}

extern "C" int slb_stem_conv3x3s2(const float* img, int64_t B, int64_t S, const uint16_t* w_planes, int64_t cout, int plane_fmt,
                                  const float* scale, const float* shift, uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && S > 0, SLB_EINVAL, "slb_stem_conv3x3s2: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(img && w_planes && scale && shift && out_planes, SLB_EINVAL, "slb_stem_conv3x3s2: null pointer");
    SLB_REQUIRE(S % 2 == 0, SLB_EUNSUPPORTED, "slb_stem_conv3x3s2: the image size must be even");
    SLB_REQUIRE(cout == 32 || cout == 64, SLB_EUNSUPPORTED, "slb_stem_conv3x3s2: 32 or 64 output channels (got %lld)", (long long)cout);
    SLB_REQUIRE(plane_fmt == SLB_PLANE_F16 || plane_fmt == SLB_PLANE_BF16, SLB_EINVAL, "slb_stem_conv3x3s2: bad plane format");
    SLB_REQUIRE(((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_stem_conv3x3s2: misaligned output");
    const int64_t M = B * (S / 2) * (S / 2);
    SlbProfScope prof("K4 stem conv (direct)", stream, 2.0 * 27.0 * (double)cout * (double)M,
                      12.0 * (double)B * (double)S * (double)S + 4.0 * (double)cout * (double)M);
    const int64_t n_work = B * (S / 2) * ((S / 2 + 3) / 4) * (cout / 8);  // (pixel quad, 8-channel group) items
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(slb_ceil_div(n_work, 256), (int64_t)slb_sm_count() * 32));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint16_t* w_lo = w_planes + cout * 64;
    uint16_t* lo = out_planes + M * cout;
    if (cout == 32)
        stem_conv_kernel<4><<<grid, 256, 0, st>>>(img, n_work, (int)S, plane_fmt, w_planes, w_lo, scale, shift, out_planes, lo);
    else
        stem_conv_kernel<8><<<grid, 256, 0, st>>>(img, n_work, (int)S, plane_fmt, w_planes, w_lo, scale, shift, out_planes, lo);
    SLB_LAUNCH_OK("stem_conv");
    return SLB_OK;
}

extern "C" int slb_im2col3x3(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, uint16_t* out_planes,
                             void* stream) {
    SLB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, SLB_EINVAL, "slb_im2col3x3: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(in_planes && out_planes, SLB_EINVAL, "slb_im2col3x3: null pointer");
    SLB_REQUIRE(C % 8 == 0, SLB_EUNSUPPORTED, "slb_im2col3x3: channels must be a multiple of 8 (got %lld)", (long long)C);
    SLB_REQUIRE(((uintptr_t)in_planes % 16) == 0 && ((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_im2col3x3: misaligned");
    const int64_t M = B * H * W, K = slb_conv_k(C, 3);
    SlbProfScope prof("K4 im2col", stream, 0.0, 4.0 * (double)M * (double)(K + C));
    SLB_REQUIRE(M < (1ll << 31) / 8, SLB_EUNSUPPORTED, "slb_im2col3x3: too many pixels (%lld)", (long long)M);
    const int C8 = (int)(C / 8);
    const bool pow2 = (C8 & (C8 - 1)) == 0;
    int shift = 0;
    while ((1 << shift) < C8) ++shift;
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(slb_ceil_div(M, 8), (int64_t)slb_sm_count() * 32)), 2);
    if (pow2)
        im2col3x3_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_planes, (int)M, (int)H, (int)W, C8, shift,
                                                                                   (int)(K / 8), out_planes);
    else
        im2col3x3_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_planes, (int)M, (int)H, (int)W, C8, shift,
                                                                                    (int)(K / 8), out_planes);
    SLB_LAUNCH_OK("im2col3x3");
    return SLB_OK;
}

extern "C" int slb_avgpool2_planes(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, int plane_fmt,
                                   uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, SLB_EINVAL, "slb_avgpool2_planes: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(in_planes && out_planes, SLB_EINVAL, "slb_avgpool2_planes: null pointer");
    SLB_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, SLB_EUNSUPPORTED,
                "slb_avgpool2_planes: H and W must be even and C a multiple of 8");
    SLB_REQUIRE(((uintptr_t)in_planes % 16) == 0 && ((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_avgpool2_planes: misaligned");
    const int64_t Mo = B * (H / 2) * (W / 2), n = Mo * (C / 8);
    SlbProfScope prof("K4 avgpool", stream, 0.0, 4.0 * 5.0 * (double)Mo * (double)C);
    avgpool2_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(in_planes, B * H * W, n, (int)(H / 2), (int)(W / 2),
                                                                                    (int)(C / 8), plane_fmt, out_planes, Mo);
    SLB_LAUNCH_OK("avgpool2");
    return SLB_OK;
}

extern "C" int slb_pool_tokens(const float* x, const float* pos, int64_t B, int64_t HW, int64_t C, int plane_fmt,
                               uint16_t* tok_planes, uint16_t* query_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && HW > 0 && C > 0, SLB_EINVAL, "slb_pool_tokens: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(x && pos && tok_planes && query_planes, SLB_EINVAL, "slb_pool_tokens: null pointer");
    SlbProfScope prof("K4 pool_tokens", stream, 0.0, 8.0 * (double)B * (double)(HW + 1) * (double)C);
    pool_tokens_kernel<<<grid_for(B * C, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, pos, B, (int)HW, (int)C, plane_fmt,
                                                                                           tok_planes, query_planes);
    SLB_LAUNCH_OK("pool_tokens");
    return SLB_OK;
}

extern "C" size_t slb_rn_workspace_bytes(const SlbRnWeights* w, int64_t B) {
    RnLayout L;
    if (B < 0 || !rn_layout(w, B, &L)) return 0;
    return L.total;
}

extern "C" int slb_rn_forward(const SlbRnWeights* w, const float* img, int64_t B, float* out, void* workspace,
                              size_t workspace_bytes, void* stream) {
    SLB_REQUIRE(w != nullptr && B >= 0, SLB_EINVAL, "slb_rn_forward: bad arguments");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(img && out && workspace, SLB_EINVAL, "slb_rn_forward: null pointer");
    RnLayout L;
    SLB_REQUIRE(rn_layout(w, B, &L), SLB_EUNSUPPORTED,
                "slb_rn_forward: width must be a multiple of 64, image_size of 32, every stage needs a bottleneck");
    SLB_REQUIRE(w->convs && w->n_convs == n_convs_of(w), SLB_EINVAL, "slb_rn_forward: expected %d convolutions, got %d",
                n_convs_of(w), w->n_convs);
    SLB_REQUIRE(w->pos && w->w_q && w->b_q && w->w_kv && w->b_kv && w->w_c && w->b_c, SLB_EINVAL,
                "slb_rn_forward: incomplete attention pool");
    const int64_t wd = w->width, S = w->image_size, E = 32 * wd;
    SLB_REQUIRE(w->heads > 0 && E % w->heads == 0 && (E / w->heads) % 4 == 0 && E / w->heads <= 128, SLB_EUNSUPPORTED,
                "slb_rn_forward: unsupported head count");
    SLB_REQUIRE(w->out_dim > 0 && w->out_dim % 8 == 0, SLB_EUNSUPPORTED, "slb_rn_forward: out_dim must be a multiple of 8");
    SLB_REQUIRE(((uintptr_t)workspace % 256) == 0, SLB_EINVAL, "slb_rn_forward: workspace must be 256-byte aligned");
    SLB_REQUIRE(workspace_bytes >= L.total, SLB_EWORKSPACE, "slb_rn_forward: workspace needs %zu bytes, got %zu", L.total,
                workspace_bytes);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    uint16_t* x_cur = reinterpret_cast<uint16_t*>(ws + L.xa);
    uint16_t* x_nxt = reinterpret_cast<uint16_t*>(ws + L.xb);
    uint16_t* t1 = reinterpret_cast<uint16_t*>(ws + L.t1);
    uint16_t* t2 = reinterpret_cast<uint16_t*>(ws + L.t2);
    uint16_t* t3 = reinterpret_cast<uint16_t*>(ws + L.t3);
    uint16_t* xp = reinterpret_cast<uint16_t*>(ws + L.xp);
    uint16_t* sc = reinterpret_cast<uint16_t*>(ws + L.f);  // a projected shortcut as planes (same 4 bytes per element as fp32)
    uint16_t* col = reinterpret_cast<uint16_t*>(ws + L.col);
    float* f = reinterpret_cast<float*>(ws + L.col);  // the last block's fp32 map: read by the pool before k | v land in `col`
    const int fmt = w->plane_fmt;
    const int kPasses = rn_passes();
    int rc;
#define SLB_TRY(call)                \
    do {                             \
        rc = (call);                 \
        if (rc != SLB_OK) return rc; \
    } while (0)
    // conv (as a GEMM over planes `a` with M rows) + BatchNorm + epilogue
    auto conv = [&](const SlbConvBn& c, const uint16_t* a, int64_t M, int epi, const float* residual, float* o32, uint16_t* opl) {
        return slb_gemm_split(a, c.w, fmt, M, c.cout, slb_conv_k(c.cin, c.ksize), kAlpha, c.shift, residual, nullptr, c.scale, epi,
                              kPasses, o32, opl, stream);
    };
    for (int i = 0; i < w->n_convs; ++i)
        SLB_REQUIRE(w->convs[i].w && w->convs[i].scale && w->convs[i].shift, SLB_EINVAL, "slb_rn_forward: convolution %d is incomplete", i);
    const SlbConvBn* cv = w->convs;
    SLB_REQUIRE(cv[0].cin == 3 && cv[0].ksize == 3 && cv[0].cout == wd / 2 && cv[2].cout == wd, SLB_EINVAL,
                "slb_rn_forward: the stem does not match width %lld", (long long)wd);

    // ---- stem ----
    int64_t H = S / 2, M = B * H * H;
    if ((cv[0].cout == 32 || cv[0].cout == 64) && !rn_explicit_im2col()) {
        SLB_TRY(slb_stem_conv3x3s2(img, B, S, cv[0].w, cv[0].cout, fmt, cv[0].scale, cv[0].shift, t1, stream));
    } else {
        SLB_TRY(slb_im2col_stem(img, B, S, fmt, col, stream));
        SLB_TRY(conv(cv[0], col, M, SLB_EPI_RELU, nullptr, nullptr, t1));
    }
    // the two 3x3 convolutions over wd/2 channels: implicit GEMMs too (32 channels = one tap per k-block); at 112 x 112 their
    // im2col matrices were 2 GB each per 128 images — 15 % of the RN50 tower's time went into writing and re-reading them
    const int64_t hw_ = wd / 2;
    if ((hw_ % 64 == 0 || hw_ == 32) && !rn_explicit_im2col()) {
        SLB_TRY(slb_conv_gemm(t1, B, H, H, hw_, 3, 1, 1, cv[1].w, cv[1].cout, fmt, kAlpha, cv[1].shift, nullptr, cv[1].scale, SLB_EPI_RELU,
                              kPasses, nullptr, t2, stream));
        SLB_TRY(slb_conv_gemm(t2, B, H, H, hw_, 3, 1, 1, cv[2].w, cv[2].cout, fmt, kAlpha, cv[2].shift, nullptr, cv[2].scale, SLB_EPI_RELU,
                              kPasses, nullptr, x_nxt, stream));
    } else {
        SLB_TRY(slb_im2col3x3(t1, B, H, H, hw_, col, stream));
        SLB_TRY(conv(cv[1], col, M, SLB_EPI_RELU, nullptr, nullptr, t2));
        SLB_TRY(slb_im2col3x3(t2, B, H, H, hw_, col, stream));
        SLB_TRY(conv(cv[2], col, M, SLB_EPI_RELU, nullptr, nullptr, x_nxt));
    }
    SLB_TRY(slb_avgpool2_planes(x_nxt, B, H, H, wd, fmt, x_cur, stream));
    H /= 2;
    M /= 4;
    cv += 3;

    // ---- bottlenecks ----
    int64_t inpl = wd;
    for (int li = 0; li < 4; ++li) {
        const int64_t pl = wd << li;
        for (int bi = 0; bi < w->blocks[li]; ++bi) {
            const bool ds = bi == 0;
            const int stride = (bi == 0 && li > 0) ? 2 : 1;
            SLB_REQUIRE(cv[0].cin == inpl && cv[0].cout == pl && cv[0].ksize == 1 && cv[1].cin == pl && cv[1].cout == pl &&
                            cv[1].ksize == 3 && cv[2].cin == pl && cv[2].cout == 4 * pl && cv[2].ksize == 1 &&
                            (!ds || (cv[3].cin == inpl && cv[3].cout == 4 * pl && cv[3].ksize == 1)),
                        SLB_EINVAL, "slb_rn_forward: stage %d bottleneck %d has unexpected convolution shapes", li + 1, bi);
            SLB_TRY(conv(cv[0], x_cur, M, SLB_EPI_RELU, nullptr, nullptr, t1));
            // the 3x3 convolution as an implicit GEMM (TMA im2col-mode loads; no 9x matrix); SLB_RN_IM2COL=1: the explicit path
            if (pl % 64 == 0 && !rn_explicit_im2col()) {
                SLB_TRY(slb_conv_gemm(t1, B, H, H, pl, 3, 1, 1, cv[1].w, cv[1].cout, fmt, kAlpha, cv[1].shift, nullptr, cv[1].scale,
                                      SLB_EPI_RELU, kPasses, nullptr, t2, stream));
            } else {
                SLB_TRY(slb_im2col3x3(t1, B, H, H, pl, col, stream));
                SLB_TRY(conv(cv[1], col, M, SLB_EPI_RELU, nullptr, nullptr, t2));
            }
            const uint16_t* main_in = t2;
            const uint16_t* short_in = x_cur;
            if (stride == 2) {
                SLB_TRY(slb_avgpool2_planes(t2, B, H, H, pl, fmt, t3, stream));
                SLB_TRY(slb_avgpool2_planes(x_cur, B, H, H, inpl, fmt, xp, stream));
                main_in = t3;
                short_in = xp;
                H /= 2;
                M /= 4;
            }
            // The residual stream lives as split planes (22 bits per value, 2^-22 relative per block): the shortcut is the
            // block's input planes or its projection written as planes, and the tail reads it through
            // SLB_EPI_ADD_RELU_PLANES. Same bytes to read as an fp32 shortcut, but no block writes an fp32 copy of its
            // output (a third of the tail convolution's traffic) — except the last one, whose fp32 map feeds the pool.
            const bool last = li == 3 && bi == w->blocks[3] - 1;
            if (rn_f32_stream()) {  // SLB_RN_F32_STREAM=1: the round-1 arrangement (fp32 shortcut buffer), kept for A/B runs
                float* f0 = reinterpret_cast<float*>(sc);
                if (ds) SLB_TRY(conv(cv[3], short_in, M, SLB_EPI_NONE, nullptr, f0, nullptr));
                SLB_TRY(conv(cv[2], main_in, M, SLB_EPI_ADD_RELU, f0, f0, x_nxt));
                if (last) SLB_CUDA_OK(cudaMemcpyAsync(f, f0, (size_t)M * 4 * pl * 4, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
                std::swap(x_cur, x_nxt);
                cv += ds ? 4 : 3;
                inpl = 4 * pl;
                continue;
            }
            const uint16_t* shortcut = x_cur;
            if (ds) {
                SLB_TRY(conv(cv[3], short_in, M, SLB_EPI_NONE, nullptr, nullptr, sc));
                shortcut = sc;
            }
            SLB_TRY(conv(cv[2], main_in, M, SLB_EPI_ADD_RELU_PLANES, reinterpret_cast<const float*>(shortcut), last ? f : nullptr,
                         last ? nullptr : x_nxt));
            std::swap(x_cur, x_nxt);
            cv += ds ? 4 : 3;
            inpl = 4 * pl;
        }
    }

    // ---- attention pool ----
    const int64_t HW = H * H, T = HW + 1, dh = E / w->heads;
    unsigned char* head = ws + L.head;
    const size_t hs = align_up((size_t)B * E * 4);
    uint16_t* q_planes = reinterpret_cast<uint16_t*>(head);
    float* q32 = reinterpret_cast<float*>(head + hs);
    uint16_t* pooled = reinterpret_cast<uint16_t*>(head + 2 * hs);
    float* kv = reinterpret_cast<float*>(col);
    uint16_t* tok = x_nxt;
    SLB_TRY(slb_pool_tokens(f, w->pos, B, HW, E, fmt, tok, q_planes, stream));
    SLB_TRY(slb_gemm_split(tok, w->w_kv, fmt, B * T, 2 * E, E, kAlpha, w->b_kv, nullptr, nullptr, nullptr, SLB_EPI_NONE, kPasses, kv, nullptr,
                           stream));
    SLB_TRY(slb_gemm_split(q_planes, w->w_q, fmt, B, E, E, kAlpha, w->b_q, nullptr, nullptr, nullptr, SLB_EPI_NONE, kPasses, q32, nullptr,
                           stream));
    SLB_TRY(slb_attention_small(q32, E, E, kv, kv + E, T * 2 * E, 2 * E, B, 1, T, w->heads, dh, 1.0f / sqrtf((float)dh), fmt, nullptr,
                                pooled, stream));
    SLB_TRY(slb_gemm_split(pooled, w->w_c, fmt, B, w->out_dim, E, kAlpha, w->b_c, nullptr, nullptr, nullptr, SLB_EPI_NONE, kPasses, out,
                           nullptr, stream));
#undef SLB_TRY
    return SLB_OK;
}
