// Building blocks of the accelerated probed-model forward (semanticlens_b200/probed.py): torchvision-style ResNets
// (Conv2d / BatchNorm2d / ReLU / MaxPool2d stacks — the probed models of BASELINE.json configs[0..3], run by the
// reference as `self.model(x)` inside ActivationComponentVisualizer._run, activation_based.py:341-358) on the same
// tcgen05 GEMM as the CLIP ModifiedResNet tower. Opt-in: the default sweep keeps the user's model in PyTorch.
//
// Layout as in rn_tower.cu: activations are channels-last split planes [2, B*H*W, C] at SLB_ACT_PLANE_SCALE, a
// convolution is slb_gemm_split over (im2col) planes with eval-mode BatchNorm / ReLU / shortcut as its epilogue. What
// this file adds is what torchvision's ResNet needs beyond the CLIP tower: a k x k / stride s stem straight from NCHW
// fp32 images (7x7 / 2), strided 3x3 im2col and pixel subsampling (stride lives in the convolution, not in an average
// pool), BatchNorm + ReLU + MaxPool2d(3, 2, 1) in one pass, and an affine + activation pass for convolutions whose RAW
// output a forward hook wants to see (the GEMM then writes the un-normalised fp32 map and this pass makes the planes).
#include "tc_common.cuh"

#include <algorithm>

namespace {

int grid_for(int64_t n, int threads) {
    return (int)std::max<int64_t>(1, std::min<int64_t>(slb_ceil_div(n, threads), (int64_t)slb_sm_count() * 16));
}

__device__ __forceinline__ void pack8(const float (&v)[8], int fmt, uint4& hi, uint4& lo) {
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    if (fmt == SLB_PLANE_F16) {
#pragma unroll
        for (int j = 0; j < 4; ++j) slb_split_pair_act_f16(v[2 * j], v[2 * j + 1], h[j], l[j]);
        hi = make_uint4(h[0], h[1], h[2], h[3]);
        lo = make_uint4(l[0], l[1], l[2], l[3]);
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint16_t hh, ll;
        slb_split2_act(v[j], fmt, hh, ll);
        h[j >> 1] |= (uint32_t)hh << ((j & 1) * 16);
        l[j >> 1] |= (uint32_t)ll << ((j & 1) * 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// ---- k x k / stride / pad im2col straight from NCHW fp32 images: thread = one 16-byte chunk (8 columns) of one output
// row, both planes; column (ky * k + kx) * C + c, zero outside the image and past C k k --------------------------------
__global__ void __launch_bounds__(256) im2col_nchw_kernel(const float* __restrict__ img, int64_t n_chunks, int C, int H, int W, int Ho,
                                                          int Wo, int k, int stride, int pad, int K8, int fmt,
                                                          uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    const int Ktrue = C * k * k;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int kc = (int)(i % K8);
        const int64_t m = i / K8;
        const int x = (int)(m % Wo);
        const int y = (int)((m / Wo) % Ho);
        const int64_t b = m / ((int64_t)Wo * Ho);
        const float* base = img + b * C * (int64_t)H * W;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = kc * 8 + j;
            v[j] = 0.f;
            if (col < Ktrue) {
                const int tap = col / C, c = col - tap * C;
                const int ky = tap / k, kx = tap - ky * k;
                const int yy = y * stride + ky - pad, xx = x * stride + kx - pad;
                if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) v[j] = __ldg(base + ((int64_t)c * H + yy) * W + xx);
            }
        }
        uint4 h4, l4;
        pack8(v, fmt, h4, l4);
        *reinterpret_cast<uint4*>(hi + m * (int64_t)K8 * 8 + kc * 8) = h4;
        *reinterpret_cast<uint4*>(lo + m * (int64_t)K8 * 8 + kc * 8) = l4;
    }
}

// The same im2col for the common case of a whole output row per CTA: the k input rows a row of outputs needs are staged in
// shared memory with coalesced loads (zero padding included), then every thread assembles 16-byte chunks from there. The
// direct kernel above gathers 4 bytes at a time from up to 32 cache lines per warp instruction and ran at 1 TB/s on the
// 7x7 / 2 stem of a ResNet (2.5 GB of planes per 256 images); this one is bound by the plane stores.
__global__ void __launch_bounds__(256) im2col_nchw_rows_kernel(const float* __restrict__ img, int C, int H, int W, int Ho, int Wo, int k,
                                                               int stride, int pad, int K8, int Wn, int fmt,
                                                               uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
    extern __shared__ float rows[];  // [C][k][Wn]: input column xx - pad of input row y * stride + ky - pad
    int* tab = reinterpret_cast<int*>(rows + C * k * Wn);  // [8 K8]: column -> offset of its (channel, ky) row + kx, -1 past C k k
    for (int col = threadIdx.x; col < K8 * 8; col += blockDim.x) {
        const int tap = col / C, c = col - tap * C;
        const int ky = tap / k, kx = tap - ky * k;
        tab[col] = col < C * k * k ? (c * k + ky) * Wn + kx : -1;
    }
    const int y = blockIdx.x % Ho;
    const int64_t b = blockIdx.x / Ho;
    const float* base = img + b * C * (int64_t)H * W;
    for (int i = threadIdx.x; i < C * k * Wn; i += blockDim.x) {
        const int xx = i % Wn, r = i / Wn;
        const int ky = r % k, c = r / k;
        const int yy = y * stride + ky - pad, xi = xx - pad;
        rows[i] = ((unsigned)yy < (unsigned)H && (unsigned)xi < (unsigned)W) ? __ldg(base + ((int64_t)c * H + yy) * W + xi) : 0.f;
    }
    __syncthreads();
    const int64_t m0 = (b * Ho + y) * (int64_t)Wo;
    for (int q = threadIdx.x; q < Wo * K8; q += blockDim.x) {
        const int x = q / K8, kc = q - x * K8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int o = tab[kc * 8 + j];
            v[j] = o >= 0 ? rows[o + x * stride] : 0.f;
        }
        uint4 h4, l4;
        pack8(v, fmt, h4, l4);
        *reinterpret_cast<uint4*>(hi + (m0 + x) * (int64_t)K8 * 8 + kc * 8) = h4;
        *reinterpret_cast<uint4*>(lo + (m0 + x) * (int64_t)K8 * 8 + kc * 8) = l4;
    }
}

// ---- 3x3 / pad 1 / stride s im2col over channels-last planes: a warp per output pixel, lanes sweep the row's 16-byte
// chunks (tap-major, channels contiguous). grid.y = plane --------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col3x3_strided_kernel(const uint16_t* __restrict__ in, int64_t M_in, int M_out, int H, int W,
                                                                int Ho, int Wo, int stride, int C8, int K8, uint16_t* __restrict__ out) {
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)blockIdx.y * M_in * C8;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)blockIdx.y * M_out * K8;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M_out; m += warps) {
        const int xo = m % Wo;
        const int yo = (m / Wo) % Ho;
        const int b = m / (Wo * Ho);
        const int y = yo * stride, x = xo * stride;
        const uint4* row_src = src + ((int64_t)(b * H + y) * W + x) * C8;
        uint4* row_dst = dst + (int64_t)m * K8;
        for (int kc = lane; kc < K8; kc += 32) {
            const int tap = kc / C8;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (tap < 9) {
                const int c8 = kc - tap * C8;
                const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                if ((unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W) v = __ldg(row_src + (dy * W + dx) * C8 + c8);
            }
            row_dst[kc] = v;
        }
    }
}

// ---- every second pixel of every second row (the input of a 1x1 / stride 2 convolution); grid.y = plane ----------------
__global__ void __launch_bounds__(256) subsample2_kernel(const uint16_t* __restrict__ in, int64_t M_in, int64_t M_out, int H, int W,
                                                         int Ho, int Wo, int C8, uint16_t* __restrict__ out) {
    const uint4* src = reinterpret_cast<const uint4*>(in) + (int64_t)blockIdx.y * M_in * C8;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)blockIdx.y * M_out * C8;
    const int64_t n = M_out * C8;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const int64_t mo = i / C8;
        const int xo = (int)(mo % Wo);
        const int yo = (int)((mo / Wo) % Ho);
        const int64_t b = mo / ((int64_t)Wo * Ho);
        dst[i] = __ldg(src + ((b * H + 2 * yo) * W + 2 * xo) * C8 + c8);
    }
}

// ---- y = act(raw * scale[c] + shift[c] (+ residual)) over a channels-last fp32 map: thread = 8 channels of a pixel ------
__global__ void __launch_bounds__(256) affine_act_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, const float* residual, int64_t n8, int C8,
                                                         int relu, int fmt, float* out_f32, uint16_t* __restrict__ hi,
                                                         uint16_t* __restrict__ lo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(raw) + 2 * i), a1 = __ldg(reinterpret_cast<const float4*>(raw) + 2 * i + 1);
        float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], __ldg(scale + c + j), __ldg(shift + c + j));
        if (residual) {  // may alias out_f32: every element is read before it is written, by the same thread
            const float4 r0 = reinterpret_cast<const float4*>(residual)[2 * i], r1 = reinterpret_cast<const float4*>(residual)[2 * i + 1];
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
            v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = v[j] < 0.f ? 0.f : v[j];  // NaN stays NaN, like torch.relu
        }
        if (out_f32) {
            reinterpret_cast<float4*>(out_f32)[2 * i] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(out_f32)[2 * i + 1] = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (hi) {
            uint4 h4, l4;
            pack8(v, fmt, h4, l4);
            reinterpret_cast<uint4*>(hi)[i] = h4;
            reinterpret_cast<uint4*>(lo)[i] = l4;
        }
    }
}

// ---- BatchNorm + ReLU + MaxPool2d(3, stride 2, pad 1) over a channels-last fp32 map: thread = 8 channels of an output
// pixel. The pool's -inf padding never wins after the ReLU, so border taps are simply skipped -----------------------------
__global__ void __launch_bounds__(256) bn_relu_maxpool_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int64_t n8, int H, int W, int Ho, int Wo,
                                                              int C8, int fmt, float* __restrict__ out_f32, uint16_t* __restrict__ hi,
                                                              uint16_t* __restrict__ lo) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const int64_t mo = i / C8;
        const int xo = (int)(mo % Wo);
        const int yo = (int)((mo / Wo) % Ho);
        const int64_t b = mo / ((int64_t)Wo * Ho);
        float sc[8], sh[8], best[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = __ldg(scale + c8 * 8 + j);
            sh[j] = __ldg(shift + c8 * 8 + j);
            best[j] = 0.f;  // the ReLU's floor
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = 2 * yo + t / 3 - 1, xx = 2 * xo + t % 3 - 1;
            if ((unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W) {
                const float4* p = reinterpret_cast<const float4*>(raw) + (((b * H + yy) * W + xx) * C8 + c8) * 2;
                const float4 a0 = __ldg(p), a1 = __ldg(p + 1);
                const float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float z = fmaf(v[j], sc[j], sh[j]);
                    if (z > best[j] || z != z) best[j] = z;  // NaN propagates like torch's max_pool2d
                }
            }
        }
        if (out_f32) {
            reinterpret_cast<float4*>(out_f32)[2 * i] = make_float4(best[0], best[1], best[2], best[3]);
            reinterpret_cast<float4*>(out_f32)[2 * i + 1] = make_float4(best[4], best[5], best[6], best[7]);
        }
        if (hi) {
            uint4 h4, l4;
            pack8(best, fmt, h4, l4);
            reinterpret_cast<uint4*>(hi)[i] = h4;
            reinterpret_cast<uint4*>(lo)[i] = l4;
        }
    }
}

}  // namespace

extern "C" int slb_im2col_nchw(const float* img, int64_t B, int64_t C, int64_t H, int64_t W, int ksize, int stride, int pad,
                               int plane_fmt, uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && ksize > 0 && stride > 0 && pad >= 0, SLB_EINVAL, "slb_im2col_nchw: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(img && out_planes, SLB_EINVAL, "slb_im2col_nchw: null pointer");
    SLB_REQUIRE(((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_im2col_nchw: misaligned output");
    const int64_t Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    SLB_REQUIRE(Ho > 0 && Wo > 0, SLB_EINVAL, "slb_im2col_nchw: empty output");
    const int64_t M = B * Ho * Wo, K = slb_conv_k(C, ksize);
    SlbProfScope prof("conv im2col", stream, 0.0, 4.0 * (double)B * (double)C * (double)H * (double)W + 4.0 * (double)K * (double)M);
    const int64_t Wn = (Wo - 1) * stride + ksize;  // staged columns per input row (padding included)
    const size_t smem = (size_t)C * ksize * Wn * sizeof(float) + (size_t)K * sizeof(int);
    if (smem <= 48 * 1024 && B * Ho < (1ll << 31)) {
        im2col_nchw_rows_kernel<<<(unsigned)(B * Ho), 256, smem, static_cast<cudaStream_t>(stream)>>>(
            img, (int)C, (int)H, (int)W, (int)Ho, (int)Wo, ksize, stride, pad, (int)(K / 8), (int)Wn, plane_fmt, out_planes, out_planes + M * K);
    } else {
        im2col_nchw_kernel<<<grid_for(M * (K / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            img, M * (K / 8), (int)C, (int)H, (int)W, (int)Ho, (int)Wo, ksize, stride, pad, (int)(K / 8), plane_fmt, out_planes,
            out_planes + M * K);
    }
    SLB_LAUNCH_OK("im2col_nchw");
    return SLB_OK;
}

extern "C" int slb_im2col3x3_strided(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, int stride,
                                     uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && (stride == 1 || stride == 2), SLB_EINVAL, "slb_im2col3x3_strided: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(in_planes && out_planes, SLB_EINVAL, "slb_im2col3x3_strided: null pointer");
    SLB_REQUIRE(C % 8 == 0, SLB_EUNSUPPORTED, "slb_im2col3x3_strided: channels must be a multiple of 8 (got %lld)", (long long)C);
    SLB_REQUIRE(((uintptr_t)in_planes % 16) == 0 && ((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_im2col3x3_strided: misaligned");
    const int64_t Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const int64_t M = B * Ho * Wo, K = slb_conv_k(C, 3);
    SLB_REQUIRE(B * H * W < (1ll << 31) / 8, SLB_EUNSUPPORTED, "slb_im2col3x3_strided: too many pixels");
    SlbProfScope prof("conv im2col", stream, 0.0, 4.0 * (double)M * (double)(K + C));
    dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>(slb_ceil_div(M, 8), (int64_t)slb_sm_count() * 32)), 2);
    im2col3x3_strided_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_planes, B * H * W, (int)M, (int)H, (int)W, (int)Ho,
                                                                                 (int)Wo, stride, (int)(C / 8), (int)(K / 8), out_planes);
    SLB_LAUNCH_OK("im2col3x3_strided");
    return SLB_OK;
}

extern "C" int slb_subsample2_planes(const uint16_t* in_planes, int64_t B, int64_t H, int64_t W, int64_t C, uint16_t* out_planes,
                                     void* stream) {
    SLB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, SLB_EINVAL, "slb_subsample2_planes: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(in_planes && out_planes, SLB_EINVAL, "slb_subsample2_planes: null pointer");
    SLB_REQUIRE(C % 8 == 0, SLB_EUNSUPPORTED, "slb_subsample2_planes: channels must be a multiple of 8");
    SLB_REQUIRE(((uintptr_t)in_planes % 16) == 0 && ((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_subsample2_planes: misaligned");
    const int64_t Ho = (H + 1) / 2, Wo = (W + 1) / 2, Mo = B * Ho * Wo;
    SlbProfScope prof("conv subsample", stream, 0.0, 8.0 * (double)Mo * (double)C);
    dim3 grid((unsigned)grid_for(Mo * (C / 8), 256), 2);
    subsample2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_planes, B * H * W, Mo, (int)H, (int)W, (int)Ho, (int)Wo,
                                                                          (int)(C / 8), out_planes);
    SLB_LAUNCH_OK("subsample2");
    return SLB_OK;
}

extern "C" int slb_affine_act(const float* raw, int64_t M, int64_t C, const float* scale, const float* shift, const float* residual,
                              int relu, int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(M >= 0 && C > 0, SLB_EINVAL, "slb_affine_act: bad size");
    if (M == 0) return SLB_OK;
    SLB_REQUIRE(raw && scale && shift && (out_f32 || out_planes), SLB_EINVAL, "slb_affine_act: null pointer");
    SLB_REQUIRE(C % 8 == 0, SLB_EUNSUPPORTED, "slb_affine_act: channels must be a multiple of 8");
    SLB_REQUIRE(((uintptr_t)raw % 16) == 0 && ((uintptr_t)residual % 16) == 0 && ((uintptr_t)out_f32 % 16) == 0 &&
                    ((uintptr_t)out_planes % 16) == 0, SLB_EINVAL, "slb_affine_act: misaligned");
    const int64_t n8 = M * (C / 8);
    SlbProfScope prof("conv affine_act", stream, 0.0, (double)M * (double)C * (4.0 + (residual ? 4.0 : 0.0) + (out_f32 ? 4.0 : 0.0) + (out_planes ? 4.0 : 0.0)));
    affine_act_kernel<<<grid_for(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        raw, scale, shift, residual, n8, (int)(C / 8), relu, plane_fmt, out_f32, out_planes, out_planes ? out_planes + M * C : nullptr);
    SLB_LAUNCH_OK("affine_act");
    return SLB_OK;
}

extern "C" int slb_bn_relu_maxpool(const float* raw, int64_t B, int64_t H, int64_t W, int64_t C, const float* scale, const float* shift,
                                   int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0, SLB_EINVAL, "slb_bn_relu_maxpool: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(raw && scale && shift && (out_f32 || out_planes), SLB_EINVAL, "slb_bn_relu_maxpool: null pointer");
    SLB_REQUIRE(C % 8 == 0, SLB_EUNSUPPORTED, "slb_bn_relu_maxpool: channels must be a multiple of 8");
    SLB_REQUIRE(((uintptr_t)raw % 16) == 0 && ((uintptr_t)out_f32 % 16) == 0 && ((uintptr_t)out_planes % 16) == 0, SLB_EINVAL,
                "slb_bn_relu_maxpool: misaligned");
    const int64_t Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1, Mo = B * Ho * Wo, n8 = Mo * (C / 8);
    SlbProfScope prof("conv bn_relu_maxpool", stream, 0.0, 4.0 * (double)B * (double)H * (double)W * (double)C + (double)Mo * (double)C * (out_f32 ? 8.0 : 4.0));
    bn_relu_maxpool_kernel<<<grid_for(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        raw, scale, shift, n8, (int)H, (int)W, (int)Ho, (int)Wo, (int)(C / 8), plane_fmt, out_f32, out_planes,
        out_planes ? out_planes + Mo * C : nullptr);
    SLB_LAUNCH_OK("bn_relu_maxpool");
    return SLB_OK;
}
