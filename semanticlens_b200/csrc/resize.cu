// Bicubic Resize + CenterCrop of open_clip's eval transform on the GPU, bit-exact with Pillow.
//
// The reference preprocesses every image on the host (foundation_models/clip.py:157-160 -> torchvision Resize(S, BICUBIC)
// + CenterCrop(S) on PIL images -> Pillow's ImagingResample, Resample.c). This is the same algorithm:
// separable Keys bicubic (a = -0.5) whose support widens with the down-scaling factor, coefficients normalised in double
// precision and rounded to 22-bit fixed point, a horizontal pass into an 8-bit intermediate, then a vertical pass, both
// rounding with (acc + 2^21) >> 22 and clamping to [0, 255]. Integer arithmetic end to end, so the result must equal
// Pillow's byte for byte (tests/test_resize_gpu.py); the coefficient kernel uses explicitly rounded double operations
// (__dmul_rn / __dadd_rn: no FMA contraction) in the C source's evaluation order.
// Only the crop window is produced: rows/columns that CenterCrop would discard are never computed.
#include "slb_common.cuh"

#include <algorithm>
#include <cmath>

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

struct Axis {
    int in_size, out_size, crop0, n_out, ksize;
    double scale, support, ss;
};

Axis make_axis(int64_t in_size, int64_t out_size, int64_t crop0, int64_t n_out) {
    Axis a;
    a.in_size = (int)in_size;
    a.out_size = (int)out_size;
    a.crop0 = (int)crop0;
    a.n_out = (int)n_out;
    a.scale = (double)((float)in_size - 0.0f) / (double)out_size;  // the box is held in C floats
    const double filterscale = a.scale < 1.0 ? 1.0 : a.scale;
    a.support = 2.0 * filterscale;
    a.ss = 1.0 / filterscale;
    a.ksize = (int)std::ceil(a.support) * 2 + 1;
    return a;
}

__device__ __forceinline__ double bicubic(double x) {
    x = fabs(x);
    if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
    if (x < 2.0) return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
    return 0.0;
}

// bounds[i] = (first tap, tap count), kk[i][ksize] fixed-point weights for output index crop0 + i
__global__ void resize_coeffs_kernel(Axis a, int* __restrict__ bounds, int* __restrict__ kk) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_out) return;
    const double center = __dmul_rn((double)(a.crop0 + i) + 0.5, a.scale);
    int xmin = __double2int_rz(__dadd_rn(__dsub_rn(center, a.support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, a.support), 0.5));
    if (xmax > a.in_size) xmax = a.in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x)
        ww = __dadd_rn(ww, bicubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), a.ss)));
    int* k = kk + (size_t)i * a.ksize;
    for (int x = 0; x < a.ksize; ++x) {
        int v = 0;
        if (x < xmax) {
            double w = bicubic(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), a.ss));
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double f = __dmul_rn(w, (double)(1 << kPrecisionBits));
            v = __double2int_rz(w < 0.0 ? __dadd_rn(-0.5, f) : __dadd_rn(0.5, f));
        }
        k[x] = v;
    }
    bounds[2 * i] = xmin;
    bounds[2 * i + 1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;
    return (uint8_t)min(max(v, 0), 255);
}

// tmp[r][xo][c] for source rows r0 + r: horizontal pass over the crop's columns
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, int w, int r0, int n_rows, int n_out, int ksize,
                                                       const int* __restrict__ bounds, const int* __restrict__ kk,
                                                       uint8_t* __restrict__ tmp) {
    const int64_t n = (int64_t)n_rows * n_out;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % n_out), r = (int)(i / n_out);
        const int xmin = bounds[2 * xo], cnt = bounds[2 * xo + 1];
        const int* k = kk + (size_t)xo * ksize;
        const uint8_t* p = src + ((int64_t)(r0 + r) * w + xmin) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < cnt; ++x) {
            const int kv = k[x];
            s0 += p[3 * x] * kv;
            s1 += p[3 * x + 1] * kv;
            s2 += p[3 * x + 2] * kv;
        }
        uint8_t* o = tmp + i * 3;
        o[0] = clip8(s0);
        o[1] = clip8(s1);
        o[2] = clip8(s2);
    }
}

// dst[c][yo][xo]: vertical pass over the intermediate, channels-first output
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ tmp, int r0, int n_cols, int n_out, int ksize,
                                                       const int* __restrict__ bounds, const int* __restrict__ kk,
                                                       uint8_t* __restrict__ dst) {
    const int64_t n = (int64_t)n_out * n_cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int xo = (int)(i % n_cols), yo = (int)(i / n_cols);
        const int ymin = bounds[2 * yo], cnt = bounds[2 * yo + 1];
        const int* k = kk + (size_t)yo * ksize;
        const uint8_t* p = tmp + ((int64_t)(ymin - r0) * n_cols + xo) * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < cnt; ++y) {
            const int kv = k[y];
            const uint8_t* q = p + (int64_t)y * n_cols * 3;
            s0 += q[0] * kv;
            s1 += q[1] * kv;
            s2 += q[2] * kv;
        }
        dst[i] = clip8(s0);
        dst[n + i] = clip8(s1);
        dst[2 * n + i] = clip8(s2);
    }
}

// Conservative range of source rows the vertical pass of the crop touches (one row of slack on both sides covers any
// last-bit difference between this host estimate and the device's coefficient kernel).
void row_range(const Axis& v, int* r0, int* r1) {
    const double c_first = ((double)v.crop0 + 0.5) * v.scale, c_last = ((double)(v.crop0 + v.n_out - 1) + 0.5) * v.scale;
    *r0 = std::max(0, (int)(c_first - v.support + 0.5) - 1);
    *r1 = std::min(v.in_size, (int)(c_last + v.support + 0.5) + 1);
}

struct ResizeLayout {
    size_t hb, hk, vb, vk, tmp, total;
};
size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

bool resize_layout(int64_t h, int64_t w, int64_t out_w, int64_t out_h, int64_t cl, int64_t ct, int64_t cw, int64_t ch, Axis* ah,
                   Axis* av, ResizeLayout* L) {
    if (h <= 0 || w <= 0 || out_w <= 0 || out_h <= 0 || cw <= 0 || ch <= 0 || cl < 0 || ct < 0 || cl + cw > out_w || ct + ch > out_h)
        return false;
    if (h >= (1 << 24) || w >= (1 << 24) || out_w >= (1 << 24) || out_h >= (1 << 24)) return false;
    *ah = make_axis(w, out_w, cl, cw);
    *av = make_axis(h, out_h, ct, ch);
    int r0, r1;
    row_range(*av, &r0, &r1);
    size_t o = 0;
    L->hb = o;  o += align256((size_t)cw * 2 * 4);
    L->hk = o;  o += align256((size_t)cw * ah->ksize * 4);
    L->vb = o;  o += align256((size_t)ch * 2 * 4);
    L->vk = o;  o += align256((size_t)ch * av->ksize * 4);
    L->tmp = o; o += align256((size_t)(r1 - r0) * cw * 3);
    L->total = o;
    return true;
}

}  // namespace

extern "C" size_t slb_resize_workspace_bytes(int64_t h, int64_t w, int64_t out_w, int64_t out_h, int64_t crop_left,
                                             int64_t crop_top, int64_t crop_w, int64_t crop_h) {
    Axis ah, av;
    ResizeLayout L;
    if (!resize_layout(h, w, out_w, out_h, crop_left, crop_top, crop_w, crop_h, &ah, &av, &L)) return 0;
    return L.total;
}

extern "C" int slb_resize_bicubic_u8(const uint8_t* src_hwc, int64_t h, int64_t w, int64_t out_w, int64_t out_h,
                                     int64_t crop_left, int64_t crop_top, int64_t crop_w, int64_t crop_h, uint8_t* dst_chw,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    Axis ah, av;
    ResizeLayout L;
    SLB_REQUIRE(resize_layout(h, w, out_w, out_h, crop_left, crop_top, crop_w, crop_h, &ah, &av, &L), SLB_EINVAL,
                "slb_resize_bicubic_u8: bad sizes (%lld x %lld -> %lld x %lld, crop %lld,%lld %lld x %lld)", (long long)w,
                (long long)h, (long long)out_w, (long long)out_h, (long long)crop_left, (long long)crop_top, (long long)crop_w,
                (long long)crop_h);
    SLB_REQUIRE(src_hwc && dst_chw && workspace, SLB_EINVAL, "slb_resize_bicubic_u8: null pointer");
    SLB_REQUIRE(((uintptr_t)workspace % 256) == 0, SLB_EINVAL, "slb_resize_bicubic_u8: workspace must be 256-byte aligned");
    SLB_REQUIRE(workspace_bytes >= L.total, SLB_EWORKSPACE, "slb_resize_bicubic_u8: workspace needs %zu bytes, got %zu", L.total,
                workspace_bytes);
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    int* hb = reinterpret_cast<int*>(ws + L.hb);
    int* hk = reinterpret_cast<int*>(ws + L.hk);
    int* vb = reinterpret_cast<int*>(ws + L.vb);
    int* vk = reinterpret_cast<int*>(ws + L.vk);
    uint8_t* tmp = ws + L.tmp;
    int r0, r1;
    row_range(av, &r0, &r1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SlbProfScope prof("K3 resize_bicubic", stream, 0.0, 3.0 * ((double)(r1 - r0) * (double)w + 2.0 * (double)(r1 - r0) * (double)crop_w +
                                                               (double)crop_w * (double)crop_h));
    resize_coeffs_kernel<<<(unsigned)slb_ceil_div(crop_w, 128), 128, 0, st>>>(ah, hb, hk);
    SLB_LAUNCH_OK("resize_coeffs");
    resize_coeffs_kernel<<<(unsigned)slb_ceil_div(crop_h, 128), 128, 0, st>>>(av, vb, vk);
    SLB_LAUNCH_OK("resize_coeffs");
    const int64_t nh = (int64_t)(r1 - r0) * crop_w, nv = crop_w * crop_h;
    resize_h_kernel<<<(unsigned)std::min<int64_t>(slb_ceil_div(nh, 256), 4096), 256, 0, st>>>(src_hwc, (int)w, r0, r1 - r0, (int)crop_w,
                                                                                             ah.ksize, hb, hk, tmp);
    SLB_LAUNCH_OK("resize_h");
    resize_v_kernel<<<(unsigned)std::min<int64_t>(slb_ceil_div(nv, 256), 4096), 256, 0, st>>>(tmp, r0, (int)crop_w, (int)crop_h, av.ksize,
                                                                                             vb, vk, dst_chw);
    SLB_LAUNCH_OK("resize_v");
    return SLB_OK;
}
