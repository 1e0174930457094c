// K4 attention for head_dim 64 — softmax(scale · Q Kᵀ) V of nn.MultiheadAttention inside the CLIP / SigLIP ViT blocks
// (reference foundation_models/clip.py:118 -> open_clip VisionTransformer -> nn.MultiheadAttention), on the tensor cores
// with fp32-grade accuracy.
//
// Sequences here are short (50 / 197 / 257 tokens) and heads are 64 wide, so one (image, head) is far below a tcgen05
// tile; this kernel uses warp-level mma.sync m16n8k16 (fp16 x fp16 -> fp32) in the FlashAttention-2 arrangement: a warp
// owns 16 query rows, the head's K and V live in shared memory for the whole CTA, scores never leave registers.
// fp32-grade: every operand is split x = hi + lo/2^11 (two fp16, 22 significant bits) and each product is three MMAs
// (hi·hi into the main accumulator; hi·lo and lo·hi into a correction accumulator that is added back /2^11), for both
// S = Q Kᵀ and O = P V; the softmax itself (max, exp, sum) is fp32.
//
// Shared-memory layouts are chosen so that every B-fragment register is ONE conflict-free 32-bit load:
//   K   [key][72]  fp16 (row = key, 64 dims + 8 pad)        -> b = K[key0 + lane/4][d0 + 2*(lane%4) .. +1]
//   V^T [dim][Tkp+8] fp16 (row = dim, keys contiguous)       -> b = V^T[d0 + lane/4][key0 + 2*(lane%4) .. +1]
#include "tc_common.cuh"

namespace {

constexpr int kDh = 64;
constexpr int kKPad = 72;     // K row stride in halves (36 words: 8 consecutive keys hit 8 distinct bank quads)
constexpr int kKeyBlock = 64; // keys per online-softmax block

struct AttnMmaParams {
    const float* q; int64_t q_bs, q_rs;
    const float* k; const float* v; int64_t kv_bs, kv_rs;
    int Tq, Tk, Tkp, H;
    float scale;
    float* out_f32; uint16_t* out_hi; uint16_t* out_lo; int fmt;
    int64_t out_rows_per_batch;  // = Tq
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x0, x1) -> packed fp16 hi pair and packed fp16 (lo * 2^11) pair
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    x0 = fminf(fmaxf(x0, -65504.0f), 65504.0f);
    x1 = fminf(fmaxf(x1, -65504.0f), 65504.0f);
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((x0 - hf.x) * 2048.0f, (x1 - hf.y) * 2048.0f);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) attention_mma_kernel(AttnMmaParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tkp = p.Tkp;
    const int vstride = Tkp + 8;
    __half* Kh = reinterpret_cast<__half*>(smem_raw);   // [Tkp][72]
    __half* Kl = Kh + (size_t)Tkp * kKPad;
    __half* Vh = Kl + (size_t)Tkp * kKPad;              // [64][Tkp + 8]
    __half* Vl = Vh + (size_t)kDh * vstride;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.x / p.H;
    const int h = (int)(blockIdx.x % p.H);
    const float* kb = p.k + b * p.kv_bs + (int64_t)h * kDh;
    const float* vb = p.v + b * p.kv_bs + (int64_t)h * kDh;

    // ---- stage K (row-major) and V (transposed) of this head as fp16 hi / lo planes ----
    for (int i = tid; i < Tkp * (kDh / 4); i += NW * 32) {
        const int key = i >> 4, d4 = (i & 15) * 4;
        float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
        if (key < p.Tk) {
            kk = *reinterpret_cast<const float4*>(kb + (int64_t)key * p.kv_rs + d4);
            vv = *reinterpret_cast<const float4*>(vb + (int64_t)key * p.kv_rs + d4);
        }
        uint32_t h0, l0, h1, l1;
        split_pair(kk.x, kk.y, h0, l0);
        split_pair(kk.z, kk.w, h1, l1);
        *reinterpret_cast<uint2*>(Kh + (size_t)key * kKPad + d4) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(Kl + (size_t)key * kKPad + d4) = make_uint2(l0, l1);
        split_pair(vv.x, vv.y, h0, l0);
        split_pair(vv.z, vv.w, h1, l1);
        const __half* hh0 = reinterpret_cast<const __half*>(&h0);
        const __half* hh1 = reinterpret_cast<const __half*>(&h1);
        const __half* ll0 = reinterpret_cast<const __half*>(&l0);
        const __half* ll1 = reinterpret_cast<const __half*>(&l1);
        Vh[(size_t)(d4 + 0) * vstride + key] = hh0[0];
        Vh[(size_t)(d4 + 1) * vstride + key] = hh0[1];
        Vh[(size_t)(d4 + 2) * vstride + key] = hh1[0];
        Vh[(size_t)(d4 + 3) * vstride + key] = hh1[1];
        Vl[(size_t)(d4 + 0) * vstride + key] = ll0[0];
        Vl[(size_t)(d4 + 1) * vstride + key] = ll0[1];
        Vl[(size_t)(d4 + 2) * vstride + key] = ll1[0];
        Vl[(size_t)(d4 + 3) * vstride + key] = ll1[1];
    }
    __syncthreads();

    const int r0 = blockIdx.y * (16 * NW) + warp * 16;
    if (r0 >= p.Tq) return;
    const int g = lane >> 2, t4 = lane & 3;

    // ---- Q fragments (scaled), split ----
    uint32_t qh[4][4], ql[4][4];
    {
        const float* qb = p.q + b * p.q_bs + (int64_t)h * kDh;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = r0 + g + (i & 1) * 8;
                const int d = ks * 16 + t4 * 2 + (i >> 1) * 8;
                float2 x = make_float2(0.f, 0.f);
                if (row < p.Tq) x = *reinterpret_cast<const float2*>(qb + (int64_t)row * p.q_rs + d);
                split_pair(x.x * p.scale, x.y * p.scale, qh[ks][i], ql[ks][i]);
            }
        }
    }

    float om[8][4], oc[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
        for (int j = 0; j < 4; ++j) { om[dt][j] = 0.f; oc[dt][j] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    constexpr float kInvS = 1.0f / 2048.0f;

    for (int kb0 = 0; kb0 < Tkp; kb0 += kKeyBlock) {
        const int nkt = min(kKeyBlock, Tkp - kb0) >> 3;  // n-tiles of 8 keys in this block (even: Tkp % 16 == 0)
        float s[8][4];
        // ---- S = Q K^T ----
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float sm_[4] = {0.f, 0.f, 0.f, 0.f}, sc_[4] = {0.f, 0.f, 0.f, 0.f};
            if (nt < nkt) {
                const int key = kb0 + nt * 8 + g;
                const __half* krh = Kh + (size_t)key * kKPad + t4 * 2;
                const __half* krl = Kl + (size_t)key * kKPad + t4 * 2;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(krh + ks * 16);
                    const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(krh + ks * 16 + 8);
                    const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(krl + ks * 16);
                    const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(krl + ks * 16 + 8);
                    mma16816(sm_, qh[ks], bh0, bh1);
                    mma16816(sc_, qh[ks], bl0, bl1);
                    mma16816(sc_, ql[ks], bh0, bh1);
                }
            }
            const int col = kb0 + nt * 8 + t4 * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool ok = nt < nkt && (col + (j & 1)) < p.Tk;
                s[nt][j] = ok ? fmaf(sc_[j], kInvS, sm_[j]) : -INFINITY;
            }
        }
        // ---- online softmax (rows g and g + 8) ----
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float m_new = fmaxf(m_run[r], mx[r]);
            alpha[r] = (m_run[r] == -INFINITY) ? 0.f : expf(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = expf(s[nt][j] - m_run[j >> 1]);  // exp(-inf) = 0 for masked keys
                s[nt][j] = e;
                rs[j >> 1] += e;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            om[dt][0] *= alpha[0]; om[dt][1] *= alpha[0]; om[dt][2] *= alpha[1]; om[dt][3] *= alpha[1];
            oc[dt][0] *= alpha[0]; oc[dt][1] *= alpha[0]; oc[dt][2] *= alpha[1]; oc[dt][3] *= alpha[1];
        }
        // ---- O += P V ----
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            if (2 * kt < nkt) {
                uint32_t ph[4], pl[4];
                split_pair(s[2 * kt][0], s[2 * kt][1], ph[0], pl[0]);
                split_pair(s[2 * kt][2], s[2 * kt][3], ph[1], pl[1]);
                split_pair(s[2 * kt + 1][0], s[2 * kt + 1][1], ph[2], pl[2]);
                split_pair(s[2 * kt + 1][2], s[2 * kt + 1][3], ph[3], pl[3]);
                const int key = kb0 + kt * 16 + t4 * 2;
#pragma unroll
                for (int dt = 0; dt < 8; ++dt) {
                    const __half* vrh = Vh + (size_t)(dt * 8 + g) * vstride + key;
                    const __half* vrl = Vl + (size_t)(dt * 8 + g) * vstride + key;
                    const uint32_t bh0 = *reinterpret_cast<const uint32_t*>(vrh);
                    const uint32_t bh1 = *reinterpret_cast<const uint32_t*>(vrh + 8);
                    const uint32_t bl0 = *reinterpret_cast<const uint32_t*>(vrl);
                    const uint32_t bl1 = *reinterpret_cast<const uint32_t*>(vrl + 8);
                    mma16816(om[dt], ph, bh0, bh1);
                    mma16816(oc[dt], ph, bl0, bl1);
                    mma16816(oc[dt], pl, bh0, bh1);
                }
            }
        }
    }

    // ---- normalise and store ----
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const int Wd = p.H * kDh;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = r0 + g + r * 8;
        if (row >= p.Tq) continue;
        const float inv = 1.0f / l_run[r];
        const int64_t base = (b * p.out_rows_per_batch + row) * Wd + (int64_t)h * kDh + t4 * 2;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            const float o0 = fmaf(oc[dt][2 * r], kInvS, om[dt][2 * r]) * inv;
            const float o1 = fmaf(oc[dt][2 * r + 1], kInvS, om[dt][2 * r + 1]) * inv;
            if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + base + dt * 8) = make_float2(o0, o1);
            if (p.out_hi) {
                uint16_t h0, l0, h1, l1;
                slb_split2(o0, p.fmt, h0, l0);
                slb_split2(o1, p.fmt, h1, l1);
                *reinterpret_cast<uint32_t*>(p.out_hi + base + dt * 8) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                *reinterpret_cast<uint32_t*>(p.out_lo + base + dt * 8) = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Same algorithm, operands already split: q, k, v are read from the fp16 split planes [2][rows][3W] that the in_proj
// GEMM emits (hi, lo * 2^11), so nothing is converted here — K and V go to shared memory with 16-byte cp.async copies,
// B fragments come from ldmatrix (.trans for V, which stays row-major), Q fragments are 32-bit global loads. The softmax
// scale multiplies the fp32 logits.
// ------------------------------------------------------------------------------------------------
struct AttnPlanesParams {
    const __half* hi;  // [rows][3W]
    const __half* lo;
    int T, Tkp, H, W;
    int causal;  // 1: key j is visible to query i only if j <= i (CLIP text tower)
    float scale;
    float* out_f32; uint16_t* out_hi; uint16_t* out_lo; int fmt;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slb_smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(slb_smem_u32(smem_row)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(slb_smem_u32(smem_row)));
}

template <int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) attention_planes_kernel(AttnPlanesParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tkp = p.Tkp;
    __half* Kh = reinterpret_cast<__half*>(smem_raw);  // four [Tkp][72] tiles
    __half* Kl = Kh + (size_t)Tkp * kKPad;
    __half* Vh = Kl + (size_t)Tkp * kKPad;
    __half* Vl = Vh + (size_t)Tkp * kKPad;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.x / p.H;
    const int h = (int)(blockIdx.x % p.H);
    const int64_t ld = 3 * (int64_t)p.W;
    const int64_t row0 = b * p.T;

    // ---- stage K and V (hi and lo) with 16-byte async copies; keys >= T are zero ----
    for (int i = tid; i < Tkp * 32; i += NW * 32) {
        const int key = i >> 5, which = (i >> 3) & 3, chunk = i & 7;  // which: 0 Kh, 1 Kl, 2 Vh, 3 Vl
        __half* dst = Kh + (size_t)which * Tkp * kKPad + (size_t)key * kKPad + chunk * 8;
        if (key < p.T) {
            const __half* plane = (which & 1) ? p.lo : p.hi;
            const __half* src = plane + (row0 + key) * ld + ((which >> 1) ? 2 : 1) * p.W + h * kDh + chunk * 8;
            cp_async16(dst, src);
        } else {
            *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    const int r0 = blockIdx.y * (16 * NW) + warp * 16;
    const int g = lane >> 2, t4 = lane & 3;
    // ---- Q fragments straight from the planes (overlaps the copies above) ----
    uint32_t qh[4][4], ql[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = r0 + g + (i & 1) * 8;
            const int d = ks * 16 + t4 * 2 + (i >> 1) * 8;
            uint32_t vh = 0u, vl = 0u;
            if (row < p.T) {
                const int64_t off = (row0 + row) * ld + h * kDh + d;
                vh = *reinterpret_cast<const uint32_t*>(p.hi + off);
                vl = *reinterpret_cast<const uint32_t*>(p.lo + off);
            }
            qh[ks][i] = vh;
            ql[ks][i] = vl;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (r0 >= p.T) return;

    float om[8][4], oc[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
        for (int j = 0; j < 4; ++j) { om[dt][j] = 0.f; oc[dt][j] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    constexpr float kInvS = 1.0f / 2048.0f;
    const float sc_main = p.scale, sc_corr = p.scale * kInvS;
    // ldmatrix row addresses: lane -> (matrix j = lane / 8, row r = lane % 8)
    const int lm_j = lane >> 3, lm_r = lane & 7;

    for (int kb0 = 0; kb0 < Tkp; kb0 += kKeyBlock) {
        const int nkt = min(kKeyBlock, Tkp - kb0) >> 3;
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float sm_[4] = {0.f, 0.f, 0.f, 0.f}, sc_[4] = {0.f, 0.f, 0.f, 0.f};
            if (nt < nkt) {
                // matrices: (keys nt*8.., d = ks2*32 + j*8 ..): j = 0,1 -> k-step 2*ks2 (b0, b1); j = 2,3 -> k-step 2*ks2 + 1
                const size_t roff = (size_t)(kb0 + nt * 8 + lm_r) * kKPad + lm_j * 8;
#pragma unroll
                for (int ks2 = 0; ks2 < 2; ++ks2) {
                    uint32_t bh[4], bl[4];
                    ldmatrix_x4(bh, Kh + roff + ks2 * 32);
                    ldmatrix_x4(bl, Kl + roff + ks2 * 32);
                    mma16816(sm_, qh[2 * ks2], bh[0], bh[1]);
                    mma16816(sc_, qh[2 * ks2], bl[0], bl[1]);
                    mma16816(sc_, ql[2 * ks2], bh[0], bh[1]);
                    mma16816(sm_, qh[2 * ks2 + 1], bh[2], bh[3]);
                    mma16816(sc_, qh[2 * ks2 + 1], bl[2], bl[3]);
                    mma16816(sc_, ql[2 * ks2 + 1], bh[2], bh[3]);
                }
            }
            const int col = kb0 + nt * 8 + t4 * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = col + (j & 1);
                const bool ok = nt < nkt && c < p.T && (!p.causal || c <= r0 + g + (j >> 1) * 8);
                s[nt][j] = ok ? fmaf(sc_[j], sc_corr, sm_[j] * sc_main) : -INFINITY;
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float alpha[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float m_new = fmaxf(m_run[r], mx[r]);
            alpha[r] = (m_run[r] == -INFINITY) ? 0.f : expf(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = expf(s[nt][j] - m_run[j >> 1]);
                s[nt][j] = e;
                rs[j >> 1] += e;
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
        if (kb0 > 0) {
#pragma unroll
            for (int dt = 0; dt < 8; ++dt) {
                om[dt][0] *= alpha[0]; om[dt][1] *= alpha[0]; om[dt][2] *= alpha[1]; om[dt][3] *= alpha[1];
                oc[dt][0] *= alpha[0]; oc[dt][1] *= alpha[0]; oc[dt][2] *= alpha[1]; oc[dt][3] *= alpha[1];
            }
        }
#pragma unroll
        for (int kt = 0; kt < 4; ++kt) {
            if (2 * kt < nkt) {
                uint32_t ph[4], pl[4];
                split_pair(s[2 * kt][0], s[2 * kt][1], ph[0], pl[0]);
                split_pair(s[2 * kt][2], s[2 * kt][3], ph[1], pl[1]);
                split_pair(s[2 * kt + 1][0], s[2 * kt + 1][1], ph[2], pl[2]);
                split_pair(s[2 * kt + 1][2], s[2 * kt + 1][3], ph[3], pl[3]);
                // matrices: j & 1 -> keys +8 (b1), j >> 1 -> next 8 dims (n-tile dt + 1)
                const size_t roff = (size_t)(kb0 + kt * 16 + (lm_j & 1) * 8 + lm_r) * kKPad + (lm_j >> 1) * 8;
#pragma unroll
                for (int dt2 = 0; dt2 < 4; ++dt2) {
                    uint32_t bh[4], bl[4];
                    ldmatrix_x4_trans(bh, Vh + roff + dt2 * 16);
                    ldmatrix_x4_trans(bl, Vl + roff + dt2 * 16);
                    mma16816(om[2 * dt2], ph, bh[0], bh[1]);
                    mma16816(oc[2 * dt2], ph, bl[0], bl[1]);
                    mma16816(oc[2 * dt2], pl, bh[0], bh[1]);
                    mma16816(om[2 * dt2 + 1], ph, bh[2], bh[3]);
                    mma16816(oc[2 * dt2 + 1], ph, bl[2], bl[3]);
                    mma16816(oc[2 * dt2 + 1], pl, bh[2], bh[3]);
                }
            }
        }
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = r0 + g + r * 8;
        if (row >= p.T) continue;
        const float inv = 1.0f / l_run[r];
        const int64_t base = (row0 + row) * p.W + (int64_t)h * kDh + t4 * 2;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            const float o0 = fmaf(oc[dt][2 * r], kInvS, om[dt][2 * r]) * inv;
            const float o1 = fmaf(oc[dt][2 * r + 1], kInvS, om[dt][2 * r + 1]) * inv;
            if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + base + dt * 8) = make_float2(o0, o1);
            if (p.out_hi) {
                uint16_t h0, l0, h1, l1;
                slb_split2(o0, p.fmt, h0, l0);
                slb_split2(o1, p.fmt, h1, l1);
                *reinterpret_cast<uint32_t*>(p.out_hi + base + dt * 8) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                *reinterpret_cast<uint32_t*>(p.out_lo + base + dt * 8) = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        }
    }
}

template <int NW, int MINB>
int launch_attn_planes(const AttnPlanesParams& p, int64_t B, cudaStream_t st) {
    const size_t smem = (size_t)4 * p.Tkp * kKPad * sizeof(__half);
    SLB_REQUIRE(smem <= 227 * 1024, SLB_EUNSUPPORTED, "slb_attention_planes: K/V of one head do not fit shared memory");
    if (smem > 48 * 1024)
        SLB_CUDA_OK(cudaFuncSetAttribute(attention_planes_kernel<NW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(B * p.H), (unsigned)slb_ceil_div(p.T, 16 * NW));
    attention_planes_kernel<NW, MINB><<<grid, NW * 32, smem, st>>>(p);
    SLB_LAUNCH_OK("attention_planes");
    return SLB_OK;
}

template <int NW>
int launch_attn(const AttnMmaParams& p, int64_t B, cudaStream_t st) {
    const size_t smem = ((size_t)2 * p.Tkp * kKPad + (size_t)2 * kDh * (p.Tkp + 8)) * sizeof(__half);
    SLB_REQUIRE(smem <= 227 * 1024, SLB_EUNSUPPORTED, "slb_attention_small: K/V of one head do not fit shared memory");
    if (smem > 48 * 1024)
        SLB_CUDA_OK(cudaFuncSetAttribute(attention_mma_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(B * p.H), (unsigned)slb_ceil_div(p.Tq, 16 * NW));
    attention_mma_kernel<NW><<<grid, NW * 32, smem, st>>>(p);
    SLB_LAUNCH_OK("attention_mma");
    return SLB_OK;
}

}  // namespace

// Called by slb_attention_small (vit_kernels.cu) when head_dim == 64 and the strides allow vector loads.
int slb_attention_mma_dh64(const float* q, int64_t q_bs, int64_t q_rs, const float* k, const float* v, int64_t kv_bs,
                           int64_t kv_rs, int64_t B, int64_t Tq, int64_t Tk, int64_t H, float scale, int plane_fmt,
                           float* out_f32, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st) {
    AttnMmaParams p{};
    p.q = q; p.q_bs = q_bs; p.q_rs = q_rs;
    p.k = k; p.v = v; p.kv_bs = kv_bs; p.kv_rs = kv_rs;
    p.Tq = (int)Tq; p.Tk = (int)Tk; p.Tkp = (int)((Tk + 15) / 16 * 16); p.H = (int)H;
    p.scale = scale;
    p.out_f32 = out_f32; p.out_hi = out_hi; p.out_lo = out_lo; p.fmt = plane_fmt;
    p.out_rows_per_batch = Tq;
    SLB_REQUIRE(B * H <= 0x7FFFFFFF, SLB_EUNSUPPORTED, "slb_attention_small: grid too large");
    if (Tq <= 64) return launch_attn<4>(p, B, st);
    return launch_attn<8>(p, B, st);
}

extern "C" int slb_attention_planes(const uint16_t* qkv_planes, int64_t B, int64_t T, int64_t H, int64_t dh, float scale,
                                    int causal, int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && T > 0 && H > 0, SLB_EINVAL, "slb_attention_planes: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(qkv_planes && (out_f32 || out_planes), SLB_EINVAL, "slb_attention_planes: null pointer");
    SLB_REQUIRE(dh == kDh, SLB_EUNSUPPORTED, "slb_attention_planes: head_dim must be 64 (got %lld)", (long long)dh);
    SLB_REQUIRE(plane_fmt == SLB_PLANE_F16, SLB_EUNSUPPORTED, "slb_attention_planes: fp16 planes only");
    SLB_REQUIRE(((uintptr_t)qkv_planes % 16) == 0, SLB_EINVAL, "slb_attention_planes: planes must be 16-byte aligned");
    SLB_REQUIRE(B * H <= 0x7FFFFFFF && T <= 4096, SLB_EUNSUPPORTED, "slb_attention_planes: problem too large");
    const int64_t W = H * dh, rows = B * T;
    SlbProfScope prof("K4 attention", stream, 4.0 * (double)B * (double)H * (double)T * (double)T * (double)dh * 3.0,
                      4.0 * (double)rows * (double)W * 4.0);
    AttnPlanesParams p{};
    p.hi = reinterpret_cast<const __half*>(qkv_planes);
    p.lo = p.hi + rows * 3 * W;
    p.T = (int)T; p.Tkp = (int)((T + 15) / 16 * 16); p.H = (int)H; p.W = (int)W;
    p.scale = scale;
    p.causal = causal ? 1 : 0;
    p.out_f32 = out_f32; p.out_hi = out_planes; p.out_lo = out_planes ? out_planes + rows * W : nullptr; p.fmt = plane_fmt;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (T <= 64) return launch_attn_planes<4, 3>(p, B, st);
    return launch_attn_planes<8, 1>(p, B, st);
}
