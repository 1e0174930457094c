// K4 attention for head_dim 64 — softmax(scale · Q Kᵀ) V of nn.MultiheadAttention inside the CLIP / SigLIP ViT blocks
// (reference foundation_models/clip.py:118 -> open_clip VisionTransformer -> nn.MultiheadAttention), on the tensor cores
// with fp32-grade accuracy.
//
// Sequences here are short (50 / 197 / 257 tokens) and heads are 64 wide, so one (image, head) is far below a tcgen05
// tile; this kernel uses warp-level mma.sync m16n8k16 (fp16 x fp16 -> fp32) in the FlashAttention-2 arrangement: a warp
// owns 16 query rows, the head's K and V live in shared memory for the whole CTA, scores never leave registers.
// fp32-grade: every operand is split x·s = hi + lo (two fp16, 22 significant bits, s = SLB_ACT_PLANE_SCALE shared by q, k
// and v) and each product is three MMAs hi·hi + hi·lo + lo·hi, for both S = Q Kᵀ and O = P V (P planes carry 2^10); the
// cross terms accumulate apart from hi·hi and are added in fp32 at the end; the softmax itself (max, exp, sum) is fp32.
//
// Shared-memory layouts are chosen so that every B-fragment register is ONE conflict-free 32-bit load:
//   K   [key][72]  fp16 (row = key, 64 dims + 8 pad)        -> b = K[key0 + lane/4][d0 + 2*(lane%4) .. +1]
//   V^T [dim][Tkp+8] fp16 (row = dim, keys contiguous)       -> b = V^T[d0 + lane/4][key0 + 2*(lane%4) .. +1]
#include "tc_common.cuh"

#include <stdlib.h>

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

// the same for values known to lie in [0, 1] (softmax numerators): no saturation needed
// Probabilities p in [0, 1] as split planes with the scale kPScale (hi + lo = p * 2^10; lo stays a normal fp16 down to
// p ~ 2^-14 and is exact to 2^-34 below that), the same format as the GEMM operands.
constexpr float kPScale = 1024.0f;
__device__ __forceinline__ void split_pair_unit(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    x0 *= kPScale;
    x1 *= kPScale;
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
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
                slb_split2_act(o0, p.fmt, h0, l0);
                slb_split2_act(o1, p.fmt, h1, l1);
                *reinterpret_cast<uint32_t*>(p.out_hi + base + dt * 8) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                *reinterpret_cast<uint32_t*>(p.out_lo + base + dt * 8) = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Same algorithm, operands already split: q, k, v are read from the fp16 split planes [2][rows][3W] that the in_proj
// GEMM emits (hi, lo at SLB_ACT_PLANE_SCALE), so nothing is converted here — K and V go to shared memory with 16-byte cp.async copies,
// B fragments come from ldmatrix (.trans for V, which stays row-major), Q fragments are 32-bit global loads. The softmax
// scale multiplies the fp32 logits.
// ------------------------------------------------------------------------------------------------
struct AttnPlanesParams {
    const __half* hi;  // [rows][3W]
    const __half* lo;
    int T, Tkp, H, W;
    int q_row0;  // first query row handled by this launch (rows below it belong to the tcgen05 tiles)
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

// kSplit (a tail of <= 16 query rows, e.g. the class token left over by the tcgen05 tiles of a 257-token tower): the NW
// warps share the one 16-row query tile and split every 64-key block four ways (warp w takes keys [16w, 16w + 16) of
// each block), then merge their (max, sum, O) through shared memory — the serial chain per CTA is a quarter as long.
template <int NW, int MINB, bool kSplit = false>
__global__ void __launch_bounds__(NW * 32, MINB) attention_planes_kernel(AttnPlanesParams p) {
    static_assert(!kSplit || NW == 4, "key splitting assumes 4 warps x 16 keys per block");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Tkp = p.Tkp;
    // one stage = a 64-key block of Kh, Kl, Vh, Vl ([64][72] halves each); two stages when the sequence has several blocks:
    // block j + 1 streams in (cp.async) while block j is being multiplied
    constexpr int kTileHalves = kKeyBlock * kKPad;
    constexpr int kStageHalves = 4 * kTileHalves;
    __half* smem = reinterpret_cast<__half*>(smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.x / p.H;
    const int h = (int)(blockIdx.x % p.H);
    const int64_t ld = 3 * (int64_t)p.W;
    const int64_t row0 = b * p.T;
    const int nblk = (Tkp + kKeyBlock - 1) / kKeyBlock;

    auto stage = [&](int blk) {  // 16-byte async copies; keys >= T are zero
        __half* base = smem + (size_t)(blk & 1) * kStageHalves;
        for (int i = tid; i < kKeyBlock * 32; i += NW * 32) {
            const int kl = i >> 5, which = (i >> 3) & 3, chunk = i & 7;  // which: 0 Kh, 1 Kl, 2 Vh, 3 Vl
            const int key = blk * kKeyBlock + kl;
            __half* dst = base + (size_t)which * kTileHalves + (size_t)kl * kKPad + chunk * 8;
            if (key < p.T) {
                const __half* plane = (which & 1) ? p.lo : p.hi;
                cp_async16(dst, plane + (row0 + key) * ld + ((which >> 1) ? 2 : 1) * p.W + h * kDh + chunk * 8);
            } else if (key < Tkp) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0);

    const int r0 = kSplit ? p.q_row0 : p.q_row0 + blockIdx.y * (16 * NW) + warp * 16;
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
    const bool active = r0 < p.T;  // warps past the last query row still help to stage K / V

    float om[8][4], oc[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt)
#pragma unroll
        for (int j = 0; j < 4; ++j) { om[dt][j] = 0.f; oc[dt][j] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};
    // q, k, v planes carry SLB_ACT_PLANE_SCALE with unscaled lo planes: the three products of Q K^T share one accumulator.
    // Logits are kept in the log2 domain (scale * log2 e folded in) so that the exponentials are single ex2 instructions.
    constexpr float kInvAct = 1.0f / SLB_ACT_PLANE_SCALE;
    const float sc_main = p.scale * 1.4426950408889634f * kInvAct * kInvAct;
    // ldmatrix row addresses: lane -> (matrix j = lane / 8, row r = lane % 8)
    const int lm_j = lane >> 3, lm_r = lane & 7;

    for (int blk = 0; blk < nblk; ++blk) {
        const int kb0 = blk * kKeyBlock;
        if (blk + 1 < nblk) {
            stage(blk + 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // block `blk` has landed for every thread's copies
        const __half* Kh = smem + (size_t)(blk & 1) * kStageHalves;
        const __half* Kl = Kh + kTileHalves;
        const __half* Vh = Kl + kTileHalves;
        const __half* Vl = Vh + kTileHalves;
        if (active) {
        const int nkt = min(kKeyBlock, Tkp - kb0) >> 3;
        const bool need_mask = (kb0 + kKeyBlock > p.T) || (p.causal && kb0 + kKeyBlock > r0);
        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float sm_[4] = {0.f, 0.f, 0.f, 0.f};
            const bool mine = !kSplit || (nt >> 1) == warp;
            if (nt < nkt && mine) {
                // matrices: (keys nt*8.., d = ks2*32 + j*8 ..): j = 0,1 -> k-step 2*ks2 (b0, b1); j = 2,3 -> k-step 2*ks2 + 1
                const size_t roff = (size_t)(nt * 8 + lm_r) * kKPad + lm_j * 8;
#pragma unroll
                for (int ks2 = 0; ks2 < 2; ++ks2) {
                    uint32_t bh[4], bl[4];
                    ldmatrix_x4(bh, Kh + roff + ks2 * 32);
                    ldmatrix_x4(bl, Kl + roff + ks2 * 32);
                    mma16816(sm_, qh[2 * ks2], bh[0], bh[1]);
                    mma16816(sm_, qh[2 * ks2], bl[0], bl[1]);
                    mma16816(sm_, ql[2 * ks2], bh[0], bh[1]);
                    mma16816(sm_, qh[2 * ks2 + 1], bh[2], bh[3]);
                    mma16816(sm_, qh[2 * ks2 + 1], bl[2], bl[3]);
                    mma16816(sm_, ql[2 * ks2 + 1], bh[2], bh[3]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) s[nt][j] = sm_[j] * sc_main;
            if (need_mask || kSplit) {  // block-uniform: the last key block (padding), causal blocks on / above the diagonal
                const int col = kb0 + nt * 8 + t4 * 2;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = col + (j & 1);
                    const bool ok = mine && nt < nkt && c < p.T && (!p.causal || c <= r0 + g + (j >> 1) * 8);
                    if (!ok) s[nt][j] = -INFINITY;
                }
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
            alpha[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] - m_new);
            m_run[r] = m_new;
        }
        float rs[2] = {0.f, 0.f};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // (a warp of the split variant may not have seen a valid key yet: keep exp2(-inf - -inf) out)
                const float mr = kSplit && m_run[j >> 1] == -INFINITY ? 0.f : m_run[j >> 1];
                const float e = exp2f(s[nt][j] - mr);
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
            if (2 * kt < nkt && (!kSplit || kt == warp)) {
                uint32_t ph[4], pl[4];
                split_pair_unit(s[2 * kt][0], s[2 * kt][1], ph[0], pl[0]);
                split_pair_unit(s[2 * kt][2], s[2 * kt][3], ph[1], pl[1]);
                split_pair_unit(s[2 * kt + 1][0], s[2 * kt + 1][1], ph[2], pl[2]);
                split_pair_unit(s[2 * kt + 1][2], s[2 * kt + 1][3], ph[3], pl[3]);
                // matrices: j & 1 -> keys +8 (b1), j >> 1 -> next 8 dims (n-tile dt + 1)
                const size_t roff = (size_t)(kt * 16 + (lm_j & 1) * 8 + lm_r) * kKPad + (lm_j >> 1) * 8;
#pragma unroll
                for (int dt2 = 0; dt2 < 4; ++dt2) {
                    uint32_t bh[4], bl[4];
                    ldmatrix_x4_trans(bh, Vh + roff + dt2 * 16);
                    ldmatrix_x4_trans(bl, Vl + roff + dt2 * 16);
                    mma16816(om[2 * dt2], ph, bh[0], bh[1]);
                    mma16816(oc[2 * dt2], ph, bl[0], bl[1]);      // the two cross terms (same scale as hi . hi) are kept
                    mma16816(oc[2 * dt2], pl, bh[0], bh[1]);      // apart and added in fp32 at the end
                    mma16816(om[2 * dt2 + 1], ph, bh[2], bh[3]);
                    mma16816(oc[2 * dt2 + 1], ph, bl[2], bl[3]);
                    mma16816(oc[2 * dt2 + 1], pl, bh[2], bh[3]);
                }
            }
        }
        }  // active
        if (blk + 2 < nblk) __syncthreads();  // everyone is done with this stage before block blk + 2 overwrites it
    }
    if (!active) return;

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    if constexpr (kSplit) {
        // merge the four key quarters: xm / xl [warp][16 rows], xo [warp][16][64] in the (now idle) staging memory
        __syncthreads();
        float* xm = reinterpret_cast<float*>(smem_raw);
        float* xl = xm + NW * 16;
        float* xo = xl + NW * 16;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = g + r * 8;
            if (t4 == 0) { xm[warp * 16 + row] = m_run[r]; xl[warp * 16 + row] = l_run[r]; }
#pragma unroll
            for (int dt = 0; dt < 8; ++dt)
                *reinterpret_cast<float2*>(xo + ((size_t)warp * 16 + row) * 64 + dt * 8 + t4 * 2) =
                    make_float2(om[dt][2 * r] + oc[dt][2 * r], om[dt][2 * r + 1] + oc[dt][2 * r + 1]);
        }
        __syncthreads();
        const int row = tid >> 3, c0 = (tid & 7) * 8;  // 128 threads x 8 columns = the 16 x 64 tile
        if (r0 + row >= p.T) return;
        float mw[NW], mmax = -INFINITY, lsum = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) { mw[w] = xm[w * 16 + row]; mmax = fmaxf(mmax, mw[w]); }
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float sc = mw[w] == -INFINITY ? 0.f : exp2f(mw[w] - mmax);
            lsum += xl[w * 16 + row] * sc;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] += xo[((size_t)w * 16 + row) * 64 + c0 + j] * sc;
        }
        const float inv = kInvAct / (kPScale * lsum);
        const int64_t base = (row0 + r0 + row) * p.W + (int64_t)h * kDh + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= inv;
        if (p.out_f32) {
            *reinterpret_cast<float4*>(p.out_f32 + base) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(p.out_f32 + base + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
        if (p.out_hi) {
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                uint16_t h0, l0, h1, l1;
                slb_split2_act(o[2 * e], p.fmt, h0, l0);
                slb_split2_act(o[2 * e + 1], p.fmt, h1, l1);
                hh[e] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                ll[e] = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
            *reinterpret_cast<uint4*>(p.out_hi + base) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<uint4*>(p.out_lo + base) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        }
        return;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int row = r0 + g + r * 8;
        if (row >= p.T) continue;
        const float inv = kInvAct / (kPScale * l_run[r]);  // V planes carry the activation scale, P planes kPScale
        const int64_t base = (row0 + row) * p.W + (int64_t)h * kDh + t4 * 2;
#pragma unroll
        for (int dt = 0; dt < 8; ++dt) {
            const float o0 = (om[dt][2 * r] + oc[dt][2 * r]) * inv;
            const float o1 = (om[dt][2 * r + 1] + oc[dt][2 * r + 1]) * inv;
            if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + base + dt * 8) = make_float2(o0, o1);
            if (p.out_hi) {
                uint16_t h0, l0, h1, l1;
                slb_split2_act(o0, p.fmt, h0, l0);
                slb_split2_act(o1, p.fmt, h1, l1);
                *reinterpret_cast<uint32_t*>(p.out_hi + base + dt * 8) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                *reinterpret_cast<uint32_t*>(p.out_lo + base + dt * 8) = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        }
    }
}

// A tail of one to four query rows (the last token left over by the tcgen05 tiles of a 257-token tower) as plain fp32
// SIMT work: one CTA per (image, head), its 4 warps take the 32-key groups round-robin. Per group a lane first owns two of
// the 64 dims (coalesced 128-byte reads of the K hi / lo planes), the 32 partial dot products are summed by a transposing
// butterfly (31 shuffles for 32 keys: lane j ends with key j's logit), then an online-softmax step and the P V update with
// the lane back on its two dims. The warps' (max, sum, O) merge through shared memory. ~6 k instructions per (image, head)
// instead of a 16-row tensor-core tile that is 94 % padding and latency-bound (48 -> ~10 us per layer at B = 64, H = 16).
__global__ void __launch_bounds__(128) attention_rows_kernel(AttnPlanesParams p) {
    __shared__ float s_m[4], s_l[4], s_o[4][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int64_t ld = 3 * (int64_t)p.W;
    const int64_t row_base = (int64_t)b * p.T;
    const uint32_t* khi = reinterpret_cast<const uint32_t*>(p.hi + row_base * ld + p.W + h * 64) + lane;
    const uint32_t* klo = reinterpret_cast<const uint32_t*>(p.lo + row_base * ld + p.W + h * 64) + lane;
    const int64_t ldw = ld / 2;  // row pitch in 32-bit words
    constexpr float kInvAct = 1.0f / SLB_ACT_PLANE_SCALE;
    const float c_log2 = p.scale * 1.4426950408889634f * kInvAct * kInvAct;
    const int ngroups = (p.T + 31) >> 5;
    auto unpack = [](uint32_t hi, uint32_t lo, float& x0, float& x1) {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&lo));
        x0 = a.x + c.x;
        x1 = a.y + c.y;
    };
    for (int qr = p.q_row0; qr < p.T; ++qr) {
        float q0, q1;
        {
            const int64_t off = (row_base + qr) * ld + h * 64;
            unpack(reinterpret_cast<const uint32_t*>(p.hi + off)[lane], reinterpret_cast<const uint32_t*>(p.lo + off)[lane], q0, q1);
        }
        float m = -INFINITY, l = 0.f, o0 = 0.f, o1 = 0.f;
        // the same trip count for every warp (a data-independent loop keeps the shuffles plain SHFLs); a warp whose last
        // round has no group left works on keys past T, which are masked
        for (int g0 = 0; g0 < ngroups; g0 += 4) {
            const int key0 = (g0 + warp) * 32;
            float part[32];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
                // keys past T read the last key again (no branch: the 64 loads of a group stay in flight together); their
                // logits are masked below
                const int64_t ko = (int64_t)min(key0 + jj, p.T - 1) * ldw;
                float k0, k1;
                unpack(khi[ko], klo[ko], k0, k1);
                part[jj] = fmaf(q0, k0, q1 * k1);
            }
            // transposing butterfly: after the step with stride s a lane keeps the half of its values whose key index has bit
            // s equal to its own lane bit; lane j ends with the full dot product of key j
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
                for (int i = 0; i < s; ++i) {
                    const bool up = (lane & s) != 0;
                    const float keep = up ? part[i + s] : part[i], send = up ? part[i] : part[i + s];
                    part[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
                }
            }
            const float logit = key0 + lane < p.T ? part[0] * c_log2 : -INFINITY;
            float gm = logit;
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, s));
            const float m_new = fmaxf(m, gm);
            const float m_use = m_new == -INFINITY ? 0.f : m_new;  // nothing seen yet and an empty group
            const float alpha = exp2f(m - m_use), pj = exp2f(logit - m_use);
            l = l * alpha + slb_warp_sum_butterfly(pj);
            o0 *= alpha;
            o1 *= alpha;
            m = m_new;
            const uint32_t* vhi = khi + p.W / 2;  // V follows K in the packed q | k | v row
            const uint32_t* vlo = klo + p.W / 2;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) {
                const float pv = __shfl_sync(0xffffffffu, pj, jj);  // 0 for keys past T
                const int64_t vo = (int64_t)min(key0 + jj, p.T - 1) * ldw;
                float v0, v1;
                unpack(vhi[vo], vlo[vo], v0, v1);
                o0 = fmaf(pv, v0, o0);
                o1 = fmaf(pv, v1, o1);
            }
        }
        __syncthreads();  // the previous row's merge has been read
        if (lane == 0) { s_m[warp] = m; s_l[warp] = l; }
        s_o[warp][2 * lane] = o0;
        s_o[warp][2 * lane + 1] = o1;
        __syncthreads();
        if (warp == 0) {
            float mm = -INFINITY;
#pragma unroll
            for (int w = 0; w < 4; ++w) mm = fmaxf(mm, s_m[w]);
            float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const float f = s_m[w] == -INFINITY ? 0.f : exp2f(s_m[w] - mm);  // a warp without a key group
                L += f * s_l[w];
                a0 += f * s_o[w][2 * lane];
                a1 += f * s_o[w][2 * lane + 1];
            }
            const float inv = kInvAct / L;  // V planes carry the activation scale
            a0 *= inv;
            a1 *= inv;
            const int64_t off = (row_base + qr) * p.W + h * 64 + 2 * lane;
            if (p.out_f32) *reinterpret_cast<float2*>(p.out_f32 + off) = make_float2(a0, a1);
            if (p.out_hi) {
                uint16_t h0, l0, h1, l1;
                slb_split2_act(a0, p.fmt, h0, l0);
                slb_split2_act(a1, p.fmt, h1, l1);
                *reinterpret_cast<uint32_t*>(p.out_hi + off) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                *reinterpret_cast<uint32_t*>(p.out_lo + off) = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        }
    }
}

template <int NW, int MINB, bool kSplit = false>
int launch_attn_planes(const AttnPlanesParams& p, int64_t B, cudaStream_t st) {
    const size_t smem = (size_t)(p.Tkp > kKeyBlock ? 2 : 1) * 4 * kKeyBlock * kKPad * sizeof(__half);
    if (smem > 48 * 1024)
        SLB_CUDA_OK(cudaFuncSetAttribute(attention_planes_kernel<NW, MINB, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
    dim3 grid((unsigned)(B * p.H), kSplit ? 1u : (unsigned)slb_ceil_div(p.T - p.q_row0, 16 * NW));
    attention_planes_kernel<NW, MINB, kSplit><<<grid, NW * 32, smem, st>>>(p);
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

int slb_attention_ts_tiles(const uint16_t* qkv_planes, int64_t B, int64_t T, int64_t H, float scale, int n_tiles, int tail_keys,
                           int causal, int plane_fmt, float* out_f32, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st);  // attention_ts.cu

extern "C" int slb_attention_planes(const uint16_t* qkv_planes, int64_t B, int64_t T, int64_t H, int64_t dh, float scale,
                                    int causal, int plane_fmt, float* out_f32, uint16_t* out_planes, void* stream) {
    SLB_REQUIRE(B >= 0 && T > 0 && H > 0, SLB_EINVAL, "slb_attention_planes: bad size");
    if (B == 0) return SLB_OK;
    SLB_REQUIRE(qkv_planes && (out_f32 || out_planes), SLB_EINVAL, "slb_attention_planes: null pointer");
    SLB_REQUIRE(dh == kDh, SLB_EUNSUPPORTED, "slb_attention_planes: head_dim must be 64 (got %lld)", (long long)dh);
    SLB_REQUIRE(plane_fmt == SLB_PLANE_F16, SLB_EUNSUPPORTED, "slb_attention_planes: fp16 planes only");
    SLB_REQUIRE(((uintptr_t)qkv_planes % 16) == 0, SLB_EINVAL, "slb_attention_planes: planes must be 16-byte aligned");
    SLB_REQUIRE(B * H <= 0x7FFFFFFF && T <= 64 * 65535, SLB_EUNSUPPORTED, "slb_attention_planes: problem too large");
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
    // Long sequences: full 128-query tiles run on tcgen05 (attention_ts.cu: P as a tensor-memory operand); a tail of fewer
    // than 64 rows (the last row of a 257-token tower) and short / causal sequences stay on the mma.sync kernel.
    static const bool no_tc = [] { const char* e = getenv("SLB_ATTN_TC"); return e && e[0] == '0'; }();
    // Short sequences (ViT-B/32's 50 tokens, the text towers' 77 / 64, causal or not) run on the tcgen05 kernel too: every image
    // gets a slot of 16 / 32 / 64 / 128 tile rows (two 50-token images per tile), the softmax masks block-diagonally. Slots
    // start on multiples of 16 keys, so an image's bits do not depend on its neighbours or its position (the batch-composition
    // invariance the sharded sweep relies on; a first version that packed images back to back did not have it).
    // SLB_ATTN_PACK=0 keeps short sequences on the mma.sync kernel.
    static const bool no_pack = [] { const char* e = getenv("SLB_ATTN_PACK"); return e && e[0] == '0'; }();
    if (!no_pack && !no_tc && T < 128)
        return slb_attention_ts_tiles(qkv_planes, B, T, H, scale, 1, 0, causal, plane_fmt, out_f32, p.out_hi, p.out_lo, st);
    if (!causal && !no_tc && T >= 128) {
        int n_tiles = (int)(T / 128);
        const int64_t rem = T - 128 * (int64_t)n_tiles;
        if (rem >= 64) n_tiles += 1;
        // T = 128 n + 1 (ViT-L/14's 257 tokens): the last key rides in the softmax threads instead of a key block of its own
        // (SLB_ATTN_KEY_TAIL=0: the old arrangement, a third block holding one key)
        static const bool no_tail = [] { const char* e = getenv("SLB_ATTN_KEY_TAIL"); return e && e[0] == '0'; }();
        const int tail_keys = (!no_tail && rem == 1) ? 1 : 0;  // (2..4 leftover keys: a block of their own, as before)
        int rc = slb_attention_ts_tiles(qkv_planes, B, T, H, scale, n_tiles, tail_keys, 0, plane_fmt, out_f32, p.out_hi, p.out_lo, st);
        if (rc != SLB_OK) return rc;
        if ((int64_t)n_tiles * 128 >= T) return SLB_OK;
        p.q_row0 = n_tiles * 128;
        if (T - p.q_row0 <= 4) {  // one to four leftover rows (the last token of a 257-token tower): fp32 SIMT rows
            attention_rows_kernel<<<(unsigned)(B * H), 128, 0, st>>>(p);
            SLB_LAUNCH_OK("attention_rows");
            return SLB_OK;
        }
        if (T - p.q_row0 <= 16) return launch_attn_planes<4, 3, true>(p, B, st);  // a short tail: split the keys
    }
    return launch_attn_planes<4, 3>(p, B, st);  // 64 query rows per CTA, 3 CTAs / SM (168 registers, <= 72 KB smem)
}

// =================================================================================================
// Long sequences (T >= 128): full 128-row query tiles run on tcgen05 with P as a tensor-memory operand — attention_ts.cu.
// SLB_ATTN_TRACE=1 records one CTA's hand-off timeline into host-mapped words (scripts/trace_attention.py).
// =================================================================================================
static unsigned int* g_attn_trace_host = nullptr;
static unsigned int* g_attn_trace_dev = nullptr;

// the trace words of the last tcgen05 attention launch (null unless SLB_ATTN_TRACE=1): [64 + type * 16 + index] SM clocks
extern "C" const unsigned int* slb_attention_trace() { return g_attn_trace_host; }

// host-mapped trace words (allocated on first use, zeroed on every call); returns the device alias or null
unsigned int* slb_attention_trace_buffer() {
    if (!g_attn_trace_host) {
        if (cudaHostAlloc(reinterpret_cast<void**>(&g_attn_trace_host), 256 * 4, cudaHostAllocMapped) != cudaSuccess) return nullptr;
        if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&g_attn_trace_dev), g_attn_trace_host, 0) != cudaSuccess) return nullptr;
    }
    for (int i = 0; i < 256; ++i) g_attn_trace_host[i] = 0;
    return g_attn_trace_dev;
}
