// K4 attention, tcgen05 with P as a TENSOR-MEMORY operand (TS-mode MMA) — full 128-row query tiles of the long-sequence
// towers (T >= 128: ViT-B/16, ViT-L/14, SigLIP; reference foundation_models/clip.py:118 -> open_clip VisionTransformer ->
// nn.MultiheadAttention). Successor of attention_tc_kernel (attention_mma.cu), which wrote P into shared memory.
//
// One CTA = one (image, head, 128-query tile); 320 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 softmax
// (thread = query row = TMEM lane; two warps per lane quarter, each on 64 of a block's 128 keys). Q and the K / V blocks
// (128 tokens x 64 dims x {hi, lo} planes = 32 KB) arrive through one 3-D tensor map over the in_proj GEMM's split
// planes; K and V travel through a ring of two slots. 97 KB of shared memory and 256 TMEM columns per CTA: two CTAs share
// an SM and fill each other's hand-off bubbles.
//
// What changed against the shared-memory-P kernel, and why (its ncu capture: tensor pipe 16 %, the SM's shared-memory
// port was the limiter — an SS-mode 128x64x16 MMA reads 6 KB of operands for 32 clocks of math):
//   * key blocks of 128: S = Q K^T runs as N = 128 MMAs (8 KB of operands per 64 clocks = the port's rate);
//   * P never touches shared memory: the softmax threads convert exp2(S - max) to split fp16 planes in registers and
//     write them with tcgen05.st INTO THE COLUMNS S OCCUPIED (S fp32 128 columns -> P hi | lo, 16 + 16 columns per 32
//     keys); O += P V is a TS-mode MMA (A = P from TMEM, B = V from its row-major tile as an MN-major operand: 2 KB of
//     shared-memory reads per MMA instead of 6) — no P store, no proxy fence, no swizzle arithmetic;
//   * one S accumulator instead of main + corr: the two cross products (hi.lo, lo.hi) are issued FIRST, so only the four
//     hi.hi MMAs accumulate onto a large value (the tensor core truncates at every accumulate; same error as before),
//     and the softmax threads read 512 bytes per row and block instead of 1024;
//   * the hi plane of P is the fp32 value with its low 13 mantissa bits cleared (exact in fp16), lo the exact
//     remainder rounded to fp16: one AND + one subtract per element instead of two conversions back to fp32.
// Two sweeps over the keys as before (no rescaling of O): sweep 1 takes row maxima from hi.hi logits alone (the maximum
// is only the softmax's stabiliser), double-buffered over the two halves of TMEM (the O columns are idle until sweep 2);
// sweep 2 recomputes S with all three products. S(j+1) is issued right behind P V(j): MMAs of one thread execute in
// issue order, so S(j+1) may overwrite the columns P(j) was read from.
// Accumulators (TMEM columns): [0,128) S / P, [128,192) O main (P hi . V hi), [192,256) O corr (cross terms).
#include "tc_common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace {

constexpr int kSoftmaxWarps = 8;
constexpr int kThreads = (2 + kSoftmaxWarps) * 32;
constexpr int kTile = 128;                       // queries per CTA
constexpr int kKeys = 128;                       // keys per block
constexpr int kPlane = kTile * 64 * 2;           // one 128 x 64 fp16 plane tile: 16 KB
constexpr int kOffQ = 0;                         // Q hi | lo           32 KB
constexpr int kOffKV = 2 * kPlane;               // two slots of hi | lo  64 KB
constexpr int kSlots = 2;
constexpr int kOffStage = kOffKV + kSlots * 2 * kPlane;  // 2 KB per softmax warp: transpose of the output tile
constexpr int kStageWarp = 2048;
constexpr int kOffXch = kOffStage + kSoftmaxWarps * kStageWarp;  // [128] floats: row max / row sum exchange of the column halves
constexpr int kOffBars = kOffXch + kTile * 4;
static_assert(2 * (kOffBars + 128 + 1024) <= 228 * 1024, "two CTAs must fit an SM");
constexpr size_t kSmem = (size_t)kOffBars + 128;
constexpr float kLog2PScale = 10.0f;             // P planes carry 2^10 (hi + lo = 1024 p)

struct AttnTsParams {
    int T, H, W;   // T = keys of a work item: the sequence length, or 128 in packed mode
    // Packed mode (short sequences, Tseq < 128): a work item is (group of `pack` consecutive images, head). Every image gets a
    // SLOT of 2^slot_shift >= Tseq rows of the tile (16, 32, 64 or 128; pack = 128 / slot), loaded by its own TMA boxes; the tile
    // is ONE query tile and ONE key block and the softmax masks the logits block-diagonally (a query sees the keys of its own
    // image only; `causal`: and no later token). Slots start on multiples of 16 keys, so an image's keys fall into the same MMA
    // k-steps and the same summation order wherever it sits in the tile: results are bit-identical for any batch composition.
    // pack = 0: one image per item.
    int pack, slot_shift, Tseq, B, causal;
    // Key tail (T = 128 n + 1: the class token on top of a power-of-two patch grid, ViT-L/14's 257): the last key (tail = 1) does
    // not get a key block of their own (a block costs the same hand-offs whether it holds 1 key or 128: 91 -> 136 us per layer)
    // — the softmax threads take them in fp32 SIMT: logit from the Q row in shared memory, probability into the row sum, p v
    // added to O in the epilogue. T then counts the keys that go through the tensor cores; rows_seq is the real length.
    int tail, rows_seq;
    const uint16_t* qkv; int64_t plane_stride;  // the qkv planes (hi at qkv, lo at qkv + plane_stride) for the tail keys' K / V rows
    float scale_log2;  // scale * log2(e)
    float* out_f32; uint16_t* out_hi; uint16_t* out_lo; int fmt;
    int n_tiles, n_items;  // query tiles per (image, head); work items = B * H * n_tiles
    unsigned int* dbg;     // SLB_ATTN_TRACE=1: host-mapped words; CTA SLB_ATTN_TRACE_CTA records the hand-off timeline of
    int dbg_cta, dbg_item; // its SLB_ATTN_TRACE_ITEM-th work item
};

// [64 + type * 16 + index] = SM clocks since the CTA's start (scripts/trace_attention.py)
#define TS_TRACE(type, idx)                                                                                 \
    do {                                                                                                    \
        if (trace_on && (idx) < 16) p.dbg[64 + (type) * 16 + (idx)] = (unsigned int)(clock64() - t_start);  \
    } while (0)

__device__ __forceinline__ void ts_wait(uint64_t* bar, uint32_t parity) {
    if (slb_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!slb_mbar_try_wait(bar, parity))
        if (clock64() - t0 > 4000000000ll) __trap();  // a lost hand-off must fail the launch, not hang the GPU
}

// MN-major operand tile (rows = K index, 128 B per row = 64 contiguous MN elements), 128-byte swizzle:
// SBO = 1024 B between 8-row groups along K; LBO = bytes between 64-element MN atoms (unused for N = 64).
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// D[tmem] (+)= A[tmem] * B[smem]: A = 128 lanes x 8 columns of packed 16-bit pairs (K = 16)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// 32 lanes x 16 consecutive columns: thread i of the warp writes row (lane base + i), columns [col, col + 16)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ask the TMA unit to bring a box into L2 (no shared-memory destination): hides the DRAM latency of the next work item
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kThreads, 2)
attention_ts_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_pack, AttnTsParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
    uint64_t* q_full = bars;        // TMA -> MMA
    uint64_t* q_empty = bars + 1;   // MMA -> TMA: the work item's last S MMAs have read Q
    uint64_t* kv_full = bars + 2;   // [2] TMA -> MMA
    uint64_t* kv_empty = bars + 4;  // [2] MMA -> TMA
    uint64_t* s_full = bars + 6;    // [2] MMA -> softmax, one per S buffer
    uint64_t* s_free = bars + 8;    // [2] softmax (8 warps) -> MMA: S is in registers
    uint64_t* p_full = bars + 10;   // softmax (8 warps) -> MMA: P is in tensor memory
    uint64_t* o_full = bars + 11;   // MMA -> softmax
    uint64_t* o_free = bars + 12;   // softmax (8 warps) -> MMA: O is in registers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nblk = (p.T + kKeys - 1) / kKeys;
    // rows of the qkv matrix per item index b (an image, or a packed group of images)
    const int item_rows = p.pack ? p.pack * p.Tseq : (p.tail ? p.rows_seq : p.T);
    const int n_iter = 2 * nblk;
    // S buffer of iteration `it` of a work item (sweep 1 alternates the two buffers, sweep 2 uses buffer 0) and how many
    // earlier iterations OF THE ITEM used that buffer; a work item uses buffer 0 uses0 times and buffer 1 uses1 times
    auto s_buf = [&](int it) { return it < nblk ? (it & 1) : 0; };
    auto s_idx = [&](int it) { return it < nblk ? (it >> 1) : ((nblk + 1) >> 1) + (it - nblk); };
    const int uses0 = ((nblk + 1) >> 1) + nblk, uses1 = nblk >> 1;
    auto keys_of = [&](int blk) { return min(kKeys, (p.T - blk * kKeys + 15) & ~15); };  // padded to the MMA N step

    if (warp == 0 && lane == 0) {
        slb_prefetch_tmap(&tm);
        if ((slb_smem_u32(smem) & 1023u) != 0) __trap();
        slb_mbar_init(q_full, 1);
        slb_mbar_init(q_empty, 1);
        for (int i = 0; i < kSlots; ++i) {
            slb_mbar_init(&kv_full[i], 1);
            slb_mbar_init(&kv_empty[i], 1);
            slb_mbar_init(&s_full[i], 1);
            slb_mbar_init(&s_free[i], kSoftmaxWarps);
        }
        slb_mbar_init(p_full, kSoftmaxWarps);
        slb_mbar_init(o_full, 1);
        slb_mbar_init(o_free, kSoftmaxWarps);
        slb_fence_mbar_init();
    }
    if (warp == 1) slb_tmem_alloc<256>(tmem_slot);
    slb_tc_fence_before();
    __syncthreads();
    slb_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t t_o = tmem_base + 128;
    const bool trace_cta = p.dbg != nullptr && (int)blockIdx.x == p.dbg_cta;
    const long long t_start = clock64();
    if (trace_cta && warp == 2 && lane == 0) { unsigned int smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); p.dbg[63] = smid; }

    // Persistent: CTA c works on items c, c + gridDim.x, ... (item = (image, head, query tile)); every barrier keeps
    // running phases across items, so the producer warp prefetches the next item's Q and first K blocks while the softmax
    // warps are still in the current item's last block and epilogue.
    if (warp == 0) {
        if (lane == 0) {
            int t = 0;  // K / V ring item counter (running)
            int n = 0;  // work items done by this CTA
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                const bool trace_on = trace_cta && n == p.dbg_item;
                const int bh = item / p.n_tiles, tile = item % p.n_tiles;
                const int b = bh / p.H, h = bh % p.H;
                const int row_base = b * item_rows;
                ts_wait(q_empty, (uint32_t)(n & 1) ^ 1u);
                // one 128-row box of both planes; packed mode: ONE 4-D box {64 columns, slot tokens, pack images, 2 planes} over the
                // planes seen as (plane, image, token, column) — it lands as [plane][image][token] = the slot layout, tokens past
                // the sequence and images past the batch zero-filled by the TMA unit
                auto load_tile = [&](unsigned char* dst, int col, int row, uint64_t* bar) {
                    if (!p.pack)
                        slb_tma_load_3d(dst, &tm, col, row, 0, bar);
                    else
                        slb_tma_load_4d(dst, &tm_pack, col, 0, b * p.pack, 0, bar);
                };
                slb_mbar_arrive_expect_tx(q_full, 2u * kPlane);
                load_tile(smem + kOffQ, h * 64, row_base + tile * kTile, q_full);
                {   // the ring only holds two blocks: start the NEXT item's operands on their way to L2 now
                    const int nxt = item + (int)gridDim.x;
                    if (nxt < p.n_items) {
                        const int bh2 = nxt / p.n_tiles, b2 = bh2 / p.H, h2 = bh2 % p.H, rb2 = b2 * item_rows;
                        tma_prefetch_l2_3d(&tm, h2 * 64, rb2 + (nxt % p.n_tiles) * kTile, 0);
                        for (int blk = 0; blk < nblk; ++blk) {
                            tma_prefetch_l2_3d(&tm, p.W + h2 * 64, rb2 + blk * kKeys, 0);
                            tma_prefetch_l2_3d(&tm, 2 * p.W + h2 * 64, rb2 + blk * kKeys, 0);
                        }
                    }
                }
                // ring items in consumption order: K_0 .. K_{n-1} (sweep 1), then — sweep 2 walks the blocks BACKWARDS, so
                // the last K block of sweep 1 is reused in place — V_{n-1}, K_{n-2}, V_{n-2}, ..., K_0, V_0; item t reuses
                // the slot of item t - 2 (a K slot is released by its S MMAs: early; a V slot by its P V MMAs).
                int t_item = 0;
                auto load_item = [&](int col, int row) {
                    const int slot = t % kSlots;
                    ts_wait(&kv_empty[slot], (uint32_t)((t / kSlots) & 1) ^ 1u);
                    slb_mbar_arrive_expect_tx(&kv_full[slot], 2u * kPlane);
                    load_tile(smem + kOffKV + slot * 2 * kPlane, col, row, &kv_full[slot]);
                    TS_TRACE(0, t_item);
                    ++t;
                    ++t_item;
                };
                for (int blk = 0; blk < nblk; ++blk) load_item(p.W + h * 64, row_base + blk * kKeys);  // sweep 1: K
                for (int j = 0; j < nblk; ++j) {  // sweep 2, blocks in reverse order: K_{n-1} is still resident
                    const int blk = nblk - 1 - j;
                    if (j > 0) load_item(p.W + h * 64, row_base + blk * kKeys);
                    load_item(2 * p.W + h * 64, row_base + blk * kKeys);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t qa = slb_smem_u32(smem + kOffQ);
            const uint64_t dq_hi = slb_umma_desc_sw128(qa), dq_lo = slb_umma_desc_sw128(qa + kPlane);
            const uint64_t dk0 = slb_umma_desc_sw128(slb_smem_u32(smem + kOffKV));
            const uint32_t idesc_o = slb_umma_idesc_f16(0, kTile, 64) | (1u << 16);    // B (= V) is MN-major
            const uint32_t idesc_o2 = slb_umma_idesc_f16(0, kTile, 128) | (1u << 16);  // B = [V hi | V lo]
            int t0 = 0;                 // ring items consumed by earlier work items
            int base[2] = {0, 0};       // S-buffer uses of earlier work items
            int pblk = 0;               // sweep-2 blocks of earlier work items
            int n = 0;
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
                const bool trace_on = trace_cta && n == p.dbg_item;
                // ring item holding the K block of iteration `it`; sweep 2 (j = it - nblk) reuses item nblk - 1 for j = 0
                auto item_k = [&](int it) { return t0 + (it < nblk ? it : (it == nblk ? nblk - 1 : nblk + 2 * (it - nblk) - 1)); };
                auto block_of = [&](int it) { return it < nblk ? it : n_iter - 1 - it; };
                auto issue_s = [&](int it) {
                    const int blk = block_of(it);
                    const bool sweep2 = it >= nblk;
                    const int nk = keys_of(blk);
                    const int t = item_k(it), slot = t % kSlots;
                    const int buf = s_buf(it), idx = base[buf] + s_idx(it);
                    ts_wait(&kv_full[slot], (uint32_t)((t / kSlots) & 1));
                    TS_TRACE(1, it);  // K landed
                    if (idx > 0) ts_wait(&s_free[buf], (uint32_t)((idx - 1) & 1));  // the buffer's previous S is in registers
                    slb_tc_fence_after();
                    TS_TRACE(2, it);  // S issue
                    const uint64_t dk_hi = dk0 + (uint64_t)(slot * (2 * kPlane >> 4)), dk_lo = dk_hi + (kPlane >> 4);
                    const uint32_t idesc_s = slb_umma_idesc_f16(0, kTile, nk);
                    const uint32_t d = tmem_base + (uint32_t)(buf * 128);
                    if (sweep2) {
                        // cross terms first: only the four hi.hi MMAs accumulate onto a large value
#pragma unroll
                        for (int k = 0; k < 4; ++k) slb_umma_f16(d, dq_hi + 2 * k, dk_lo + 2 * k, idesc_s, k != 0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) slb_umma_f16(d, dq_lo + 2 * k, dk_hi + 2 * k, idesc_s, true);
#pragma unroll
                        for (int k = 0; k < 4; ++k) slb_umma_f16(d, dq_hi + 2 * k, dk_hi + 2 * k, idesc_s, true);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) slb_umma_f16(d, dq_hi + 2 * k, dk_hi + 2 * k, idesc_s, k != 0);
                    }
                    slb_umma_commit(&s_full[buf]);
                    if (it != nblk - 1) slb_umma_commit(&kv_empty[slot]);  // sweep 1's last K block is read again by sweep 2
                    if (it == n_iter - 1) slb_umma_commit(q_empty);        // the item's last read of Q
                };
                auto issue_pv = [&](int it) {
                    const int j = it - nblk, blk = block_of(it);
                    const int nk = keys_of(blk);
                    const int t = t0 + nblk + 2 * j, slot = t % kSlots;
                    ts_wait(&kv_full[slot], (uint32_t)((t / kSlots) & 1));
                    TS_TRACE(8, j);  // V landed
                    ts_wait(p_full, (uint32_t)((pblk + j) & 1));
                    slb_tc_fence_after();
                    TS_TRACE(3, j);  // P V issue
                    const uint32_t va = slb_smem_u32(smem + kOffKV) + (uint32_t)(slot * 2 * kPlane);
                    const uint64_t dv_hi = desc_mn_sw128(va, 1024), dv_both = desc_mn_sw128(va, kPlane);
                    const int ksteps = nk >> 4;
                    // k-step s = keys [16 s, 16 s + 16): P hi at column 32 (s / 2) + 8 (s % 2), P lo 16 columns further.
                    // Two MMAs per k-step: P hi . [V hi | V lo] (one N = 128 operand: the lo plane is the second MN atom,
                    // kPlane bytes after the hi plane) -> [O main | O corr], then P lo . V hi -> O corr. (An N = 64 MMA takes as
                    // long as an N = 128 one when it is issued from a single stream: scripts/micro/umma_rates.cu.)
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        if (s < ksteps) {
                            const uint32_t a_hi = tmem_base + (uint32_t)(32 * (s >> 1) + 8 * (s & 1)), a_lo = a_hi + 16;
                            const uint64_t koff = (uint64_t)((2048 >> 4) * s);
                            umma_f16_ts(t_o, a_hi, dv_both + koff, idesc_o2, (j | s) != 0);
                            umma_f16_ts(t_o + 64, a_lo, dv_hi + koff, idesc_o, true);
                        }
                    }
                    slb_umma_commit(&kv_empty[slot]);
                };
                ts_wait(q_full, (uint32_t)(n & 1));
                // the previous item's O (columns [128, 256) = S buffer 1) must be in the softmax warps' registers
                if (n > 0) ts_wait(o_free, (uint32_t)((n - 1) & 1));
                // sweep 1 runs one block ahead of the softmax warps (two S buffers); sweep 2: S(j), P V(j), S(j+1), ...
                for (int it = 0; it <= nblk; ++it) issue_s(it);
                for (int it = nblk; it < n_iter; ++it) {
                    issue_pv(it);
                    if (it + 1 < n_iter) issue_s(it + 1);
                }
                slb_umma_commit(o_full);
                t0 += 3 * nblk - 1;
                base[0] += uses0;
                base[1] += uses1;
                pblk += nblk;
            }
        }
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;            // which 64 keys of every 128-key block
        const int r = quarter * 32 + lane;           // row of the tile = TMEM lane
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        constexpr float kInvAct = 1.0f / SLB_ACT_PLANE_SCALE;
        const float c_main = p.scale_log2 * kInvAct * kInvAct;
        // 2 KB of staging per warp; the two warps of a lane quarter (column halves 0 and 1) exchange row maxima / sums through
        // one word per row: half 1 writes, barrier, half 0 combines and writes back, barrier, half 1 reads
        unsigned char* stage_own = smem + kOffStage + (half * 4 + quarter) * kStageWarp;
        float* xw = reinterpret_cast<float*>(smem + kOffXch) + r;
        const int pair_bar = 1 + quarter;  // named barrier of the quarter's two warps
        int base[2] = {0, 0};
        int pblk = 0;
        int n = 0;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++n) {
            const bool trace_on = trace_cta && n == p.dbg_item && warp == 2 && lane == 0;
            const int bh = item / p.n_tiles, tile = item % p.n_tiles;
            const int b = bh / p.H, h = bh % p.H;
            const int row_base = b * item_rows;
            // keys this thread's query row may see: [k_lo, k_lo + k_span); rows of the item that exist: rows_item
            int k_lo = 0;
            int k_span = p.pack ? p.Tseq : p.T;
            if (p.pack) k_lo = (r >> p.slot_shift) << p.slot_shift;
            if (p.causal) k_span = min(k_span, r - k_lo + 1);  // causal: keys up to the query's own position
            if (p.tail && warp == 2 && lane < 4) {
                // the tail keys' K and V rows (hi / lo: four 128-byte lines per key) into L1 now: the softmax threads read them
                // between the sweeps and in the epilogue, where a trip to L2 would sit on every warp's critical path
                const int which = lane & 3;
                const uint16_t* a = p.qkv + ((int64_t)row_base + p.T) * (3 * p.W) + ((which & 1) ? 2 : 1) * p.W + h * 64 +
                                    ((which & 2) ? p.plane_stride : 0);
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
            }
            float m_row = -INFINITY, l_row = 0.f, neg_m = 0.f;
            float s_tail = 0.f;
            for (int it = 0; it < n_iter; ++it) {
                const int blk = it < nblk ? it : n_iter - 1 - it;  // sweep 2 walks the blocks backwards
                const bool sweep2 = it >= nblk;
                const int nk = keys_of(blk);
                const int buf = s_buf(it), idx = base[buf] + s_idx(it);
                const uint32_t t_s = tmem_base + lane_addr + (uint32_t)(buf * 128);
                const bool full_block = !p.pack && (blk + 1) * kKeys <= p.T;
                if (it == nblk) {
                    if (p.tail) {
                        // the tail key's logit for this thread's query row, fp32: q from the Q tile (both planes, 128-byte
                        // swizzle: 16-byte chunk c of row r sits at chunk c ^ (r & 7)), k from the planes in global memory.
                        // (Rolled loop, one key: the first version unrolled four keys and grew the kernel from 79 to 126 KB of
                        // code — ncu then showed 23 % of the samples waiting for instructions.)
                        ts_wait(q_full, (uint32_t)(n & 1));  // (long complete: makes the TMA's writes visible to this thread)
                        const unsigned char* qrow = smem + kOffQ + r * 128;
                        const uint16_t* kh = p.qkv + ((int64_t)row_base + p.T) * (3 * p.W) + p.W + h * 64;
                        float dot = 0.f;
#pragma unroll 1
                        for (int ch = 0; ch < 8; ++ch) {
                            const uint4 qh = *reinterpret_cast<const uint4*>(qrow + ((ch ^ (r & 7)) << 4));
                            const uint4 ql = *reinterpret_cast<const uint4*>(qrow + kPlane + ((ch ^ (r & 7)) << 4));
                            const uint4 kh4 = __ldg(reinterpret_cast<const uint4*>(kh + ch * 8));
                            const uint4 kl4 = __ldg(reinterpret_cast<const uint4*>(kh + p.plane_stride + ch * 8));
                            const uint32_t qhw[4] = {qh.x, qh.y, qh.z, qh.w}, qlw[4] = {ql.x, ql.y, ql.z, ql.w};
                            const uint32_t khw[4] = {kh4.x, kh4.y, kh4.z, kh4.w}, klw[4] = {kl4.x, kl4.y, kl4.z, kl4.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 qa = __half22float2(*reinterpret_cast<const __half2*>(&qhw[e]));
                                const float2 qb = __half22float2(*reinterpret_cast<const __half2*>(&qlw[e]));
                                const float2 ka = __half22float2(*reinterpret_cast<const __half2*>(&khw[e]));
                                const float2 kb = __half22float2(*reinterpret_cast<const __half2*>(&klw[e]));
                                dot = fmaf(qa.x + qb.x, ka.x + kb.x, dot);
                                dot = fmaf(qa.y + qb.y, ka.y + kb.y, dot);
                            }
                        }
                        s_tail = dot * c_main;  // planes carry the activation scale on both sides, like S
                        m_row = fmaxf(m_row, s_tail);
                    }
                    // end of sweep 1: the two warps of a row combine their partial maxima
                    if (half) *xw = m_row;
                    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                    if (!half) {
                        m_row = fmaxf(m_row, *xw);
                        *xw = m_row;
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                    if (half) m_row = *xw;
                    neg_m = kLog2PScale - m_row;  // exp2(s - m + 10) = 1024 p
                }
                ts_wait(&s_full[buf], (uint32_t)(idx & 1));
                slb_tc_fence_after();
                TS_TRACE(4, it);  // S visible
                if (!sweep2) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c = half * 64 + u * 32;
                        if (c < nk) {  // warp-uniform
                            uint32_t a[32];
                            slb_tmem_ld_32x32(t_s + c, a);
                            slb_tmem_ld_wait();
                            if (full_block) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) m_row = fmaxf(m_row, __uint_as_float(a[j]) * c_main);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if ((unsigned)(blk * kKeys + c + j - k_lo) < (unsigned)k_span) m_row = fmaxf(m_row, __uint_as_float(a[j]) * c_main);
                            }
                        }
                    }
                    slb_tc_fence_before();
                    __syncwarp();
                    if (lane == 0) slb_mbar_arrive(&s_free[buf]);
                    TS_TRACE(5, it);  // maxima done
                    continue;
                }
                // sweep 2: 32 keys at a time, S columns [c, c + 32) -> P hi columns [c, c + 16) | P lo columns [c + 16, c + 32)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c = half * 64 + u * 32;
                    if (c < nk) {  // warp-uniform
                        uint32_t a[32];
                        slb_tmem_ld_32x32(t_s + c, a);
                        slb_tmem_ld_wait();
                        uint32_t hh[16], ll[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float x0 = fmaf(__uint_as_float(a[2 * e]), c_main, neg_m);
                            float x1 = fmaf(__uint_as_float(a[2 * e + 1]), c_main, neg_m);
                            if (!full_block) {
                                if ((unsigned)(blk * kKeys + c + 2 * e - k_lo) >= (unsigned)k_span) x0 = -INFINITY;
                                if ((unsigned)(blk * kKeys + c + 2 * e + 1 - k_lo) >= (unsigned)k_span) x1 = -INFINITY;
                            }
                            const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
                            l_row += p0 + p1;
                            // hi = the value with its low 13 mantissa bits cleared (exact in fp16), lo = the exact remainder
                            const float h0 = __uint_as_float(__float_as_uint(p0) & 0xFFFFE000u);
                            const float h1 = __uint_as_float(__float_as_uint(p1) & 0xFFFFE000u);
                            const __half2 hp = __floats2half2_rn(h0, h1);
                            const __half2 lp = __floats2half2_rn(p0 - h0, p1 - h1);
                            hh[e] = *reinterpret_cast<const uint32_t*>(&hp);
                            ll[e] = *reinterpret_cast<const uint32_t*>(&lp);
                        }
                        tmem_st_32x16(t_s + c, hh);
                        tmem_st_32x16(t_s + c + 16, ll);
                    }
                }
                TS_TRACE(6, it - nblk);  // exponentials + stores issued
                tmem_st_wait();
                slb_tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    slb_mbar_arrive(&s_free[buf]);
                    slb_mbar_arrive(p_full);
                }
                TS_TRACE(7, it - nblk);  // P published
            }
            base[0] += uses0;
            base[1] += uses1;
            pblk += nblk;
            // ---- row sums of the two column halves, then O = (main + corr) / (scales * l); each warp stores 32 of the 64 dims
            ts_wait(o_full, (uint32_t)(n & 1));
            slb_tc_fence_after();
            TS_TRACE(9, 0);  // O complete
            uint32_t a[32], cr[32];
            {
                const int c = half * 32;
                slb_tmem_ld_32x32(t_o + lane_addr + c, a);
                slb_tmem_ld_32x32(t_o + lane_addr + 64 + c, cr);
                slb_tmem_ld_wait();
            }
            // O is in registers: the next item's S MMAs may reuse its columns
            slb_tc_fence_before();
            __syncwarp();
            if (lane == 0) slb_mbar_arrive(o_free);
            if (half) *xw = l_row;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            if (!half) {
                l_row += *xw;
                *xw = l_row;
            }
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
            if (half) l_row = *xw;
            float p_tail = 0.f;
            if (p.tail) {  // (l_row is complete except for the tail key: both warps of the row add the same p, once each)
                p_tail = ex2_approx(s_tail + neg_m);
                l_row += p_tail;
            }
            const float inv = kInvAct / l_row;  // V planes carry the activation scale, l_row the 2^10 of the P planes
            const int c = half * 32;
            // tile row rt (0..127) -> row of the output matrix, or -1 when the row does not exist (past the sequence / the batch)
            auto out_row = [&](int rt) -> int64_t {
                if (!p.pack) return tile * kTile + rt < p.T + p.tail ? (int64_t)row_base + tile * kTile + rt : -1;
                const int im = rt >> p.slot_shift, tok = rt & ((1 << p.slot_shift) - 1);
                const int g_img = b * p.pack + im;
                return (tok < p.Tseq && g_img < p.B) ? (int64_t)g_img * p.Tseq + tok : -1;
            };
            const int64_t col_base = (int64_t)h * 64 + c;
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(a[j]) + __uint_as_float(cr[j]);
            if (p.tail) {
                // O += p v for the tail key: v = hi + lo of this head's 32 dims [c, c + 32) (planes at the activation scale, as the
                // accumulated P V is)
                const uint16_t* vh = p.qkv + ((int64_t)row_base + p.T) * (3 * p.W) + 2 * p.W + h * 64 + c;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const uint4 h4 = __ldg(reinterpret_cast<const uint4*>(vh + q4 * 8));
                    const uint4 l4 = __ldg(reinterpret_cast<const uint4*>(vh + p.plane_stride + q4 * 8));
                    const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                        const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
                        o[q4 * 8 + 2 * e] = fmaf(p_tail, a2.x + b2.x, o[q4 * 8 + 2 * e]);
                        o[q4 * 8 + 2 * e + 1] = fmaf(p_tail, a2.y + b2.y, o[q4 * 8 + 2 * e + 1]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] *= inv;
            const int64_t my_row = out_row(quarter * 32 + lane);
            if (p.out_f32 && my_row >= 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    reinterpret_cast<float4*>(p.out_f32 + my_row * p.W + col_base)[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            if (p.out_hi) {
                // thread = row would store 16 bytes to 32 different lines per instruction (the LSU serialises them: the
                // epilogue was store-bound); the warp transposes through its staging region so that 4 lanes cover the 64
                // bytes a row owns per plane: 8 rows per instruction. 16-byte chunk ch of row rr lives at
                // rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4): conflict-free for both access patterns.
                uint32_t hh[16], ll[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) slb_split_pair_act_f16(o[2 * e], o[2 * e + 1], hh[e], ll[e]);
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
                    __syncwarp();  // the region's previous contents (the other plane) have been read
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const uint4 v = pl == 0 ? make_uint4(hh[4 * ch], hh[4 * ch + 1], hh[4 * ch + 2], hh[4 * ch + 3])
                                                : make_uint4(ll[4 * ch], ll[4 * ch + 1], ll[4 * ch + 2], ll[4 * ch + 3]);
                        *reinterpret_cast<uint4*>(stage_own + lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4)) = v;
                    }
                    __syncwarp();
                    uint16_t* dst = pl == 0 ? p.out_hi : p.out_lo;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = i * 8 + (lane >> 2), ch = lane & 3;
                        const uint4 v = *reinterpret_cast<const uint4*>(stage_own + rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
                        const int64_t orow = out_row(quarter * 32 + rr);
                        if (orow >= 0) *reinterpret_cast<uint4*>(dst + orow * p.W + col_base + ch * 8) = v;
                    }
                }
                __syncwarp();
            }
            TS_TRACE(9, 1);  // item done
        }
    }

    slb_tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        slb_tc_fence_after();
        slb_tmem_dealloc<256>(tmem_base);
    }
}

}  // namespace

unsigned int* slb_attention_trace_buffer();  // attention_mma.cu

// Full 128-row query tiles of every (image, head) on the TS-mode tcgen05 path; the caller handles the remaining rows.
// T < 128 (n_tiles must be 1): packed mode — every image in a slot of 16 / 32 / 64 / 128 rows of the tile, block-diagonal (and,
// with `causal`, lower-triangular) softmax mask, every row is covered.
// tail_keys (0 or 1, only with T = 128 n_tiles + 1): the last key is taken by the softmax threads in fp32 (no key block).
int slb_attention_ts_tiles(const uint16_t* qkv_planes, int64_t B, int64_t T, int64_t H, float scale, int n_tiles, int tail_keys,
                           int causal, int plane_fmt, float* out_f32, uint16_t* out_hi, uint16_t* out_lo, cudaStream_t st) {
    const int64_t W = H * 64, rows = B * T;
    CUtensorMap tm;
    int rc = slb_make_plane_map(&tm, qkv_planes, rows, 3 * W, 2, kTile);
    if (rc != SLB_OK) return rc;
    AttnTsParams p{};
    p.T = (int)T; p.H = (int)H; p.W = (int)W;
    if (tail_keys) {
        SLB_REQUIRE(tail_keys == 1 && T == (int64_t)n_tiles * kTile + 1, SLB_EINVAL,
                    "slb_attention_ts_tiles: the key tail is the one key past the last full block");
        p.tail = tail_keys; p.rows_seq = (int)T; p.T = (int)(T - tail_keys);
        p.qkv = qkv_planes; p.plane_stride = rows * 3 * W;
    }
    int64_t n_batch = B;  // item indices b
    CUtensorMap tm_pack = tm;
    if (T < kTile) {
        SLB_REQUIRE(n_tiles == 1 && tail_keys == 0, SLB_EINVAL, "slb_attention_ts_tiles: a short sequence is one tile");
        int shift = 4;
        while ((1 << shift) < T) ++shift;
        p.slot_shift = shift; p.pack = kTile >> shift; p.Tseq = (int)T; p.B = (int)B; p.causal = causal ? 1 : 0;
        p.T = kTile;
        n_batch = (B + p.pack - 1) / p.pack;
        // the planes as (plane, image, token, column); box = {64 columns, slot tokens, pack images, both planes}
        slb_tmap_encode_fn enc = slb_get_tmap_encoder();
        if (!enc) return SLB_ECUDA;
        cuuint64_t dims[4] = {(cuuint64_t)(3 * W), (cuuint64_t)T, (cuuint64_t)B, 2};
        cuuint64_t strides[3] = {(cuuint64_t)(3 * W) * 2, (cuuint64_t)T * (cuuint64_t)(3 * W) * 2, (cuuint64_t)rows * (cuuint64_t)(3 * W) * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(1 << shift), (cuuint32_t)p.pack, 2};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm_pack, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<uint16_t*>(qkv_planes), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            slb_set_error("cuTensorMapEncodeTiled (packed attention map) failed with %d (B=%lld T=%lld W=%lld)", (int)r, (long long)B,
                          (long long)T, (long long)W);
            return SLB_ECUDA;
        }
    } else {
        SLB_REQUIRE(!causal, SLB_EUNSUPPORTED, "slb_attention_ts_tiles: causal masks on short sequences only");
    }
    p.scale_log2 = scale * 1.4426950408889634f;
    p.out_f32 = out_f32; p.out_hi = out_hi; p.out_lo = out_lo; p.fmt = plane_fmt;
    static const bool trace = [] { const char* e = getenv("SLB_ATTN_TRACE"); return e && e[0] == '1'; }();
    if (trace) {
        const char* e = getenv("SLB_ATTN_TRACE_CTA");
        p.dbg = slb_attention_trace_buffer();
        p.dbg_cta = e ? atoi(e) : 0;
        e = getenv("SLB_ATTN_TRACE_ITEM");
        p.dbg_item = e ? atoi(e) : 0;
    }
    SLB_CUDA_OK(cudaFuncSetAttribute(attention_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    p.n_tiles = n_tiles;
    p.n_items = (int)(n_batch * H * n_tiles);
    int dev = 0, sms = 0;
    SLB_CUDA_OK(cudaGetDevice(&dev));
    SLB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = std::min(p.n_items, 2 * sms);  // persistent: two CTAs per SM
    attention_ts_kernel<<<grid, kThreads, kSmem, st>>>(tm, tm_pack, p);
    SLB_LAUNCH_OK("attention_ts");
    return SLB_OK;
}
