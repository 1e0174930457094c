// K8 — polysemanticity_score (reference semanticlens/scores.py:132-185): per neuron, sklearn's
// KMeans(n_clusters=2, n_init=10, random_state=123) on its k example embeddings, then 1 - cos(centre_1, centre_2),
// with the small-cluster fallback of scores.py:173-184. The reference runs a serial Python loop over neurons
// (13-53 ms each on 8 cores); here one CTA owns a neuron and the whole fit runs on-chip from the neuron's Gram matrix.
//
// Why the Gram matrix: every centre sklearn forms is the mean of a subset A of the points, so with the centred Gram
// Gc = G0 - r 1^T - 1 r^T + m (r = row means of G0 = X X^T, m = grand mean)
//     x_i . c_A = (1/|A|) sum_{j in A} Gc_ij          |c_A|^2 = (1/|A|) sum_{i in A} x_i . c_A
// and distances, centre shifts, inertia, the tolerance (trace(Gc) / (k D) * 1e-4) and the final cosine of the
// un-centred centres are all functions of G0: the k x D examples are read ONCE (HBM-bound), Lloyd iterations cost
// O(k * moved points) instead of O(k D). The rows of Gc sum to zero, so the second cluster's sums are minus the first's.
//
// Phase A  G0 = X X^T in float64 (products of fp32 values are exact in f64, so assignments match sklearn's f64 fit
//          except at true f64 near-ties) on the FP64 tensor cores: mma.sync.m8n8k4.f64 (DMMA.884). The 36 32x32 tiles
//          of the upper triangle go to the CTA's 9 warps in 4 passes (one tile = 32 f64 accumulators per lane); the
//          examples arrive as raw fp32 rows through a 2-stage ring of 1-D bulk async copies (one 128-byte row piece per
//          thread and stage, evict-first in L2 so the Gram slots stay resident) and are widened to f64 when a fragment
//          is loaded (row pitch 36 floats: conflict-free). Because A and B^T are the same matrix, both fragments
//          are the same access pattern. Results go to the CTA's own k*k slot of the workspace, mirrored.
// Phase B  thread i = example i: k-means++ with sklearn's RandomState stream (data independent: the first-centre
//          index and the two local-trial uniforms per init are precomputed on the host), Lloyd to strict convergence or
//          tolerance, empty-cluster relocation, best of n_init by inertia with sklearn's _is_same_clustering guard.
//          oracle/polysem.py::kmeans2_gram is the line-by-line CPU statement of this phase.
#include "slb_common.cuh"

#include <stdlib.h>
#include <algorithm>

namespace {

constexpr int kT = 256;     // max examples per neuron (thread i < kT owns example i in phase B)
constexpr int kWarps = 9;   // 36 upper-triangle tiles of 32x32 = 4 passes x 9 warps
constexpr int kThreads = kWarps * 32;
constexpr int kStageCols = 32;  // features per pipeline stage
constexpr int kPitch = 36;      // floats per staged row: bank = (4*row + col) mod 32 -> conflict-free fragment loads
constexpr int kStages = 2;
constexpr int kMaxInit = 10;  // sklearn's n_init of the reference call (scores.py:167)
constexpr int kMaxIter = 300;  // sklearn default max_iter

struct PolyParams {
    const float* V;
    int64_t C;
    int k, D;
    int n_init;
    int replace_empty;
    int first[kMaxInit];
    double rand[2 * kMaxInit];
    double* G;  // workspace: gridDim.x slots of k(k+1)/2 doubles (packed upper triangle)
    double* out;
};

struct Smem {
    float stage[kStages][kT][kPitch];  // fp32 examples, kStageCols features per stage
    uint64_t full[kStages];
    double tA0[kMaxInit][kT];      // first-iteration sums of every restart
    double r[kThreads];
    double diag[kThreads];
    double red[kWarps * 4];
    double scan[kWarps];
    int ired[kWarps * 2];
    short moved[kT];   // examples that changed side this iteration, ascending; ~j encodes "left cluster 0"
    unsigned short in0[kThreads];  // bit `it`: the example starts restart `it` in cluster 0
    int fb[kT / 8 + 1];            // first fragment of every 8-row band of the Gram slot
    double red10[kWarps][kMaxInit];
    int seed0[kMaxInit];           // second seed of every restart (the first is PolyParams::first)
    int cnt0[kMaxInit];            // size of cluster 0 after the first assignment
    int wcnt[kWarps];
    unsigned char best_lab[kThreads];
    unsigned char best_mask[kThreads];
};

// ---- block primitives (kThreads threads, every thread must call) ----------------------------------------
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], Smem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < N; ++n) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[n] += __shfl_xor_sync(0xffffffffu, v[n], o);
    }
    __syncthreads();  // previous users of sm.red are done
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) sm.red[warp * 4 + n] = v[n];
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += sm.red[w * 4 + n];
        v[n] = t;
    }
}

// N-wide version (N <= kMaxInit): the restarts' sums behind one pair of barriers
template <int N>
__device__ __forceinline__ void block_sumN(double (&v)[N], Smem& sm) {
    static_assert(N <= kMaxInit, "red10 holds kMaxInit values per warp");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < N; ++n) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[n] += __shfl_xor_sync(0xffffffffu, v[n], o);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) sm.red10[warp][n] = v[n];
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += sm.red10[w][n];
        v[n] = t;
    }
}

__device__ __forceinline__ double block_sum1(double x, Smem& sm) {
    double v[1] = {x};
    block_sum<1>(v, sm);
    return v[0];
}

template <int N>
__device__ __forceinline__ void block_count(int (&v)[N], Smem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < N; ++n) v[n] = __reduce_add_sync(0xffffffffu, v[n]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int n = 0; n < N; ++n) sm.ired[warp * 2 + n] = v[n];
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += sm.ired[w * 2 + n];
        v[n] = t;
    }
}

// inclusive prefix sum over the threads (np.cumsum)
__device__ __forceinline__ double block_scan(double x, Smem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) sm.scan[warp] = x;
    __syncthreads();
    double base = 0.0;
    for (int w = 0; w < warp; ++w) base += sm.scan[w];
    return base + x;
}

// (max value, first index attaining it) over threads
__device__ __forceinline__ void block_argmax(double& val, int& idx, Smem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, val, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (v2 > val || (v2 == val && i2 < idx)) { val = v2; idx = i2; }
    }
    __syncthreads();
    if (lane == 0) { sm.red[warp * 4] = val; sm.ired[warp * 2] = idx; }
    __syncthreads();
    val = sm.red[0];
    idx = sm.ired[0];
    for (int w = 1; w < kWarps; ++w) {
        const double v2 = sm.red[w * 4];
        const int i2 = sm.ired[w * 2];
        if (v2 > val || (v2 == val && i2 < idx)) { val = v2; idx = i2; }
    }
}

// ---- Gram storage: upper triangle packed as DMMA fragments --------------------------------------------------
// G0 is symmetric. A slot keeps, for every band R of 8 rows and every group Kq >= 2R of 4 columns, the 8 x 4 block
// G0[8R + g][4Kq + t] as 32 doubles in lane order g * 4 + t — exactly one A fragment of mma.m8n8k4.f64 — bands back to
// back: 1056 fragments = 270 KB at k = 256. Two CTAs per SM keep 296 slots live: 80 MB, inside the 126 MB L2.
// Every full pass over the matrix in phase B is a product G0 * W with a thin weight matrix W (ones -> row means; the
// ten restarts' first memberships; the final membership) and runs on the FP64 tensor cores straight from this layout:
// an A fragment on or above the diagonal is one 256-byte contiguous load for the warp; a fragment below the diagonal
// is the transpose of half an 8 x 8 block stored above it, which in lane order is two contiguous 128-byte pieces. The
// earlier layout (plain packed rows, thread i sweeping "column i" with scalar loads) touched 32 different lines per
// warp instruction on the half of the sweep that ran along a thread's own row: ncu showed 18.5 sectors per request and the
// sweeps bound by L1 tag throughput (35-40 % of the kernel's samples).
__host__ __device__ __forceinline__ int frag_base(int R, int KQ) { return R * KQ - R * (R - 1); }  // fragments before band R
__host__ __device__ __forceinline__ int64_t slot_doubles(int64_t k) {
    const int NB = (int)((k + 7) >> 3), KQ = (int)((k + 3) >> 2);
    return (int64_t)frag_base(NB, KQ) * 32;
}
// element (i, j) of the symmetric matrix; fb = the band offsets (in fragments) in shared memory
__device__ __forceinline__ int frag_elem(const int* __restrict__ fb, int i, int j) {
    const int lo = min(i, j), hi = max(i, j), R = lo >> 3;
    return ((fb[R] + (hi >> 2) - 2 * R) << 5) + ((lo & 7) << 2) + (hi & 3);
}

// ---- phase A: G0 = X X^T (float64, DMMA) -----------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// (Measured and dropped: TWO kernels per chunk of 8192 neurons — a Gram kernel whose two resident CTAs per SM are always in
// phase A and a k-means kernel among its own kind, the Gram slots making one round trip through HBM. Phase A then takes
// 346 us per neuron and CTA = 173 us per neuron and SM = 31.6 TFLOP/s of FP64 tensor work, 79 % of B200's 40 TFLOP/s, and
// phase B 143 us per neuron and CTA (248 us at three CTAs per SM: the G0 x W products contend for the same pipe): 130 ms for
// 65 536 neurons, the same as the fused kernel, for 2.2 GB of workspace instead of 80 MB. So the fused kernel's overlap
// already hides phase B about as well as a split can; what is left is the FP64 tensor throughput itself.)
// (Measured and dropped: ONE 2-D TMA box per stage behind a 4-slot full / empty mbarrier ring instead of a bulk copy per row
// and a block barrier per stage: phase A of a lone CTA 238 -> 234 us, the whole kernel 130 -> 136 ms — the copies and
// barriers were not the bound. With SLB_POLYSEM_CTAS_PER_SM=1 a lone CTA needs 238 us per Gram matrix where the DMMA
// sub-pipe would allow ~150: one warp issues a DMMA only every ~56 clocks, and it takes both CTAs' 18 warps in phase A
// to fill the pipe. The same run shows phase B at 117 us alone against 230 us next to a neighbour's phase A.)
// (Measured and dropped: widening fp32 -> f64 with integer shifts instead of cvt.f64.f32. The conversions run on the XU
// pipe, which ncu shows busy, but phase A is bound by the DMMA sub-pipe itself — 84 % active while a CTA is in phase A —
// and the extra integer instructions cost more issue slots than the XU conversions: 300 -> 330 us per neuron and CTA.)
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            slb_smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(slb_smem_u32(bar)), "l"(pol)
        : "memory");
}

// upper-triangle 32x32 tiles in row-major order: tile t of pass p belongs to warp t - 9 p
__constant__ unsigned char kTileRow[36] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7};
__constant__ unsigned char kTileCol[36] = {0, 1, 2, 3, 4, 5, 6, 7, 1, 2, 3, 4, 5, 6, 7, 2, 3, 4, 5, 6, 7, 3, 4, 5, 6, 7, 4, 5, 6, 7, 5, 6, 7, 6, 7, 7};

struct Ring {
    uint32_t issued = 0;    // stages requested so far (all threads agree)
    uint32_t consumed = 0;  // stages waited for so far
};

// Thread i (< kT) requests columns [d0, d0 + 32) of example row i into ring slot issued % kStages. `bulk`: X and D allow
// 16-byte async copies; otherwise the row piece is copied with plain loads. Every row thread arrives once per stage, so a
// consumer's wait also orders the plain stores (tail zeroes, fallback copies) before its reads.
__device__ __forceinline__ void request_stage(Smem& sm, Ring& ring, const float* __restrict__ X, int k, int D, int d0, bool bulk,
                                              uint64_t pol) {
    const int i = threadIdx.x;
    const int slot = ring.issued % kStages;
    ring.issued++;
    if (i >= kT) return;
    uint64_t* bar = &sm.full[slot];
    float* dst = sm.stage[slot][i];
    const int valid = min(kStageCols, D - d0);
    if (i >= k) {  // rows past the last example stay zero (zeroed once at kernel start; nothing else writes the ring)
        slb_mbar_arrive(bar);
        return;
    }
    const float* src = X + (int64_t)i * D + d0;
    if (valid < kStageCols)
        for (int c = valid; c < kStageCols; ++c) dst[c] = 0.f;
    if (bulk) {
        slb_mbar_arrive_expect_tx(bar, (uint32_t)valid * 4u);
        bulk_g2s_hint(dst, src, (uint32_t)valid * 4u, bar, pol);
    } else {
        for (int c = 0; c < valid; ++c) dst[c] = __ldg(src + c);
        slb_mbar_arrive(bar);
    }
}

__device__ void gram_f64(const float* __restrict__ X, int k, int D, double* __restrict__ G, Smem& sm, Ring& ring, bool bulk,
                         uint64_t pol) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int n_stage = (D + kStageCols - 1) / kStageCols;
    const int nb = (k + 31) >> 5;  // 32-row blocks that hold examples
    for (int pass = 0; pass < 4; ++pass) {
        const int tile = pass * kWarps + warp;
        const int bi = kTileRow[tile], bj = kTileCol[tile];
        const bool live = bj < nb;  // bi <= bj: the tile touches real examples
        bool any = false;  // uniform over the CTA: a pass whose nine tiles are all past the last example is skipped
#pragma unroll
        for (int w = 0; w < kWarps; ++w) any |= kTileCol[pass * kWarps + w] < nb;
        if (!any) continue;
        double acc[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
        // prologue: fill the ring
        __syncthreads();  // every reader of the previous pass is done with the ring slots
        const int pre = min(kStages, n_stage);
        for (int s = 0; s < pre; ++s) request_stage(sm, ring, X, k, D, s * kStageCols, bulk, pol);
        for (int s = 0; s < n_stage; ++s) {
            const int slot = ring.consumed % kStages;
            slb_mbar_wait(&sm.full[slot], (ring.consumed / kStages) & 1u);
            ring.consumed++;
            if (live) {
                const float* ra = &sm.stage[slot][bi * 32 + g][t];
                const float* rb = &sm.stage[slot][bj * 32 + g][t];
#pragma unroll
                for (int kk = 0; kk < kStageCols / 4; ++kk) {
                    double av[4], bv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) av[a] = (double)ra[a * 8 * kPitch + kk * 4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) bv[b] = (double)rb[b * 8 * kPitch + kk * 4];
                    if (bi == bj) {  // a diagonal tile keeps its upper 8x8 blocks only: 10 of 16 DMMAs (warp-uniform)
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                if (b >= a) dmma884(acc[a][b][0], acc[a][b][1], av[a], bv[b]);
                    } else {
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], av[a], bv[b]);
                    }
                }
            }
            __syncthreads();  // the slot is free again
            if (s + kStages < n_stage) request_stage(sm, ring, X, k, D, (s + kStages) * kStageCols, bulk, pol);
        }
        if (live) {
            // Row sums of G0 while the tile is still in registers (they were a separate G0 x 1 product over the stored slot:
            // 20 us per neuron and CTA). psum[tile][0..31] = sums of the tile's rows -> rows of block bi; psum[tile][32..63] = sums
            // of its columns -> by symmetry, rows of block bj. A diagonal tile holds its upper 8x8 blocks only: the column sums of
            // the strictly upper blocks are folded into its row part. Fixed shuffle order: deterministic.
            double* ps = &sm.tA0[0][0] + tile * 64;
            const bool diag = bi == bj;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                double rs = 0.0;
#pragma unroll
                for (int b = 0; b < 4; ++b) rs += acc[a][b][0] + acc[a][b][1];  // (skipped blocks of a diagonal tile are zero)
                rs += __shfl_xor_sync(0xffffffffu, rs, 1);
                rs += __shfl_xor_sync(0xffffffffu, rs, 2);
                if (t == 0 && !diag) ps[8 * a + g] = rs;
                // diagonal tile: row part of rows 8a + g, completed below with the column sums of blocks (a' < a, a)
                double c0 = 0.0, c1 = 0.0;  // columns 8a + 2t, 8a + 2t + 1 of block column a
#pragma unroll
                for (int a2 = 0; a2 < 4; ++a2)
                    if (!diag || a2 < a) { c0 += acc[a2][a][0]; c1 += acc[a2][a][1]; }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                }
                if (!diag) {
                    if (g == 0) { ps[32 + 8 * a + 2 * t] = c0; ps[32 + 8 * a + 2 * t + 1] = c1; }
                } else {
                    // row 8a + g's share is the sum of column g of block column a: component g % 2 of any lane with t = g / 2
                    const double pick0 = __shfl_sync(0xffffffffu, c0, g >> 1), pick1 = __shfl_sync(0xffffffffu, c1, g >> 1);
                    if (t == 0) ps[8 * a + g] = rs + ((g & 1) ? pick1 : pick0);
                }
            }
            // accumulator (a, b) of lane (g, t) holds G0[32 bi + 8a + g][32 bj + 8b + 2t + {0, 1}]: two adjacent doubles of
            // fragment (band 4 bi + a, group 8 bj + 2b + t / 2) — one 16-byte store; fragments below the diagonal are not kept
            const int NB = (k + 7) >> 3, KQ = (k + 3) >> 2;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int R = bi * 4 + a, Kq = bj * 8 + 2 * b + (t >> 1);
                    if ((bi < bj || b >= a) && R < NB && Kq < KQ)
                        *reinterpret_cast<double2*>(G + (((sm.fb[R] + Kq - 2 * R) << 5) + (g << 2) + ((t & 1) << 1))) =
                            make_double2(acc[a][b][0], acc[a][b][1]);
                }
            }
        }
    }
    __syncthreads();
}

// ---- phase B building block: Y = G0 * W on the FP64 tensor cores -------------------------------------------
// W is k x (8 NT), given as wf(j, n); Y[row][n] is handed to out(row, n, y_n, y_n+1) two columns at a time. Warps 0..7 each
// own four bands of 8 rows (warp w: bands w, w + 8, w + 16, w + 24) and walk the 4-column groups with 16 fragment loads
// in flight per lane; the B fragments (weights) are shared by a warp's four bands. Warp 8 has no rows.
template <int NT, typename WF, typename OUT>
__device__ __forceinline__ void gram_times(const double* __restrict__ U, const int* __restrict__ fb, int k, WF&& wf, OUT&& out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= 8) return;
    const int g = lane >> 2, t = lane & 3;
    const int NB = (k + 7) >> 3, KQ = (k + 3) >> 2;
    int band[4], base[4];
    bool live[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        band[m] = warp + 8 * m;
        live[m] = band[m] < NB;
        base[m] = live[m] ? fb[band[m]] - 2 * band[m] : 0;
    }
    const int lo_off = (t << 2) + (g & 3);  // lane offset inside the half fragment that holds a transposed block
    double acc[4][NT][2];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n) acc[m][n][0] = acc[m][n][1] = 0.0;
    for (int q0 = 0; q0 < KQ; q0 += 4) {
        double a[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int Kq = min(q0 + u, KQ - 1);  // a clamped group is loaded but not used
            const int Rp = Kq >> 1;
            const int tbase = ((fb[Rp] - 2 * Rp + (g >> 2)) << 5) + ((Kq & 1) << 4) + lo_off;  // + 2 * band * 32
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                // on / above the diagonal: fragment (band, Kq) as stored; below: rows 4 (Kq & 1) .. +3 of the two fragments
                // (Kq / 2, 2 band) and (Kq / 2, 2 band + 1), read transposed — G0[8 band + g][4 Kq + t] either way
                const int idx = (Kq >= 2 * band[m]) ? ((base[m] + Kq) << 5) + lane : tbase + (band[m] << 6);
                a[u][m] = live[m] ? U[idx] : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int Kq = q0 + u;
            if (Kq < KQ) {
                double bw[NT];
#pragma unroll
                for (int n = 0; n < NT; ++n) bw[n] = wf(4 * Kq + t, n * 8 + g);
#pragma unroll
                for (int m = 0; m < 4; ++m)
#pragma unroll
                    for (int n = 0; n < NT; ++n) dmma884(acc[m][n][0], acc[m][n][1], a[u][m], bw[n]);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
        if (live[m]) {
#pragma unroll
            for (int n = 0; n < NT; ++n) out(band[m] * 8 + g, n * 8 + 2 * t, acc[m][n][0], acc[m][n][1]);
        }
}

// ---- the kernel ------------------------------------------------------------------------------------
// SM-clock totals of thread 0 of every CTA per phase (slb_polysem_phase_clocks): Gram, row means, k-means++ and first
// assignments, first-iteration sums, Lloyd restarts, score. A handful of clock reads per neuron.
__device__ unsigned long long g_phase_clk[8];

__global__ void __launch_bounds__(kThreads, 2) polysem_kernel(PolyParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int i = threadIdx.x;
    const int k = p.k;
    const bool on = i < k;
    double* G = p.G + (int64_t)blockIdx.x * slot_doubles(k);
    for (int e = i; e < kStages * kT * kPitch; e += kThreads) (&sm.stage[0][0][0])[e] = 0.f;
    if (i <= kT / 8) sm.fb[i] = frag_base(i, (k + 3) >> 2);
    if (i == 0) {
        for (int s = 0; s < kStages; ++s) slb_mbar_init(&sm.full[s], kT);
        slb_fence_mbar_init();
    }
    __syncthreads();
    Ring ring;
    const uint64_t pol = evict_first_policy();
    const bool bulk = (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.V) & 15) == 0);

    unsigned long long clk[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t neuron = blockIdx.x; neuron < p.C; neuron += gridDim.x) {
        long long t0 = clock64(), t1;
        gram_f64(p.V + neuron * (int64_t)k * p.D, k, p.D, G, sm, ring, bulk, pol);
        __threadfence_block();
        t1 = clock64(); clk[0] += t1 - t0; t0 = t1;

        // row means, grand mean, centred diagonal, tolerance
        // row means from the per-tile sums phase A left in sm.tA0 (fixed order: blocks left of the diagonal, then the row's own)
        double ri = 0.0;
        if (on) {
            const int B = i >> 5, o = i & 31, nb = (k + 31) >> 5;
            const double* ps = &sm.tA0[0][0];
            for (int b2 = 0; b2 < B; ++b2) ri += ps[(b2 * 8 - b2 * (b2 - 1) / 2 + (B - b2)) * 64 + 32 + o];
            for (int b2 = B; b2 < nb; ++b2) ri += ps[(B * 8 - B * (B - 1) / 2 + (b2 - B)) * 64 + o];
            ri /= (double)k;
        }
        __syncthreads();  // sm.tA0 is free again
        const double m = block_sum1(ri, sm) / (double)k;
        const double g_ii = on ? G[frag_elem(sm.fb, i, i)] : 0.0;
        const double di = on ? g_ii - 2.0 * ri + m : 0.0;
        sm.r[i] = ri;
        sm.diag[i] = di;
        double tv[2] = {di, g_ii};
        block_sum<2>(tv, sm);  // also publishes sm.r / sm.diag
        const double tol = tv[0] / ((double)k * (double)p.D) * 1e-4;
        const bool nonfinite = !(tv[1] < 0x1p255);  // an Inf / NaN example, or a true overflow: sklearn raises on such input

        t1 = clock64(); clk[1] += t1 - t0; t0 = t1;
        auto gc = [&](int j) -> double {  // Gc[i][j] from the packed upper triangle
            return G[frag_elem(sm.fb, i, j)] - ri - sm.r[j] + m;
        };

        double best_inertia = 0.0;
        bool have_best = false;

        // ---- pass 1 over the restarts: k-means++ and the FIRST assignment of every restart -------------------------
        // In the first Lloyd iteration about half of the examples enter cluster 0, so updating tA costs ~k/2 row reads per
        // restart. The first assignments only depend on the seeds, so they are all taken first and ONE sweep over the
        // Gram matrix then serves every restart (each entry is loaded once and added into up to n_init independent sums).
        unsigned labw = 0, in0w = 0;  // bit `it`: first label of example i / example i starts in cluster 0 (after relocation)
        for (int it = 0; it < p.n_init; ++it) {
            // k-means++ (sklearn _kmeans_plusplus, n_local_trials = 2)
            const int i0 = p.first[it];
            const double gi0 = on ? gc(i0) : 0.0;
            const double closest = on ? fmax(di - 2.0 * gi0 + sm.diag[i0], 0.0) : 0.0;
            const double pot = block_sum1(closest, sm);
            const double cum = block_scan(closest, sm);
            int cnts[2] = {(on && cum < p.rand[2 * it] * pot) ? 1 : 0, (on && cum < p.rand[2 * it + 1] * pot) ? 1 : 0};
            block_count<2>(cnts, sm);
            const int cand0 = min(cnts[0], k - 1), cand1 = min(cnts[1], k - 1);
            const double gc0 = on ? gc(cand0) : 0.0, gc1 = on ? gc(cand1) : 0.0;
            double pots[2] = {on ? fmin(closest, fmax(di - 2.0 * gc0 + sm.diag[cand0], 0.0)) : 0.0,
                              on ? fmin(closest, fmax(di - 2.0 * gc1 + sm.diag[cand1], 0.0)) : 0.0};
            block_sum<2>(pots, sm);
            const int i1 = (pots[1] < pots[0]) ? cand1 : cand0;
            // first E-step (+ empty-cluster relocation) from the two seeds
            const double s0 = gi0, s1 = (i1 == cand1) ? gc1 : gc0;
            const double n0 = sm.diag[i0], n1 = sm.diag[i1];
            const int lab = (on && (n1 - 2.0 * s1) < (n0 - 2.0 * s0)) ? 1 : 0;
            int mk = lab;
            int c1[1] = {(on && lab == 0) ? 1 : 0};
            block_count<1>(c1, sm);
            int cnt0 = c1[0], cnt1 = k - cnt0;
            if (cnt0 == 0 || cnt1 == 0) {
                const int o = (cnt1 == 0) ? 0 : 1;
                double dist = on ? di - 2.0 * (o ? s1 : s0) + (o ? n1 : n0) : -1.0;
                int far = i;
                block_argmax(dist, far, sm);
                if (dist > 0.0) {
                    if (i == far) mk = 1 - o;
                    if (o == 0) { cnt0 -= 1; } else { cnt0 = 1; }
                }
            }
            labw |= (unsigned)lab << it;
            in0w |= (unsigned)(on && mk == 0) << it;
            if (i == 0) { sm.seed0[it] = i1; sm.cnt0[it] = cnt0; }
        }
        __syncthreads();
        sm.in0[i] = (unsigned short)in0w;
        __syncthreads();
        t1 = clock64(); clk[2] += t1 - t0; t0 = t1;
        // Y[i][q] = sum over the first members j of restart q of G0[i][j], all restarts in one product; centring:
        //   sum_j w_j Gc[i][j] = Y[i] - r_i c - S + m c,   c = |members|,   S = sum_j w_j r_j = (1/k) sum_i Y[i]  (G0 symmetric)
        gram_times<2>(G, sm.fb, k, [&](int j, int n) { return (((unsigned)sm.in0[j] >> n) & 1u) ? 1.0 : 0.0; },
                      [&](int row, int col, double y0, double y1) {
                          if (col < p.n_init) sm.tA0[col][row] = y0;
                          if (col + 1 < p.n_init) sm.tA0[col + 1][row] = y1;
                      });
        __syncthreads();
        {
            double tot[kMaxInit];
#pragma unroll
            for (int q = 0; q < kMaxInit; ++q) tot[q] = (on && q < p.n_init) ? sm.tA0[q][i] : 0.0;
            block_sumN<kMaxInit>(tot, sm);
            if (on) {
#pragma unroll
                for (int q = 0; q < kMaxInit; ++q)
                    if (q < p.n_init) {
                        const double c = (double)sm.cnt0[q];
                        sm.tA0[q][i] = sm.tA0[q][i] - ri * c - tot[q] / (double)k + m * c;
                    }
            }
        }
        __syncthreads();
        t1 = clock64(); clk[3] += t1 - t0; t0 = t1;
        // ---- pass 2: the Lloyd iterations of every restart, best of n_init ------------------------------------
        for (int it = 0; it < p.n_init; ++it) {
            const int i0 = p.first[it], i1 = sm.seed0[it];
            double s0 = on ? gc(i0) : 0.0, s1 = on ? gc(i1) : 0.0;
            double n0 = sm.diag[i0], n1 = sm.diag[i1];
            double tA = on ? sm.tA0[it][i] : 0.0;  // sum over members of cluster 0 of Gc[i][j]
            int cur = 1;       // current M-step membership of example i
            int lab = 0, lab_old = -1;
            bool strict = false;
            for (int iter = 0; iter < kMaxIter; ++iter) {
                int mk, cnt0, cnt1;
                bool same;
                if (iter == 0) {  // taken in pass 1
                    lab = (labw >> it) & 1u;
                    mk = on ? (((in0w >> it) & 1u) ? 0 : 1) : 0;
                    cnt0 = sm.cnt0[it];
                    cnt1 = k - cnt0;
                    same = false;
                } else {
                    lab = (on && (n1 - 2.0 * s1) < (n0 - 2.0 * s0)) ? 1 : 0;
                    mk = lab;
                    int c2[2] = {(on && lab == 0) ? 1 : 0, (on && lab != lab_old) ? 1 : 0};
                    block_count<2>(c2, sm);
                    cnt0 = c2[0];
                    cnt1 = k - cnt0;
                    same = c2[1] == 0;
                    if (cnt0 == 0 || cnt1 == 0) {
                        // _relocate_empty_clusters_dense: the point farthest from its centre founds the empty cluster
                        const int o = (cnt1 == 0) ? 0 : 1;
                        double dist = on ? di - 2.0 * (o ? s1 : s0) + (o ? n1 : n0) : -1.0;
                        int far = i;
                        block_argmax(dist, far, sm);
                        if (dist > 0.0) {
                            if (i == far) mk = 1 - o;
                            if (o == 0) { cnt0 -= 1; cnt1 = 1; } else { cnt1 -= 1; cnt0 = 1; }
                        }
                    }
                    // incremental update of tA with the examples that changed side: the movers are compacted into an
                    // ascending list (ballot + per-warp offsets) so that every thread walks only them, eight independent
                    // L2 loads in flight at a time; the additions keep the ascending order (same bits as a full scan)
                    const int dj = on ? (mk == 0) - (cur == 0) : 0;
                    const unsigned bal = __ballot_sync(0xffffffffu, dj != 0);
                    __syncthreads();  // previous readers of sm.moved / sm.wcnt are done
                    if ((i & 31) == 0) sm.wcnt[i >> 5] = __popc(bal);
                    __syncthreads();
                    int base = 0, n_moved = 0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) {
                        const int c = sm.wcnt[w];
                        if (w < (i >> 5)) base += c;
                        n_moved += c;
                    }
                    if (dj != 0) sm.moved[base + __popc(bal & ((1u << (i & 31)) - 1u))] = (short)(dj > 0 ? i : ~i);
                    __syncthreads();
                    if (on) {
                        for (int e0 = 0; e0 < n_moved; e0 += 8) {
                            double gv[8];
                            int jj[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int e = min(e0 + u, n_moved - 1);
                                const int code = sm.moved[e];
                                jj[u] = code;
                                const int j = code >= 0 ? code : ~code;
                                gv[u] = G[frag_elem(sm.fb, i, j)];
                            }
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                if (e0 + u < n_moved) {
                                    const int j = jj[u] >= 0 ? jj[u] : ~jj[u];
                                    const double gcj = gv[u] - ri - sm.r[j] + m;
                                    tA += jj[u] >= 0 ? gcj : -gcj;
                                }
                            }
                        }
                    }
                }
                double s0n, s1n, n0n, n1n, cross0, cross1;
                cur = mk;
                if (cnt0 == 0 || cnt1 == 0) {
                    // every point coincides with the surviving centre; sklearn leaves the empty centre at 0
                    const int e = (cnt0 == 0) ? 0 : 1;
                    const double cn = (double)(e ? cnt0 : cnt1);
                    const double t = e ? tA : -tA;  // sum over the surviving side (= all points)
                    const double sn = on ? t / cn : 0.0;
                    double v[2] = {on ? sn : 0.0, on ? (e ? s0 : s1) : 0.0};
                    block_sum<2>(v, sm);
                    if (e) { s0n = sn; s1n = 0.0; n0n = v[0] / cn; n1n = 0.0; cross0 = v[1] / cn; cross1 = 0.0; }
                    else   { s1n = sn; s0n = 0.0; n1n = v[0] / cn; n0n = 0.0; cross1 = v[1] / cn; cross0 = 0.0; }
                } else {
                    s0n = on ? tA / (double)cnt0 : 0.0;
                    s1n = on ? -tA / (double)cnt1 : 0.0;
                    double v[4] = {(on && mk == 0) ? s0n : 0.0, (on && mk == 1) ? s1n : 0.0, (on && mk == 0) ? s0 : 0.0,
                                   (on && mk == 1) ? s1 : 0.0};
                    block_sum<4>(v, sm);
                    n0n = v[0] / (double)cnt0;
                    n1n = v[1] / (double)cnt1;
                    cross0 = v[2] / (double)cnt0;
                    cross1 = v[3] / (double)cnt1;
                }
                const double shift = (n0n - 2.0 * cross0 + n0) + (n1n - 2.0 * cross1 + n1);
                s0 = s0n; s1 = s1n; n0 = n0n; n1 = n1n;
                if (same) { strict = true; break; }
                if (shift <= tol) break;
                lab_old = lab;
            }
            if (!strict) lab = (on && (n1 - 2.0 * s1) < (n0 - 2.0 * s0)) ? 1 : 0;
            const double inertia = block_sum1(on ? di - 2.0 * (lab ? s1 : s0) + (lab ? n1 : n0) : 0.0, sm);

            // _is_same_clustering(labels, best_labels): labels -> best_labels must be a function
            bool take = !have_best;
            if (have_best) {
                int c4a[2] = {(on && lab == 0 && sm.best_lab[i] == 0) ? 1 : 0, (on && lab == 0 && sm.best_lab[i] == 1) ? 1 : 0};
                int c4b[2] = {(on && lab == 1 && sm.best_lab[i] == 0) ? 1 : 0, (on && lab == 1 && sm.best_lab[i] == 1) ? 1 : 0};
                block_count<2>(c4a, sm);
                block_count<2>(c4b, sm);
                const bool same_clustering = !((c4a[0] > 0 && c4a[1] > 0) || (c4b[0] > 0 && c4b[1] > 0));
                take = inertia < best_inertia && !same_clustering;
            }
            __syncthreads();
            if (take) {
                sm.best_lab[i] = (unsigned char)lab;
                sm.best_mask[i] = (unsigned char)cur;
                best_inertia = inertia;
                have_best = true;
            }
            __syncthreads();
        }

        t1 = clock64(); clk[4] += t1 - t0; t0 = t1;
        // ---- score ----
        int cl[2] = {(on && sm.best_lab[i] == 0) ? 1 : 0, (on && sm.best_mask[i] == 0) ? 1 : 0};
        block_count<2>(cl, sm);
        const int l0 = cl[0], l1 = k - cl[0];
        const int ca = cl[1], cb = k - cl[1];
        double result;
        if (p.replace_empty && min(l0, l1) < 2) {
            // 1 - mean_{i < min(10,k)} cos(mean_k V, V[:, i])   (scores.py:178-184)
            const int ns = min(10, k);
            const double c = (on && i < ns) ? ri / (fmax(sqrt(m), 1e-12) * fmax(sqrt(g_ii), 1e-12)) : 0.0;
            result = 1.0 - block_sum1(c, sm) / (double)ns;
        } else {
            gram_times<1>(G, sm.fb, k,
                          [&](int j, int n) { return (j < k && n < 2 && (int)sm.best_mask[j] == n) ? 1.0 : 0.0; },
                          [&](int row, int col, double y0, double y1) {
                              if (col == 0) { sm.tA0[0][row] = y0; sm.tA0[1][row] = y1; }
                          });
            __syncthreads();
            const double wa = on ? sm.tA0[0][i] : 0.0, wb = on ? sm.tA0[1][i] : 0.0;
            const bool ina = on && sm.best_mask[i] == 0, inb = on && sm.best_mask[i] == 1;
            double v[4] = {ina ? wa : 0.0, inb ? wa : 0.0, inb ? wb : 0.0, on ? (ca == 0 ? (inb ? ri : 0.0) : (ina ? ri : 0.0)) : 0.0};
            block_sum<4>(v, sm);
            if (ca == 0 || cb == 0) {
                // degenerate: the empty centre sits at X_mean after sklearn's `best_centers += X_mean`
                const double co = (double)(ca == 0 ? cb : ca);
                const double soo = (ca == 0 ? v[2] : v[0]) / (co * co);
                const double som = v[3] / co;
                result = 1.0 - som / (fmax(sqrt(soo), 1e-12) * fmax(sqrt(m), 1e-12));
            } else {
                const double saa = v[0] / ((double)ca * (double)ca);
                const double sab = v[1] / ((double)ca * (double)cb);
                const double sbb = v[2] / ((double)cb * (double)cb);
                result = 1.0 - sab / (fmax(sqrt(saa), 1e-12) * fmax(sqrt(sbb), 1e-12));
            }
        }
        if (i == 0) p.out[neuron] = nonfinite ? __longlong_as_double(0x7FF8000000000000ll) : result;
        __syncthreads();
        t1 = clock64(); clk[5] += t1 - t0;
        clk[6] += 1;
    }
    if (i == 0)
        for (int c = 0; c < 7; ++c) atomicAdd(&g_phase_clk[c], clk[c]);
}

// SLB_POLYSEM_CTAS_PER_SM=1 (diagnostic): one CTA per SM, to time the phases without a neighbour on the SM
int poly_ctas_per_sm() {
    static int v = 0;
    if (v == 0) {
        const char* e = getenv("SLB_POLYSEM_CTAS_PER_SM");
        v = (e && e[0] == '1') ? 1 : 2;
    }
    return v;
}
int poly_grid(int64_t C) { return (int)std::min<int64_t>(C, (int64_t)slb_sm_count() * poly_ctas_per_sm()); }

}  // namespace

extern "C" int slb_polysem_phase_clocks(uint64_t* out7, int reset) {
    SLB_REQUIRE(out7, SLB_EINVAL, "slb_polysem_phase_clocks: null pointer");
    unsigned long long host[8] = {0};
    SLB_CUDA_OK(cudaMemcpyFromSymbol(host, g_phase_clk, sizeof(host)));
    for (int c = 0; c < 7; ++c) out7[c] = host[c];
    if (reset) {
        unsigned long long zero[8] = {0};
        SLB_CUDA_OK(cudaMemcpyToSymbol(g_phase_clk, zero, sizeof(zero)));
    }
    return SLB_OK;
}

extern "C" size_t slb_polysem_workspace_bytes(int64_t C, int64_t k) {
    if (C <= 0 || k <= 0 || k > kT) return 0;
    return (size_t)poly_grid(C) * (size_t)slot_doubles(k) * sizeof(double);
}

extern "C" int slb_polysem_2means(const float* V, int64_t C, int64_t k, int64_t D, const int64_t* first_centers,
                                  const double* local_trial_uniforms, int n_init, int replace_empty_clusters, double* out,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    SLB_REQUIRE(C >= 0 && k > 0 && D > 0, SLB_EINVAL, "slb_polysem_2means: bad size");
    if (C == 0) return SLB_OK;
    SLB_REQUIRE(V && out && workspace && first_centers && local_trial_uniforms, SLB_EINVAL,
                "slb_polysem_2means: null pointer (first_centers / local_trial_uniforms are HOST arrays)");
    SLB_REQUIRE(k <= kT, SLB_EUNSUPPORTED, "slb_polysem_2means: at most %d examples per neuron (got %lld)", kT, (long long)k);
    SLB_REQUIRE(D < (1ll << 31), SLB_EUNSUPPORTED, "slb_polysem_2means: D too large");
    SLB_REQUIRE(n_init >= 1 && n_init <= kMaxInit, SLB_EUNSUPPORTED, "slb_polysem_2means: 1 <= n_init <= %d", kMaxInit);
    SLB_REQUIRE(((uintptr_t)workspace % 16) == 0, SLB_EINVAL, "slb_polysem_2means: workspace must be 16-byte aligned");
    const size_t need = slb_polysem_workspace_bytes(C, k);
    SLB_REQUIRE(workspace_bytes >= need, SLB_EWORKSPACE, "slb_polysem_2means: workspace needs %zu bytes, got %zu", need,
                workspace_bytes);
    SlbProfScope prof("K8 polysem_2means", stream, 0.0, 4.0 * (double)C * (double)k * (double)D);
    PolyParams p{};
    p.V = V; p.C = C; p.k = (int)k; p.D = (int)D; p.n_init = n_init; p.replace_empty = replace_empty_clusters ? 1 : 0;
    for (int i = 0; i < n_init; ++i) {
        SLB_REQUIRE(first_centers[i] >= 0 && first_centers[i] < k, SLB_EINVAL, "slb_polysem_2means: first centre out of range");
        p.first[i] = (int)first_centers[i];
        p.rand[2 * i] = local_trial_uniforms[2 * i];
        p.rand[2 * i + 1] = local_trial_uniforms[2 * i + 1];
    }
    p.G = static_cast<double*>(workspace);
    p.out = out;
    const size_t smem = poly_ctas_per_sm() == 1 ? std::max(sizeof(Smem), (size_t)120 * 1024) : sizeof(Smem);
    SLB_CUDA_OK(cudaFuncSetAttribute(polysem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    polysem_kernel<<<poly_grid(C), kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    SLB_LAUNCH_OK("polysem_2means");
    return SLB_OK;
}
