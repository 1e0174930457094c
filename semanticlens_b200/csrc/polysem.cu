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
//          except at true f64 near-ties), 64x64 blocks of the upper triangle, 4x4 register tiles, written to the CTA's
//          own k*k slot of the workspace (stays L2 resident: one slot per resident CTA).
// Phase B  thread i = example i: k-means++ with sklearn's RandomState stream (data independent: the first-centre
//          index and the two local-trial uniforms per init are precomputed on the host), Lloyd to strict convergence or
//          tolerance, empty-cluster relocation, best of n_init by inertia with sklearn's _is_same_clustering guard.
//          oracle/polysem.py::kmeans2_gram is the line-by-line CPU statement of this phase.
#include "slb_common.cuh"

#include <algorithm>

namespace {

constexpr int kT = 256;     // threads per CTA = max examples per neuron
constexpr int kMaxInit = 16;
constexpr int kMaxIter = 300;  // sklearn default max_iter

struct PolyParams {
    const float* V;
    int64_t C;
    int k, D;
    int n_init;
    int replace_empty;
    int first[kMaxInit];
    double rand[2 * kMaxInit];
    double* G;  // workspace: gridDim.x slots of k*k doubles
    double* out;
};

struct Smem {
    double tileA[64][33];
    double tileB[64][33];
    double r[kT];
    double diag[kT];
    double red[8 * 4];
    double scan[8];
    int ired[8 * 2];
    signed char delta[kT];
    unsigned char best_lab[kT];
    unsigned char best_mask[kT];
};

// ---- block primitives (256 threads, every thread must call) ----------------------------------------
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
        for (int w = 0; w < 8; ++w) t += sm.red[w * 4 + n];
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
        for (int w = 0; w < 8; ++w) t += sm.ired[w * 2 + n];
        v[n] = t;
    }
}

// inclusive prefix sum over threads 0..255 (np.cumsum)
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
    for (int w = 1; w < 8; ++w) {
        const double v2 = sm.red[w * 4];
        const int i2 = sm.ired[w * 2];
        if (v2 > val || (v2 == val && i2 < idx)) { val = v2; idx = i2; }
    }
}

// ---- phase A: G0 = X X^T (float64) ---------------------------------------------------------------
__device__ void gram_f64(const float* __restrict__ X, int k, int D, double* __restrict__ G, Smem& sm) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int nb = (k + 63) >> 6;
    for (int bi = 0; bi < nb; ++bi) {
        for (int bj = bi; bj < nb; ++bj) {
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
            for (int d0 = 0; d0 < D; d0 += 32) {
                __syncthreads();
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int e = tid + q * kT;
                    const int row = e >> 5, col = e & 31;
                    const int d = d0 + col;
                    const int ra = bi * 64 + row, rb = bj * 64 + row;
                    sm.tileA[row][col] = (ra < k && d < D) ? (double)__ldg(X + (int64_t)ra * D + d) : 0.0;
                    if (bj != bi) sm.tileB[row][col] = (rb < k && d < D) ? (double)__ldg(X + (int64_t)rb * D + d) : 0.0;
                }
                __syncthreads();
                const double(*tb)[33] = (bj != bi) ? sm.tileB : sm.tileA;
#pragma unroll 8
                for (int d = 0; d < 32; ++d) {
                    double av[4], bv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) av[a] = sm.tileA[ty + 16 * a][d];
#pragma unroll
                    for (int b = 0; b < 4; ++b) bv[b] = tb[tx + 16 * b][d];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
                }
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int i = bi * 64 + ty + 16 * a, j = bj * 64 + tx + 16 * b;
                    if (i < k && j < k) {
                        G[(int64_t)i * k + j] = acc[a][b];
                        if (bj != bi) G[(int64_t)j * k + i] = acc[a][b];
                    }
                }
            }
        }
    }
    __syncthreads();
}

// ---- the kernel ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) polysem_kernel(PolyParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const int i = threadIdx.x;
    const int k = p.k;
    const bool on = i < k;
    double* G = p.G + (int64_t)blockIdx.x * k * k;

    for (int64_t neuron = blockIdx.x; neuron < p.C; neuron += gridDim.x) {
        gram_f64(p.V + neuron * (int64_t)k * p.D, k, p.D, G, sm);
        __threadfence_block();

        // row means, grand mean, centred diagonal, tolerance
        double ri = 0.0;
        if (on) {
            for (int j = 0; j < k; ++j) ri += __ldcg(G + (int64_t)j * k + i);
            ri /= (double)k;
        }
        const double m = block_sum1(on ? ri : 0.0, sm) / (double)k;
        const double g_ii = on ? __ldcg(G + (int64_t)i * k + i) : 0.0;
        const double di = on ? g_ii - 2.0 * ri + m : 0.0;
        sm.r[i] = ri;
        sm.diag[i] = di;
        const double tol = block_sum1(di, sm) / ((double)k * (double)p.D) * 1e-4;  // also publishes sm.r / sm.diag

        auto gc = [&](int j) -> double {  // Gc[i][j], read down column i of the symmetric G0 (coalesced over i)
            return __ldcg(G + (int64_t)j * k + i) - ri - sm.r[j] + m;
        };

        double best_inertia = 0.0;
        bool have_best = false;

        for (int it = 0; it < p.n_init; ++it) {
            // ---- k-means++ (sklearn _kmeans_plusplus, n_local_trials = 2) ----
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

            // ---- Lloyd ----
            double s0 = gi0, s1 = (i1 == cand1) ? gc1 : gc0;
            double n0 = sm.diag[i0], n1 = sm.diag[i1];
            double tA = 0.0;   // sum over members of cluster 0 of Gc[i][j]
            int cur = 1;       // current M-step membership of example i (nothing is in cluster 0 yet)
            int lab = 0, lab_old = -1;
            bool strict = false;
            for (int iter = 0; iter < kMaxIter; ++iter) {
                lab = (on && (n1 - 2.0 * s1) < (n0 - 2.0 * s0)) ? 1 : 0;
                int mk = lab;
                int c2[2] = {(on && lab == 0) ? 1 : 0, (on && lab != lab_old) ? 1 : 0};
                block_count<2>(c2, sm);
                int cnt0 = c2[0], cnt1 = k - cnt0;
                const bool same = c2[1] == 0;
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
                double s0n, s1n, n0n, n1n, cross0, cross1;
                // incremental update of tA with the examples that changed side
                __syncthreads();
                sm.delta[i] = on ? (signed char)((mk == 0) - (cur == 0)) : 0;
                __syncthreads();
                if (on) {
                    for (int j = 0; j < k; ++j) {
                        const int dj = sm.delta[j];
                        if (dj != 0) tA += (double)dj * gc(j);
                    }
                }
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
            double wa = 0.0, wb = 0.0;
            if (on) {
                for (int j = 0; j < k; ++j) {
                    const double g = __ldcg(G + (int64_t)j * k + i);
                    if (sm.best_mask[j] == 0) wa += g; else wb += g;
                }
            }
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
        if (i == 0) p.out[neuron] = result;
        __syncthreads();
    }
}

int poly_grid(int64_t C) { return (int)std::min<int64_t>(C, (int64_t)slb_sm_count() * 2); }

}  // namespace

extern "C" size_t slb_polysem_workspace_bytes(int64_t C, int64_t k) {
    if (C <= 0 || k <= 0 || k > kT) return 0;
    return (size_t)poly_grid(C) * (size_t)k * (size_t)k * sizeof(double);
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
    SLB_REQUIRE(((uintptr_t)workspace % 8) == 0, SLB_EINVAL, "slb_polysem_2means: workspace must be 8-byte aligned");
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
    const size_t smem = sizeof(Smem);
    SLB_CUDA_OK(cudaFuncSetAttribute(polysem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    polysem_kernel<<<poly_grid(C), kT, smem, static_cast<cudaStream_t>(stream)>>>(p);
    SLB_LAUNCH_OK("polysem_2means");
    return SLB_OK;
}
