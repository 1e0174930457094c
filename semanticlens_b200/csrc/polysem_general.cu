// K8g — polysemanticity_score for any n_clusters and any number of examples (reference semanticlens/scores.py:132-185
// with n_clusters != 2, or more than 256 examples per neuron). The fast path (polysem.cu) covers the reference's default
// call — n_clusters = 2, k <= 256 — on the neuron's Gram matrix; this kernel is the general statement of the same
// sklearn fit in sample space: KMeans(n_clusters, n_init, random_state).fit(examples) in float64, i.e.
//   centring by the feature means, tolerance = mean feature variance * 1e-4                     (_kmeans.py _tolerance)
//   k-means++ with 2 + int(log n_clusters) local trials, the RandomState draws precomputed on the host  (_kmeans_plusplus)
//   Lloyd: argmin_j |c_j|^2 - 2 x.c_j (first minimum), centre = member mean, empty clusters relocated to the farthest
//   points, clusters that stay empty placed at "the location of the biggest cluster" read in index order, strict /
//   tolerance convergence, labels recomputed after a tolerance stop                       (_k_means_lloyd.pyx, _k_means_common.pyx)
//   best of n_init by inertia unless the labelling is the same clustering                              (KMeans.fit)
// then 1 - clarity of the un-centred centres, or the reference's fallback when a cluster has fewer than two members.
// oracle/polysem.py::kmeans_direct is the line-by-line CPU statement (pinned against sklearn itself).
//
// One CTA (8 warps) per neuron, grid-stride. The examples stay fp32 in global memory (L2-resident: k D 4 bytes) and are
// centred on the fly; a warp owns an example for every dot product (lanes stride the features: coalesced), a thread owns a
// feature for every centre update (sequential over the examples: deterministic sums). Per-CTA state (centres, labels,
// distances) lives in a caller-provided workspace.
#include "slb_common.cuh"

#include <algorithm>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kMaxM = 8;      // clusters
constexpr int kMaxL = 4;      // local trials: 2 + int(log 8)
constexpr int kMaxIter = 300;  // sklearn default max_iter

struct KmParams {
    const float* V;
    int64_t C;
    int k, D, m, n_init, L, replace_empty;
    const int* first;    // device [n_init]
    const double* rand;  // device [n_init][m - 1][L]
    double* ws;
    int64_t ws_doubles;  // per CTA
    double* out;
};

struct KmSmem {
    double red[kWarps][kMaxM + 2];
    double csq[kMaxM];
    double w[kMaxM];
    double vals[kMaxL];
    double bval;
    int cand[kMaxL];
    int cnt[kMaxM];
    unsigned mask[kMaxM];
    int bidx;
    int flag;
};

// every thread calls; returns the block-wide sums of v[0..N) (fixed order: lanes by butterfly, warps in index order)
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], KmSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int n = 0; n < N; ++n)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[n] += __shfl_xor_sync(0xffffffffu, v[n], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int n = 0; n < N; ++n) sm.red[warp][n] = v[n];
    __syncthreads();
#pragma unroll
    for (int n = 0; n < N; ++n) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += sm.red[w][n];
        v[n] = t;
    }
}

__device__ __forceinline__ double block_sum1(double x, KmSmem& sm) {
    double v[1] = {x};
    block_sum<1>(v, sm);
    return v[0];
}

// (largest value, lowest index attaining it); every thread calls
__device__ __forceinline__ void block_argmax(double& val, int& idx, KmSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, val, o);
        const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
        if (v2 > val || (v2 == val && i2 < idx)) { val = v2; idx = i2; }
    }
    __syncthreads();
    if (lane == 0) { sm.red[warp][0] = val; sm.red[warp][1] = (double)idx; }
    __syncthreads();
    val = sm.red[0][0];
    idx = (int)sm.red[0][1];
    for (int w = 1; w < kWarps; ++w) {
        const double v2 = sm.red[w][0];
        const int i2 = (int)sm.red[w][1];
        if (v2 > val || (v2 == val && i2 < idx)) { val = v2; idx = i2; }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct Neuron {
    const float* X;  // (k, D) fp32
    int k, D, m;
    double* mean;    // [D]
    double* xsq;     // [k]  |x_i - mean|^2
};

// dots of centred example i with NV vectors given by vec(v, d) (centred space); lane-strided, warp-reduced (all lanes get them)
template <int NV, typename F>
__device__ __forceinline__ void warp_dots(const Neuron& n, int i, int nv, F&& vec, double (&out)[NV]) {
    const int lane = threadIdx.x & 31;
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = 0.0;
    const float* row = n.X + (int64_t)i * n.D;
    for (int d = lane; d < n.D; d += 32) {
        const double x = (double)row[d] - n.mean[d];
#pragma unroll
        for (int v = 0; v < NV; ++v)
            if (v < nv) acc[v] = fma(x, vec(v, d), acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) out[v] = warp_sum(acc[v]);
}

// labels[i] = argmin_j csq[j] - 2 x_i.cur[j] (first minimum); returns whether any label differs from lab_old
__device__ bool e_step(const Neuron& n, const double* __restrict__ cur, int* __restrict__ labels, const int* __restrict__ lab_old,
                       KmSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // centre norms
    for (int j = warp; j < n.m; j += kWarps) {
        double a = 0.0;
        for (int d = lane; d < n.D; d += 32) a = fma(cur[(int64_t)j * n.D + d], cur[(int64_t)j * n.D + d], a);
        a = warp_sum(a);
        if (lane == 0) sm.csq[j] = a;
    }
    if (threadIdx.x == 0) sm.flag = 0;
    __syncthreads();
    int changed = 0;
    for (int i = warp; i < n.k; i += kWarps) {
        double dots[kMaxM];
        warp_dots<kMaxM>(n, i, n.m, [&](int v, int d) { return cur[(int64_t)v * n.D + d]; }, dots);
        int best = 0;
        double bv = sm.csq[0] - 2.0 * dots[0];
#pragma unroll
        for (int j = 1; j < kMaxM; ++j)
            if (j < n.m) {
                const double pd = sm.csq[j] - 2.0 * dots[j];
                if (pd < bv) { bv = pd; best = j; }
            }
        if (lane == 0) {
            labels[i] = best;
            if (lab_old == nullptr || lab_old[i] != best) changed = 1;
        }
    }
    if (changed) sm.flag = 1;  // benign race: every writer stores 1
    __syncthreads();
    const bool any = sm.flag != 0;
    __syncthreads();
    return any;
}

// direct squared distance of example i to centre c (centred space), all lanes get it
__device__ __forceinline__ double warp_dist_direct(const Neuron& n, int i, const double* __restrict__ c) {
    const int lane = threadIdx.x & 31;
    const float* row = n.X + (int64_t)i * n.D;
    double a = 0.0;
    for (int d = lane; d < n.D; d += 32) {
        const double t = ((double)row[d] - n.mean[d]) - c[d];
        a = fma(t, t, a);
    }
    return warp_sum(a);
}

__global__ void __launch_bounds__(kThreads) polysem_kmeans_kernel(KmParams p) {
    __shared__ KmSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k, D = p.D, m = p.m;
    double* ws = p.ws + (int64_t)blockIdx.x * p.ws_doubles;
    double* mean = ws;
    double* bufA = mean + D;
    double* bufB = bufA + (int64_t)m * D;
    double* best_c = bufB + (int64_t)m * D;
    double* xsq = best_c + (int64_t)m * D;
    double* closest = xsq + k;
    double* dist = closest + k;
    int* labA = reinterpret_cast<int*>(dist + k);
    int* labB = labA + k;
    int* best_lab = labB + k;

    for (int64_t neuron = blockIdx.x; neuron < p.C; neuron += gridDim.x) {
        Neuron n{p.V + neuron * (int64_t)k * D, k, D, m, mean, xsq};
        // ---- feature means, mean feature variance (tolerance) ------------------------------------------------
        double var_part = 0.0;
        for (int d = tid; d < D; d += kThreads) {
            double s = 0.0;
            for (int i = 0; i < k; ++i) s += (double)n.X[(int64_t)i * D + d];
            const double mu = s / (double)k;
            double v = 0.0;
            for (int i = 0; i < k; ++i) {
                const double t = (double)n.X[(int64_t)i * D + d] - mu;
                v = fma(t, t, v);
            }
            mean[d] = mu;
            var_part += v / (double)k;
        }
        const double tol = block_sum1(var_part, sm) / (double)D * 1e-4;  // (its barriers also publish mean[])
        bool nonfinite = !(tol == tol) || !(tol < 1.7e308);
        for (int i = warp; i < k; i += kWarps) {
            const float* row = n.X + (int64_t)i * D;
            double a = 0.0;
            for (int d = lane; d < D; d += 32) {
                const double t = (double)row[d] - mean[d];
                a = fma(t, t, a);
            }
            a = warp_sum(a);
            if (lane == 0) xsq[i] = a;
        }
        __syncthreads();

        double best_inertia = 0.0;
        bool have_best = false;
        for (int it = 0; it < p.n_init && !nonfinite; ++it) {
            double* cur = bufA;
            double* nxt = bufB;
            int* labels = labA;
            int* lab_old = labB;
            // ---- k-means++ -----------------------------------------------------------------------------------
            const int i0 = p.first[it];
            for (int d = tid; d < D; d += kThreads) cur[d] = (double)n.X[(int64_t)i0 * D + d] - mean[d];
            __syncthreads();
            double pot_part = 0.0;
            for (int i = warp; i < k; i += kWarps) {
                double dots[1];
                warp_dots<1>(n, i, 1, [&](int, int d) { return cur[d]; }, dots);
                const double c = fmax(xsq[i] - 2.0 * dots[0] + xsq[i0], 0.0);
                if (lane == 0) { closest[i] = c; pot_part += c; }
            }
            double pot = block_sum1(pot_part, sm);
            for (int c = 1; c < m; ++c) {
                if (tid < p.L) sm.vals[tid] = p.rand[((int64_t)it * (m - 1) + (c - 1)) * p.L + tid] * pot;
                // candidates = searchsorted(cumsum(closest), vals): the number of prefix sums below each value
                int below[kMaxL] = {0, 0, 0, 0};
                double carry = 0.0;
                __syncthreads();
                for (int base = 0; base < k; base += kThreads) {
                    const int i = base + tid;
                    double x = i < k ? closest[i] : 0.0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double y = __shfl_up_sync(0xffffffffu, x, o);
                        if (lane >= o) x += y;
                    }
                    __syncthreads();
                    if (lane == 31) sm.red[warp][0] = x;
                    __syncthreads();
                    double pre = carry;
                    for (int w = 0; w < warp; ++w) pre += sm.red[w][0];
                    const double cum = pre + x;
                    if (i < k) {
#pragma unroll
                        for (int l = 0; l < kMaxL; ++l)
                            if (l < p.L && cum < sm.vals[l]) below[l]++;
                    }
                    double tot = carry;
                    for (int w = 0; w < kWarps; ++w) tot += sm.red[w][0];
                    carry = tot;
                }
                {
                    double bl[kMaxL];
#pragma unroll
                    for (int l = 0; l < kMaxL; ++l) bl[l] = (double)below[l];
                    block_sum<kMaxL>(bl, sm);
                    if (tid < p.L) sm.cand[tid] = min((int)bl[tid], k - 1);
                }
                __syncthreads();
                // potential of every candidate
                double pp[kMaxL] = {0.0, 0.0, 0.0, 0.0};
                for (int i = warp; i < k; i += kWarps) {
                    double dots[kMaxL];
                    warp_dots<kMaxL>(n, i, p.L,
                                     [&](int v, int d) { return (double)n.X[(int64_t)sm.cand[v] * D + d] - mean[d]; }, dots);
                    if (lane == 0) {
#pragma unroll
                        for (int l = 0; l < kMaxL; ++l)
                            if (l < p.L) pp[l] += fmin(closest[i], fmax(xsq[i] - 2.0 * dots[l] + xsq[sm.cand[l]], 0.0));
                    }
                }
                block_sum<kMaxL>(pp, sm);
                int b = 0;
                for (int l = 1; l < p.L; ++l)
                    if (pp[l] < pp[b]) b = l;
                const int ib = sm.cand[b];
                pot = pp[b];
                for (int d = tid; d < D; d += kThreads) cur[(int64_t)c * D + d] = (double)n.X[(int64_t)ib * D + d] - mean[d];
                __syncthreads();
                for (int i = warp; i < k; i += kWarps) {
                    double dots[1];
                    warp_dots<1>(n, i, 1, [&](int, int d) { return cur[(int64_t)c * D + d]; }, dots);
                    if (lane == 0) closest[i] = fmin(closest[i], fmax(xsq[i] - 2.0 * dots[0] + xsq[ib], 0.0));
                }
                __syncthreads();
            }
            // ---- Lloyd ---------------------------------------------------------------------------------------
            bool strict = false, first_iter = true;
            for (int iter = 0; iter < kMaxIter; ++iter) {
                if (!first_iter) { int* t = labels; labels = lab_old; lab_old = t; }  // labels_old[:] = labels
                const bool changed = e_step(n, cur, labels, first_iter ? nullptr : lab_old, sm);
                first_iter = false;
                // member counts
                {
                    double cw[kMaxM];
#pragma unroll
                    for (int j = 0; j < kMaxM; ++j) cw[j] = 0.0;
                    for (int i = tid; i < k; i += kThreads) {
                        const int l = labels[i];
#pragma unroll
                        for (int j = 0; j < kMaxM; ++j) cw[j] += (l == j) ? 1.0 : 0.0;
                    }
                    block_sum<kMaxM>(cw, sm);
                    if (tid < m) sm.w[tid] = cw[tid];
                }
                // member sums (thread = feature, sequential over the examples)
                for (int d = tid; d < D; d += kThreads) {
                    double s[kMaxM];
#pragma unroll
                    for (int j = 0; j < kMaxM; ++j) s[j] = 0.0;
                    const double mu = mean[d];
                    for (int i = 0; i < k; ++i) {
                        const double x = (double)n.X[(int64_t)i * D + d] - mu;
                        const int l = labels[i];
#pragma unroll
                        for (int j = 0; j < kMaxM; ++j) s[j] += (l == j) ? x : 0.0;
                    }
#pragma unroll
                    for (int j = 0; j < kMaxM; ++j)
                        if (j < m) nxt[(int64_t)j * D + d] = s[j];
                }
                __syncthreads();
                // empty clusters -> the farthest points (_relocate_empty_clusters_dense)
                int n_empty = 0;
                for (int j = 0; j < m; ++j) n_empty += sm.w[j] == 0.0;
                if (n_empty) {
                    for (int i = warp; i < k; i += kWarps) {
                        const double dd = warp_dist_direct(n, i, cur + (int64_t)labels[i] * D);
                        if (lane == 0) dist[i] = dd;
                    }
                    __syncthreads();
                    double mx = -1.0;
                    int mi = 0x7fffffff;
                    for (int i = tid; i < k; i += kThreads)
                        if (dist[i] > mx) { mx = dist[i]; mi = i; }
                    block_argmax(mx, mi, sm);
                    if (mx > 0.0) {
                        for (int e = 0; e < m; ++e) {
                            if (sm.w[e] != 0.0) continue;  // uniform (shared memory, read after a barrier)
                            double fv = -1.0;
                            int far = 0x7fffffff;
                            for (int i = tid; i < k; i += kThreads)
                                if (dist[i] > fv) { fv = dist[i]; far = i; }
                            block_argmax(fv, far, sm);
                            const int o = labels[far];
                            for (int d = tid; d < D; d += kThreads) {
                                const double x = (double)n.X[(int64_t)far * D + d] - mean[d];
                                nxt[(int64_t)o * D + d] -= x;
                                nxt[(int64_t)e * D + d] = x;
                            }
                            __syncthreads();
                            if (tid == 0) {
                                sm.w[e] = 1.0;
                                sm.w[o] -= 1.0;
                                dist[far] = -1.0;  // taken
                            }
                            __syncthreads();
                        }
                    }
                }
                // _average_centers: still-empty clusters go to the biggest cluster's row, read in index order
                int amax = 0;
                for (int j = 1; j < m; ++j)
                    if (sm.w[j] > sm.w[amax]) amax = j;
                double shift_part = 0.0;
                for (int d = tid; d < D; d += kThreads) {
                    for (int j = 0; j < m; ++j) {
                        double v = nxt[(int64_t)j * D + d];
                        v = sm.w[j] > 0.0 ? v / sm.w[j] : nxt[(int64_t)amax * D + d];
                        nxt[(int64_t)j * D + d] = v;
                        const double t = v - cur[(int64_t)j * D + d];
                        shift_part = fma(t, t, shift_part);
                    }
                }
                const double shift = block_sum1(shift_part, sm);
                { double* t = cur; cur = nxt; nxt = t; }
                if (!changed) { strict = true; break; }
                if (shift <= tol) break;
            }
            if (!strict) e_step(n, cur, labels, nullptr, sm);
            double in_part = 0.0;
            for (int i = warp; i < k; i += kWarps) {
                const double dd = warp_dist_direct(n, i, cur + (int64_t)labels[i] * D);
                if (lane == 0) in_part += dd;
            }
            const double inertia = block_sum1(in_part, sm);
            // KMeans.fit: keep the run unless it is not better or is the same clustering as the best one
            bool take = !have_best;
            if (have_best && inertia < best_inertia) {
                if (tid < kMaxM) sm.mask[tid] = 0u;
                __syncthreads();
                for (int i = tid; i < k; i += kThreads) atomicOr(&sm.mask[labels[i]], 1u << best_lab[i]);
                __syncthreads();
                bool same = true;
                for (int j = 0; j < m; ++j) same &= __popc(sm.mask[j]) <= 1;
                take = !same;
                __syncthreads();
            }
            if (take) {
                for (int i = tid; i < k; i += kThreads) best_lab[i] = labels[i];
                for (int e = tid; e < m * D; e += kThreads) best_c[e] = cur[e];
                best_inertia = inertia;
                have_best = true;
            }
            __syncthreads();
        }

        // ---- score --------------------------------------------------------------------------------------------
        double result = __longlong_as_double(0x7FF8000000000000ll);
        if (!nonfinite) {
            if (tid < kMaxM) sm.cnt[tid] = 0;
            __syncthreads();
            for (int i = tid; i < k; i += kThreads) atomicAdd(&sm.cnt[best_lab[i]], 1);
            __syncthreads();
            int cmin = sm.cnt[0];
            for (int j = 1; j < m; ++j) cmin = min(cmin, sm.cnt[j]);
            if (p.replace_empty && cmin < 2) {
                // 1 - mean_{i < min(10, k)} cos(mean_k V, V[:, i])   (scores.py:178-184: clarity of two vectors is their cosine)
                const int ns = min(10, k);
                double mm = 0.0;
                for (int d = tid; d < D; d += kThreads) mm = fma(mean[d], mean[d], mm);
                mm = block_sum1(mm, sm);
                double acc = 0.0;
                for (int i = warp; i < ns; i += kWarps) {
                    const float* row = n.X + (int64_t)i * D;
                    double dv = 0.0, vv = 0.0;
                    for (int d = lane; d < D; d += 32) {
                        const double x = (double)row[d];
                        dv = fma(x, mean[d], dv);
                        vv = fma(x, x, vv);
                    }
                    dv = warp_sum(dv);
                    vv = warp_sum(vv);
                    if (lane == 0) acc += dv / (fmax(sqrt(mm), 1e-12) * fmax(sqrt(vv), 1e-12));
                }
                result = 1.0 - block_sum1(acc, sm) / (double)ns;
            } else {
                // Gram matrix of the un-centred centres; clarity = ((|mean_j c_j / |c_j||^2 - 1/m) / (m - 1)) m
                __shared__ double cgs[kMaxM * kMaxM];
                for (int pr = warp; pr < m * m; pr += kWarps) {
                    const int a = pr / m, b = pr % m;
                    double s = 0.0;
                    for (int d = lane; d < D; d += 32)
                        s = fma(best_c[(int64_t)a * D + d] + mean[d], best_c[(int64_t)b * D + d] + mean[d], s);
                    s = warp_sum(s);
                    if (lane == 0) cgs[pr] = s;
                }
                __syncthreads();
                double s = 0.0;
                for (int a = 0; a < m; ++a)
                    for (int b = 0; b < m; ++b)
                        s += cgs[a * m + b] / (fmax(sqrt(cgs[a * m + a]), 1e-12) * fmax(sqrt(cgs[b * m + b]), 1e-12));
                s /= (double)m * (double)m;
                result = 1.0 - (s - 1.0 / (double)m) / (double)(m - 1) * (double)m;
                __syncthreads();
            }
        }
        if (tid == 0) p.out[neuron] = result;
        __syncthreads();
    }
}

int64_t km_ws_doubles(int64_t k, int64_t D, int64_t m) {
    // mean[D] + 3 centre sets [m][D] + xsq / closest / dist [k] + three label arrays [k] ints (as doubles, rounded up)
    return D + 3 * m * D + 3 * k + (3 * k + 1) / 2 + 2;
}

int km_grid(int64_t C) { return (int)std::min<int64_t>(C, (int64_t)slb_sm_count() * 4); }

}  // namespace

extern "C" size_t slb_polysem_kmeans_workspace_bytes(int64_t C, int64_t k, int64_t D, int n_clusters, int n_init) {
    if (C <= 0 || k <= 0 || D <= 0 || n_clusters < 2 || n_clusters > kMaxM || n_init < 1) return 0;
    const size_t draws = (size_t)n_init * sizeof(int) + (size_t)n_init * (n_clusters - 1) * kMaxL * sizeof(double) + 64;
    return (size_t)km_grid(C) * (size_t)km_ws_doubles(k, D, n_clusters) * sizeof(double) + draws;
}

extern "C" int slb_polysem_kmeans(const float* V, int64_t C, int64_t k, int64_t D, int n_clusters, const int64_t* first_centers,
                                  const double* local_trial_uniforms, int n_local_trials, int n_init,
                                  int replace_empty_clusters, double* out, void* workspace, size_t workspace_bytes, void* stream) {
    SLB_REQUIRE(C >= 0 && k > 0 && D > 0, SLB_EINVAL, "slb_polysem_kmeans: bad size");
    if (C == 0) return SLB_OK;
    SLB_REQUIRE(V && out && workspace && first_centers && local_trial_uniforms, SLB_EINVAL,
                "slb_polysem_kmeans: null pointer (first_centers / local_trial_uniforms are HOST arrays)");
    SLB_REQUIRE(n_clusters >= 2 && n_clusters <= kMaxM, SLB_EUNSUPPORTED, "slb_polysem_kmeans: 2 <= n_clusters <= %d (got %d)", kMaxM,
                n_clusters);
    SLB_REQUIRE(k >= n_clusters, SLB_EINVAL, "n_samples=%lld should be >= n_clusters=%d.", (long long)k, n_clusters);
    SLB_REQUIRE(n_local_trials >= 1 && n_local_trials <= kMaxL, SLB_EUNSUPPORTED, "slb_polysem_kmeans: 1 <= n_local_trials <= %d", kMaxL);
    SLB_REQUIRE(n_init >= 1 && n_init <= 64, SLB_EUNSUPPORTED, "slb_polysem_kmeans: 1 <= n_init <= 64");
    SLB_REQUIRE(k < (1ll << 30) && D < (1ll << 30), SLB_EUNSUPPORTED, "slb_polysem_kmeans: k / D too large");
    SLB_REQUIRE(((uintptr_t)workspace % 16) == 0, SLB_EINVAL, "slb_polysem_kmeans: workspace must be 16-byte aligned");
    const size_t need = slb_polysem_kmeans_workspace_bytes(C, k, D, n_clusters, n_init);
    SLB_REQUIRE(workspace_bytes >= need, SLB_EWORKSPACE, "slb_polysem_kmeans: workspace needs %zu bytes, got %zu", need, workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SlbProfScope prof("K8g polysem_kmeans", stream, 0.0, 4.0 * (double)C * (double)k * (double)D);
    KmParams p{};
    p.V = V; p.C = C; p.k = (int)k; p.D = (int)D; p.m = n_clusters; p.n_init = n_init; p.L = n_local_trials;
    p.replace_empty = replace_empty_clusters ? 1 : 0;
    p.ws_doubles = km_ws_doubles(k, D, n_clusters);
    p.ws = static_cast<double*>(workspace);
    // the host draws travel in the tail of the workspace (stream-ordered copies from a staging copy the call owns)
    unsigned char* tail = static_cast<unsigned char*>(workspace) + (size_t)km_grid(C) * (size_t)p.ws_doubles * sizeof(double);
    double* d_rand = reinterpret_cast<double*>(tail);
    int* d_first = reinterpret_cast<int*>(tail + (size_t)n_init * (n_clusters - 1) * kMaxL * sizeof(double));
    int h_first[64];
    for (int i = 0; i < n_init; ++i) {
        SLB_REQUIRE(first_centers[i] >= 0 && first_centers[i] < k, SLB_EINVAL, "slb_polysem_kmeans: first centre out of range");
        h_first[i] = (int)first_centers[i];
    }
    // pageable host -> device copies return once the source has been staged, so the caller's arrays may go away after the call
    SLB_CUDA_OK(cudaMemcpyAsync(d_rand, local_trial_uniforms, (size_t)n_init * (n_clusters - 1) * n_local_trials * sizeof(double),
                                cudaMemcpyHostToDevice, st));
    SLB_CUDA_OK(cudaMemcpyAsync(d_first, h_first, (size_t)n_init * sizeof(int), cudaMemcpyHostToDevice, st));
    p.first = d_first;
    p.rand = d_rand;
    p.out = out;
    polysem_kmeans_kernel<<<km_grid(C), kThreads, 0, st>>>(p);
    SLB_LAUNCH_OK("polysem_kmeans");
    return SLB_OK;
}
