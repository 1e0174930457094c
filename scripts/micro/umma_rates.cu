// Microbenchmark: issue-to-complete cost of tcgen05.mma kind::f16 (M = 128, K = 16) per instruction, by operand source
// (SS = A from shared memory, TS = A from tensor memory) and N, with one or two CTAs per SM. Informs the attention kernel
// (attention_ts.cu): which MMA shapes run at the 128*N/256-clock floor and which are bound by operand delivery.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rates umma_rates.cu && ./umma_rates
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t idesc_f16(int M, int N, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    if (b_mn) d |= 1u << 16;
    return d;
}

// MODE 0: SS, B K-major; 1: TS, B K-major; 2: TS, B MN-major
template <int MODE>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int n_mma, int N, int iters, int cols) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if (cols == 512) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        else asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint64_t da = desc_sw128(smem_u32(smem), 16);
        const uint64_t db = desc_sw128(smem_u32(smem) + 16384, MODE == 2 ? 16384 : 16);
        const uint32_t id = idesc_f16(128, N, MODE == 2);
        long long best = 1ll << 60;
        for (int it = 0; it < iters; ++it) {
            const long long t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const uint32_t d = tmem + ((i & 1) && N <= 64 ? 64 : 0);  // alternate accumulators when they fit
                if (MODE == 0) {
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d),
                                 "l"(da + 2 * (i & 3)), "l"(db + 2 * (i & 3)), "r"(id), "r"(1u)
                                 : "memory");
                } else {
                    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p; }" ::"r"(d),
                                 "r"(tmem + (cols == 512 ? 256 : 128) + 8 * (i & 7)), "l"(db + (MODE == 2 ? 128 * (i & 7) : 2 * (i & 3))), "r"(id), "r"(1u)
                                 : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(smem_u32(&bar)), "r"((uint32_t)(it & 1)) : "memory");
            }
            const long long dt = clock64() - t0;
            if (dt < best) best = dt;
        }
        out[blockIdx.x] = best;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        if (cols == 512) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
    }
}

template <int MODE>
void run(const char* name, int N, int ctas_per_sm) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * ctas_per_sm;
    long long* out;
    cudaMallocManaged(&out, sizeof(long long) * grid);
    const int smem = ctas_per_sm == 1 ? 120 * 1024 : 64 * 1024;  // one CTA per SM: too large to share the SM
    const int cols = ctas_per_sm == 1 ? 512 : 256;
    cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    long long r[2];
    for (int k = 0; k < 2; ++k) {
        const int n_mma = k == 0 ? 64 : 576;
        rate_kernel<MODE><<<grid, 128, smem>>>(out, n_mma, N, 6, cols);
        cudaDeviceSynchronize();
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
        r[k] = mx;
    }
    printf("{\"mode\": \"%s\", \"N\": %d, \"ctas_per_sm\": %d, \"clk_per_mma\": %.1f, \"floor\": %d, \"fixed_clk\": %.0f, \"err\": \"%s\"}\n", name, N,
           ctas_per_sm, (r[1] - r[0]) / 512.0, 128 * N / 256, r[0] - 64 * (r[1] - r[0]) / 512.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int c : {1, 2}) {
        for (int N : {64, 128, 256}) {
            if (c == 2 && N == 256) continue;
            run<0>("SS", N, c);
            run<1>("TS", N, c);
            if (N <= 128) run<2>("TS, B MN-major", N, c);
        }
    }
    return 0;
}
