// Microbenchmark: FP64 throughput of one B200 — DFMA (CUDA cores) vs DMMA (mma.sync f64 shapes). Informs the K8 Gram design.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f64_rates f64_rates.cu && ./f64_rates
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(double* out, int iters) {
    const int lane = threadIdx.x & 31;
    double a0 = 1.0 + lane * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b0 = 0.5 + lane * 1e-9, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // DFMA: 16 independent chains
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(a0, b0, c[i]);
        } else if (MODE == 1) {  // m8n8k4: 8 independent accumulator tiles (2 regs each)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a0), "d"(b0));
        } else if (MODE == 2) {  // m16n8k4
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[4 * i]), "+d"(c[4 * i + 1]), "+d"(c[4 * i + 2]), "+d"(c[4 * i + 3]) : "d"(a0), "d"(a1), "d"(b0));
        } else if (MODE == 3) {  // m16n8k8
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[4 * i]), "+d"(c[4 * i + 1]), "+d"(c[4 * i + 2]), "+d"(c[4 * i + 3])
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        } else {  // m16n8k16
#pragma unroll
            for (int i = 0; i < 4; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[4 * i]), "+d"(c[4 * i + 1]), "+d"(c[4 * i + 2]), "+d"(c[4 * i + 3])
                             : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(a4), "d"(a5), "d"(a6), "d"(a7), "d"(b0), "d"(b1), "d"(b2), "d"(b3));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter, int ctas_per_sm) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    const int grid = sms * ctas_per_sm;
    cudaMalloc(&out, sizeof(double) * grid * 256);
    const int iters = 20000;
    rate_kernel<MODE><<<grid, 256>>>(out, 100);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    rate_kernel<MODE><<<grid, 256>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double flops = 2.0 * fma_per_thread_iter * iters * 256.0 * grid;
    printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"TFLOP/s\": %.2f, \"err\": \"%s\"}\n", name, ctas_per_sm, ms,
           flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int c : {1, 2, 4}) {
        run<0>("DFMA", 16, c);
        run<1>("DMMA m8n8k4", 8 * 256 / 32.0, c);
        run<2>("DMMA m16n8k4", 4 * 512 / 32.0, c);
        run<3>("DMMA m16n8k8", 4 * 1024 / 32.0, c);
        run<4>("DMMA m16n8k16", 4 * 2048 / 32.0, c);
    }
    return 0;
}
