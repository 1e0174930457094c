O=gpurun_out; mkdir -p $O
for v in "" single pair128 pair192 pair256; do SLB_GEMM_KERNEL=$v timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -1; done
SLB_BENCH_ONLY=vit timeout 300 python scripts/bench_kernels.py gemm 2>&1 | grep -v '"passes": 1' | cut -c28-200
SLB_GEMM_KERNEL=single SLB_BENCH_ONLY=vitb32 timeout 300 python scripts/bench_kernels.py gemm 2>&1 | grep -v '"passes": 1' | grep -v cosine | cut -c28-200
timeout 600 python scripts/bench_kernels.py embed 2>&1 | cut -c1-300
