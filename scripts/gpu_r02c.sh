O=gpurun_out; mkdir -p $O
SLB_GEMM_KERNEL=pair256 timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -p no:cacheprovider > $O/r02c_pytest_p256.log 2>&1; tail -4 $O/r02c_pytest_p256.log
timeout 300 python -m pytest tests/test_gemm_gpu.py -q -x -p no:cacheprovider > $O/r02c_pytest_auto.log 2>&1; tail -2 $O/r02c_pytest_auto.log
for v in single pair128 pair256; do
  SLB_GEMM_KERNEL=$v SLB_BENCH_ONLY=vitl14 timeout 300 python scripts/bench_kernels.py gemm > $O/r02c_gemm_$v.jsonl 2>&1
  SLB_GEMM_KERNEL=$v SLB_BENCH_ONLY=square timeout 300 python scripts/bench_kernels.py gemm >> $O/r02c_gemm_$v.jsonl 2>&1
  SLB_GEMM_KERNEL=$v SLB_BENCH_ONLY=vitb32 timeout 300 python scripts/bench_kernels.py gemm >> $O/r02c_gemm_$v.jsonl 2>&1
  echo "== $v"; grep -v '"passes": 1' $O/r02c_gemm_$v.jsonl | grep -v cosine | cut -c28-230
done
