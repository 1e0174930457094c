O=gpurun_out; mkdir -p $O
PYTHONUNBUFFERED=1 timeout 400 python -u -m pytest tests/test_gemm_gpu.py -x -q -p no:cacheprovider > $O/r02d_pytest_gemm.log 2>&1; echo "exit $?" >> $O/r02d_pytest_gemm.log; tail -12 $O/r02d_pytest_gemm.log | cut -c1-200
if grep -q "exit 0" $O/r02d_pytest_gemm.log; then
  for v in single pair128 pair256; do SLB_GEMM_KERNEL=$v timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -1; done
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02d_pytest.log 2>&1; echo "exit $?" >> $O/r02d_pytest.log; tail -12 $O/r02d_pytest.log | cut -c1-200
  for v in single pair128 pair256; do
    SLB_GEMM_KERNEL=$v SLB_BENCH_ONLY=vit timeout 300 python scripts/bench_kernels.py gemm > $O/r02d_gemm_$v.jsonl 2>&1
    SLB_GEMM_KERNEL=$v SLB_BENCH_ONLY=square timeout 300 python scripts/bench_kernels.py gemm >> $O/r02d_gemm_$v.jsonl 2>&1
    echo "== $v"; grep -v '"passes": 1' $O/r02d_gemm_$v.jsonl | grep -v cosine | cut -c28-200
  done
fi
