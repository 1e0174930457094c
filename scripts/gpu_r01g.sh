O=gpurun_out; mkdir -p $O
for v in pair single bn256; do
  case $v in pair) E="";; single) E="SLB_GEMM_SINGLE=1";; bn256) E="SLB_GEMM_SINGLE=1 SLB_GEMM_BN256=1";; esac
  env $E SLB_BENCH_ONLY=square timeout 300 python scripts/bench_kernels.py gemm > $O/r01g_gemm_$v.jsonl 2>&1
  env $E SLB_BENCH_ONLY=vitb32 timeout 300 python scripts/bench_kernels.py gemm >> $O/r01g_gemm_$v.jsonl 2>&1
  echo "== $v"; grep -v '"passes": 1' $O/r01g_gemm_$v.jsonl | cut -c1-420
done
SLB_GEMM_SINGLE=1 SLB_GEMM_BN256=1 timeout 200 python -m pytest tests/test_gemm_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
