O=gpurun_out; mkdir -p $O
timeout 240 python -m pytest tests/test_gemm_gpu.py -q -x -p no:cacheprovider --timeout 120 > $O/r01c_pytest_gemm.log 2>&1; echo "exit $?" >> $O/r01c_pytest_gemm.log; tail -15 $O/r01c_pytest_gemm.log
if grep -q "exit 0" $O/r01c_pytest_gemm.log; then
  timeout 300 python scripts/bench_kernels.py gemm > $O/r01c_micro_gemm_pair.jsonl 2> $O/r01c_micro_gemm_pair.err
  SLB_GEMM_SINGLE=1 timeout 300 python scripts/bench_kernels.py gemm > $O/r01c_micro_gemm_single.jsonl 2>&1
  timeout 600 python -m pytest tests/test_embed_gpu.py tests/test_scores_gpu.py -q -p no:cacheprovider --timeout 300 > $O/r01c_pytest_embed.log 2>&1; tail -3 $O/r01c_pytest_embed.log
  timeout 300 python scripts/bench_kernels.py embed > $O/r01c_micro_embed.jsonl 2>&1
  cat $O/r01c_micro_gemm_pair.jsonl $O/r01c_micro_embed.jsonl
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01c_launches_embed.csv python scripts/bench_kernels.py embed --batch 256 > $O/r01c_ncu_embed.log 2>&1
fi
