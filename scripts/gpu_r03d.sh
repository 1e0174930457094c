O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_rn_gpu.py tests/test_embed_gpu.py tests/test_scores_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r03d_pytest.log 2>&1; echo "exit $?" >> $O/r03d_pytest.log
tail -5 $O/r03d_pytest.log | cut -c1-220
for v in single pair128 pair256; do SLB_GEMM_KERNEL=$v timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -1; done
timeout 600 python scripts/bench_rn_convs.py 128 > $O/r03d_rn_convs.jsonl 2>&1; grep split_acc $O/r03d_rn_convs.jsonl | cut -c1-220
timeout 600 python scripts/bench_kernels.py embed > $O/r03d_embed.jsonl 2>&1; cut -c1-400 $O/r03d_embed.jsonl
SLB_BENCH_ONLY=vit timeout 300 python scripts/bench_kernels.py gemm > $O/r03d_gemm.jsonl 2>&1; grep -v '"passes": 1' $O/r03d_gemm.jsonl | cut -c28-200
