O=gpurun_out; mkdir -p $O
SLB_GEMM_KERNEL=pair192 timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_embed_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_embed_gpu.py tests/test_scores_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -1
for v in single pair192 pair256; do
  SLB_GEMM_KERNEL=$v SLB_BENCH_ONLY=vitb32 timeout 300 python scripts/bench_kernels.py gemm > $O/r03f_gemm_$v.jsonl 2>&1
  echo "== $v"; grep -v '"passes": 1' $O/r03f_gemm_$v.jsonl | grep -v cosine | cut -c28-200
done
echo "== default"; SLB_BENCH_ONLY=vit timeout 300 python scripts/bench_kernels.py gemm 2>&1 | grep -v '"passes": 1' | cut -c28-200
timeout 600 python scripts/bench_kernels.py embed > $O/r03f_embed.jsonl 2>&1; cut -c1-400 $O/r03f_embed.jsonl
