O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider > $O/r02e_pytest.log 2>&1; echo "exit $?" >> $O/r02e_pytest.log
grep -a "vs f64\|passed\|failed\|^exit\|FAILED\|AssertionError: assert" $O/r02e_pytest.log | cut -c1-200
for v in single pair128 pair256; do SLB_GEMM_KERNEL=$v timeout 300 python -m pytest tests/test_gemm_gpu.py tests/test_embed_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -1; done
timeout 600 python scripts/bench_kernels.py embed > $O/r02e_embed.jsonl 2>&1; cut -c1-400 $O/r02e_embed.jsonl
for t in "ViT-B-32 256" "ViT-L-14 64"; do timeout 300 python scripts/profile_tower.py $t > "$O/r02e_tower_${t%% *}.json" 2>&1; done
timeout 600 python bench.py > $O/r02e_bench.json 2> $O/r02e_bench.err; cut -c1-1500 $O/r02e_bench.json
