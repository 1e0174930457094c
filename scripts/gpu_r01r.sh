O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_collect_gpu.py tests/test_scores_gpu.py -q -p no:cacheprovider > $O/r01r_pytest.log 2>&1; tail -4 $O/r01r_pytest.log
timeout 300 python scripts/bench_kernels.py collect > $O/r01r_micro_collect.jsonl 2>&1; grep 'vit_b16' $O/r01r_micro_collect.jsonl | cut -c1-190
timeout 600 python scripts/bench_cfg5.py --cpu-neurons 16 > $O/r01r_cfg5.jsonl 2>$O/r01r_cfg5.err; grep polysem $O/r01r_cfg5.jsonl | cut -c1-300
