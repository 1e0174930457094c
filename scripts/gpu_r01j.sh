O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r01j_pytest.log 2>&1; echo "exit $?" >> $O/r01j_pytest.log; tail -6 $O/r01j_pytest.log
timeout 300 python scripts/bench_kernels.py collect > $O/r01j_micro_collect.jsonl 2>&1; grep '"K1"' $O/r01j_micro_collect.jsonl | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r01j_bench.json 2> $O/r01j_bench.err; tail -c 1800 $O/r01j_bench.json
