O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r01q_pytest.log 2>&1; echo "exit $?" >> $O/r01q_pytest.log; tail -15 $O/r01q_pytest.log
timeout 300 python scripts/bench_kernels.py collect > $O/r01q_micro_collect.jsonl 2>&1; grep '"K1"' $O/r01q_micro_collect.jsonl | cut -c1-190
