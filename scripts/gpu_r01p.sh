O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_embed_gpu.py -q -p no:cacheprovider > $O/r01p_pytest.log 2>&1; echo "exit $?" >> $O/r01p_pytest.log; tail -12 $O/r01p_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/r01p_smoke.log 2>&1; tail -3 $O/r01p_smoke.log
timeout 400 python scripts/bench_kernels.py embed > $O/r01p_micro_embed.jsonl 2>&1; cat $O/r01p_micro_embed.jsonl
