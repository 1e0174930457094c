O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_collect_gpu.py -q -p no:cacheprovider -k full_size > $O/r01o_pytest.log 2>&1; tail -5 $O/r01o_pytest.log
timeout 900 python scripts/bench_cfg5.py > $O/r01o_cfg5.jsonl 2> $O/r01o_cfg5.err; cat $O/r01o_cfg5.jsonl; tail -5 $O/r01o_cfg5.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:polysem -c 1 -f -o $O/r01o_prof_polysem python scripts/bench_cfg5.py --neurons 1184 --cpu-neurons 2 > $O/r01o_ncu_polysem.log 2>&1
