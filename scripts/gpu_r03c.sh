O=gpurun_out; mkdir -p $O
timeout 600 python scripts/bench_rn_convs.py 128 > $O/r03c_rn_convs.jsonl 2>&1; cat $O/r03c_rn_convs.jsonl | cut -c1-220
