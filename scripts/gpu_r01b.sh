O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_embed_gpu.py -q -p no:cacheprovider > $O/r01b_pytest_embed.log 2>&1; tail -3 $O/r01b_pytest_embed.log
timeout 300 python scripts/bench_kernels.py embed > $O/r01b_micro_embed.jsonl 2> $O/r01b_micro_embed.err; cat $O/r01b_micro_embed.jsonl
SLB_ATTN_SIMT=1 timeout 300 python scripts/bench_kernels.py embed > $O/r01b_micro_embed_simt.jsonl 2>&1
timeout 600 python scripts/exp_probed.py > $O/r01b_exp_probed.jsonl 2> $O/r01b_exp_probed.err; cat $O/r01b_exp_probed.jsonl
