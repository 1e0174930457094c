O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 10 --warmup 3 > $O/r01l_bench.json 2> $O/r01l_bench.err; tail -c 2500 $O/r01l_bench.json; tail -3 $O/r01l_bench.err
# DRAM traffic + duration of every libslb200 launch of two bench steps (no e2e / cpu legs)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'gemm_split|attention_|layernorm|patchify|assemble|u8_norm|agg_rows|agg_btf|topk_' -c 1200 --csv --log-file $O/r01l_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/r01l_ncu_traffic.log 2>&1
# full captures: one of each hot kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_split_kernel|attention_mma|agg_rows_group|agg_rows_subwarp' -s 40 -c 12 -f -o $O/r01l_prof_hot python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $O/r01l_ncu_hot.log 2>&1
ls -la $O | grep r01l
