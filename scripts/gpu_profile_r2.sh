#!/bin/bash
# Round-2 profiling call: ncu launch list of the bench step, DRAM traffic per kernel, --set full captures of the pair GEMM at
# the bench shape (M = 12 800), the polysemanticity kernel, the implicit-GEMM convolutions and the channels-last aggregation.
TAG=${1:-r2x}
O=gpurun_out; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --configs none > $O/${TAG}_ncu_launches.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'gemm_split|attention_|layernorm|patchify|assemble|u8_norm|agg_rows|agg_btf|topk_' -c 1200 --csv --log-file $O/${TAG}_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --configs none > $O/${TAG}_ncu_traffic.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 150 -c 5 -f -o $O/${TAG}_prof_gemm_bench python scripts/profile_tower.py ViT-B-32 256 > $O/${TAG}_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:polysem -s 1 -c 1 -f -o $O/${TAG}_prof_polysem python scripts/ncu_polysem.py > $O/${TAG}_ncu_polysem.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_split|im2col|avgpool' -s 120 -c 12 -f -o $O/${TAG}_prof_rn python scripts/profile_tower.py RN50 128 > $O/${TAG}_ncu_rn.log 2>&1
ls -la $O | grep ${TAG}
