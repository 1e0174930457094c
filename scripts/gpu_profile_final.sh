#!/bin/bash
# End-of-round profile refresh: ncu launch list + DRAM traffic of the cfg 2 bench step, --set full of the short-sequence
# tcgen05 attention and the LUT normalise at the bench shape.
TAG=${1:-fin}
O=gpurun_out; mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --configs none > $O/${TAG}_ncu_launches.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'gemm_split|attention_|layernorm|patchify|assemble|u8_norm|agg_rows|agg_btf|topk_' -c 1200 --csv --log-file $O/${TAG}_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --configs none > $O/${TAG}_ncu_traffic.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attention_ts|u8_norm' -s 14 -c 3 -f -o $O/${TAG}_prof_attn_short python scripts/profile_tower.py ViT-B-32 256 > $O/${TAG}_ncu_attn.log 2>&1
ls -la $O | grep ${TAG}
