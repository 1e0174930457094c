#!/bin/bash
# ncu of the RN50 tower's GEMM launches: per-launch durations of one forward, and full captures of layer1's launches.
TAG=${1:-ncu}
O=gpurun_out; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'gemm_split|stem_conv|avgpool' -s 65 -c 65 --csv --log-file $O/${TAG}_rn_launches.csv python scripts/profile_tower.py RN50 128 > $O/${TAG}_ncu1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_split_kernel -s 59 -c 10 -f -o $O/${TAG}_prof_rn_layer1 python scripts/profile_tower.py RN50 128 > $O/${TAG}_ncu2.log 2>&1
ls -la $O | grep ${TAG}
