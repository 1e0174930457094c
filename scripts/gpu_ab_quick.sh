#!/bin/bash
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
for i in 1 2; do
  for m in 8 0; do
    echo "SLB_GEMM_EPI_WARPS=$m"; SLB_GEMM_EPI_WARPS=$m timeout 120 python scripts/profile_tower.py RN50 128 2>&1 | tail -1 | tee -a $O/${TAG}_rn50.jsonl | cut -c1-250
  done
done
