#!/bin/bash
# A/B of the RN tower variants on one box. gpurun --timeout 600 -- 'bash scripts/gpu_ab_rn.sh tag'
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_probed_gpu.py tests/test_rn_gpu.py -m gpu -q -x -p no:cacheprovider -k "implicit or rn_ or shortcut or conv" > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -15 $O/${TAG}_pytest.log
for i in 1 2; do
  for m in 1 0; do
    echo "SLB_RN_IM2COL=$m"; SLB_RN_IM2COL=$m timeout 120 python scripts/profile_tower.py RN50 128 2>&1 | tail -1 | tee -a $O/${TAG}_rn50.jsonl | cut -c1-420
  done
done
