#!/bin/bash
# K8 parity suites + the per-phase timing split at cfg-5 size. gpurun --timeout 900 -- 'bash scripts/gpu_ab_polysem.sh tag'
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_scores_gpu.py tests/test_cfg5_gpu.py -m gpu -q -x -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log | cut -c1-300
timeout 300 python scripts/time_polysem.py 65536 2>&1 | tee $O/${TAG}_polysem_split.jsonl | cut -c1-400
