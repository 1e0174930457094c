#!/bin/bash
# full GPU suite + tower timings after an attention-kernel change. gpurun --timeout 900 -- 'bash scripts/gpu_ab_attn.sh tag'
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log | cut -c1-300
timeout 100 python scripts/profile_tower.py ViT-B-32 256 2>&1 | tail -1 | tee -a $O/${TAG}_vitb32.jsonl | cut -c1-420
timeout 100 python scripts/profile_tower.py ViT-L-14 64 2>&1 | tail -1 | tee -a $O/${TAG}_vitl14.jsonl | cut -c1-420
