#!/bin/bash
# attention parity + ViT-B/32 tower A/B of the slot-packed short-sequence path. gpurun --timeout 900 -- 'bash scripts/gpu_ab_attn.sh tag'
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_embed_gpu.py -m gpu -q -x -p no:cacheprovider -k "attention" > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log | cut -c1-300
for m in 0 1 0 1; do
    echo "SLB_ATTN_PACK=$m"; SLB_ATTN_PACK=$m timeout 120 python scripts/profile_tower.py ViT-B-32 256 2>&1 | tail -1 | tee -a $O/${TAG}_vitb32.jsonl | cut -c1-420
done
