#!/bin/bash
# attention parity + ViT-L/14 tower A/B of the SIMT key tail. gpurun --timeout 900 -- 'bash scripts/gpu_ab_attn.sh tag'
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_embed_gpu.py -m gpu -q -x -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -8 $O/${TAG}_pytest.log | cut -c1-300
for m in 0 1 0 1; do
    echo "SLB_ATTN_KEY_TAIL=$m"; SLB_ATTN_KEY_TAIL=$m timeout 120 python scripts/profile_tower.py ViT-L-14 64 2>&1 | tail -1 | tee -a $O/${TAG}_vitl14.jsonl | cut -c1-420
done
