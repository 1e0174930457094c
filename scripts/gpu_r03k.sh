O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_embed_gpu.py -m gpu -q -p no:cacheprovider -k "attention_from_planes" 2>&1 | tail -12 | cut -c1-160
SLB_ATTN_TRACE=1 timeout 120 python scripts/trace_attention.py 257 1 1
timeout 300 python scripts/profile_tower.py ViT-L-14 64 2>&1 | cut -c1-420 | tail -2
