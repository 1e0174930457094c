O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_embed_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r03i_pytest.log 2>&1; echo "exit $?" >> $O/r03i_pytest.log
tail -4 $O/r03i_pytest.log | cut -c1-220
SLB_ATTN_TRACE=1 timeout 120 python scripts/trace_attention.py 257 1 1
timeout 600 python scripts/bench_kernels.py embed 2>&1 | cut -c1-300
timeout 300 python scripts/profile_tower.py ViT-L-14 64 > $O/r03i_tower_ViT-L-14.json 2>&1; cut -c1-700 $O/r03i_tower_ViT-L-14.json | tail -3
