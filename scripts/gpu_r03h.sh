O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_embed_gpu.py tests/test_configs_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r03h_pytest.log 2>&1; echo "exit $?" >> $O/r03h_pytest.log
tail -5 $O/r03h_pytest.log | cut -c1-220
timeout 600 python scripts/bench_kernels.py embed 2>&1 | cut -c1-300
timeout 300 python scripts/profile_tower.py ViT-L-14 64 > $O/r03h_tower_ViT-L-14.json 2>&1; cut -c1-700 $O/r03h_tower_ViT-L-14.json | tail -3
