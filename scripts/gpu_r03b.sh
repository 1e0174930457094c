O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_resize_gpu.py tests/test_rn_gpu.py -m gpu -q -s -p no:cacheprovider > $O/r03b_pytest_rn.log 2>&1; echo "exit $?" >> $O/r03b_pytest_rn.log
grep -a "vs f64\|passed\|failed\|^exit\|FAILED\|Error" $O/r03b_pytest_rn.log | cut -c1-220 | head -40
timeout 300 python scripts/profile_tower.py RN50 128 > $O/r03b_tower_RN50.json 2>&1; cut -c1-1500 $O/r03b_tower_RN50.json | tail -3
