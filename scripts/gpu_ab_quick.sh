#!/bin/bash
TAG=${1:-ab}
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest tests/test_rn_gpu.py tests/test_gemm_gpu.py tests/test_gemm_modes_gpu.py tests/test_probed_gpu.py tests/test_scores_gpu.py -m gpu -q -x -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
for i in 1 2; do
    timeout 120 python scripts/profile_tower.py RN50 128 2>&1 | tail -1 | tee -a $O/${TAG}_rn50.jsonl | cut -c1-250
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --configs cfg2a,cfg4a > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
for k,v in d["configs"].items():
    print(k, round(v["value"],1), round(v["ms_per_step"],2), v["parity"]["ok"], {n:round(x["ms_per_step"],2) for n,x in v["roofline"]["kernels"].items() if x["ms_per_step"]>0.3})
PY
