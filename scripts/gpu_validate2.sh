#!/bin/bash
# Two-GPU validation (gpurun --gpus 2): the two-rank NCCL equality tests and the bench with all sub-records under torchrun.
TAG=${1:-val2}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log | cut -c1-300
t2=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > $O/${TAG}_bench_2gpu.json 2> $O/${TAG}_bench_2gpu.err
echo "bench rc $? $(( $(date +%s)-t2 )) s"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_2gpu.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("all_runs"), "ms/step", round(d["ms_per_step"],2))
for k,v in d["configs"].items():
    print(k, round(v.get("value",0),1), v.get("ms_per_step"), v.get("seconds"), (v.get("parity") or {}).get("ok") if isinstance(v.get("parity"),dict) else v.get("parity"))
PY
tail -3 $O/${TAG}_bench_2gpu.err | cut -c1-300
