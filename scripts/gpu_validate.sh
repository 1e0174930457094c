#!/bin/bash
# Round validation on one B200: GPU test suite, smoke, both bench arms (what the driver runs at round end).
TAG=${1:-val}
O=gpurun_out; mkdir -p $O
t0=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log | cut -c1-300
t1=$(date +%s); echo "pytest $((t1-t0)) s"
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; tail -c 300 $O/${TAG}_bench_ref.json
t2=$(date +%s)
timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
t3=$(date +%s); echo "bench $((t3-t2)) s"
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("all_runs"), "ms/step", round(d["ms_per_step"],2), "cpu", round(d["cpu_baseline"]["value"],1))
r=d["roofline"]; print(r["kernel"], round(r["achieved"],1), round(r["frac"],3), r["traffic"])
for k,v in d["configs"].items():
    print(k, round(v.get("value",0),1), v.get("ms_per_step"), v.get("seconds"), (v.get("parity") or {}).get("ok") if isinstance(v.get("parity"),dict) else v.get("parity"))
PY
tail -3 $O/${TAG}_bench.err
