#!/bin/bash
# End-of-round validation on one B200: GPU test suite, smoke, bench (both arms), launch list + ncu captures of the hot kernels.
TAG=${1:-final}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log; tail -4 $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; tail -1 $O/${TAG}_smoke.log
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
timeout 600 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["e2e"].get("all_runs"), "ms/step", round(d["ms_per_step"],2), "cpu", round(d["cpu_baseline"]["value"],1))
r=d["roofline"]; print(r["kernel"], round(r["achieved"],1), round(r["frac"],3), r["traffic"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_launches.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'gemm_split|attention_|layernorm|patchify|assemble|u8_norm|agg_rows|agg_btf|topk_' -c 1200 --csv --log-file $O/${TAG}_traffic.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_traffic.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'gemm_split_pair|attention_tc|attention_planes|layernorm_reg' -s 20 -c 8 -f -o $O/${TAG}_prof_towers python scripts/profile_tower.py ViT-L-14 16 > $O/${TAG}_ncu_towers.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'im2col3x3|avgpool2|gemm_split_kernel|pool_tokens|im2col_stem' -s 6 -c 8 -f -o $O/${TAG}_prof_rn python scripts/profile_tower.py RN50 32 > $O/${TAG}_ncu_rn.log 2>&1
timeout 600 python scripts/bench_kernels.py embed > $O/${TAG}_embed.jsonl 2>&1
for t in "ViT-B-32 256" "ViT-L-14 64" "RN50 128"; do timeout 300 python scripts/profile_tower.py $t > "$O/${TAG}_tower_${t%% *}.json" 2>&1; done
ls $O | grep ${TAG}
