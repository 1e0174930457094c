O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_configs_gpu.py -q -p no:cacheprovider > $O/r01s_pytest.log 2>&1; tail -12 $O/r01s_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r01s_bench.json 2> $O/r01s_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r01s_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"])
PY
