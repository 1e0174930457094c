O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r01t_pytest.log 2>&1; echo "exit $?" >> $O/r01t_pytest.log; tail -5 $O/r01t_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/r01t_bench.json 2> $O/r01t_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r01t_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
for k,v in d["roofline"]["kernels"].items(): print(" ", k, {a: round(b,3) for a,b in v.items()})
PY
