O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_embed_gpu.py -q -p no:cacheprovider > $O/r01m_pytest.log 2>&1; echo "exit $?" >> $O/r01m_pytest.log; tail -6 $O/r01m_pytest.log
timeout 300 python scripts/bench_kernels.py embed > $O/r01m_micro_embed.jsonl 2>&1; cat $O/r01m_micro_embed.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r01m_bench.json 2> $O/r01m_bench.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/r01m_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"])
r=d["roofline"]; print(r["kernel"], r["achieved"], r["frac"], r["traffic"])
for k,v in r["kernels"].items(): print(" ", k, v)
PY
