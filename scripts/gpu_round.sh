#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), kernel micro-benchmarks, ncu launch list + full captures.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
# Everything lands in gpurun_out/<tag>_*; nothing here reads /root/reference.
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/${TAG}_gpu.csv 2>&1
nproc > $O/${TAG}_nproc.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> $O/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
echo "bench exit $?"; tail -c 600 $O/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
for what in gemm embed scores collect; do
  timeout 300 python scripts/bench_kernels.py $what > $O/${TAG}_micro_$what.jsonl 2> $O/${TAG}_micro_$what.err
done
# ncu: launch list of one short bench run (shares, not absolutes), then full captures of the dominant kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > $O/${TAG}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_split -s 30 -c 4 -f -o $O/${TAG}_prof_gemm \
  python scripts/bench_kernels.py embed --batch 256 > $O/${TAG}_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'clarity|polysem' -c 3 -f -o $O/${TAG}_prof_scores \
  python scripts/bench_kernels.py scores > $O/${TAG}_ncu_scores.log 2>&1
ls -la $O | tail -30
