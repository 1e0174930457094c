O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/r03g_bench_n2.json 2> $O/r03g_bench_n2.err; cut -c1-1200 $O/r03g_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_multigpu.py > $O/r03g_check_n2.log 2>&1; tail -5 $O/r03g_check_n2.log
