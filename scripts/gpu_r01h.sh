O=gpurun_out; mkdir -p $O
nvidia-smi -L > $O/r01h_gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_multigpu.py > $O/r01h_check_multigpu.log 2>&1; echo "exit $?" >> $O/r01h_check_multigpu.log; tail -8 $O/r01h_check_multigpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 8 --warmup 3 > $O/r01h_bench_n2.json 2> $O/r01h_bench_n2.err; echo "exit $?"; tail -c 1500 $O/r01h_bench_n2.json; tail -3 $O/r01h_bench_n2.err
