O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 > $O/r03j_bench_n4.json 2> $O/r03j_bench_n4.err; cut -c1-400 $O/r03j_bench_n4.json; tail -3 $O/r03j_bench_n4.err
