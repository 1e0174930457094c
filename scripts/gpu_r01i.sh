O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_multigpu.py > $O/r01i_check_multigpu.log 2>&1; echo "exit $?" >> $O/r01i_check_multigpu.log; grep -E "rank|parity|exit" $O/r01i_check_multigpu.log | head
