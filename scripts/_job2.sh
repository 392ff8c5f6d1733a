N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; grep "dist_check\|Error\|error" gpurun_out/dist_check_n$N.log | tail -6
