#!/bin/bash
# round 2, fifth GPU call (2 GPUs): multi-GPU parity check (native NCCL communicator, torch communicator, sharded C-ABI),
# then the default bench invocation on 2 GPUs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/box_n2_b.txt; free -g >> gpurun_out/box_n2_b.txt; nproc >> gpurun_out/box_n2_b.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multigpu_check.py ) > gpurun_out/multigpu_check_n2_b.log 2>&1
echo "exit $?" >> gpurun_out/multigpu_check_n2_b.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 ) > gpurun_out/bench_all_n2_b.json 2> gpurun_out/bench_all_n2_b.err
echo "exit $?" >> gpurun_out/bench_all_n2_b.err
