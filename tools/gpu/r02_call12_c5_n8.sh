#!/bin/bash
# round 2, 8-GPU call: BASELINE configs[4] at full size (500k atoms x 50k frames, 300 GB streamed from pinned host memory,
# atoms sharded by ModAssignment over 8 GPUs, NCCL all-reduce of the packed partials), 2 |q| per step
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,memory.total --format=csv; free -g; nproc; } > gpurun_out/box_n8.txt 2>&1
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload C5 --nq 2 --steps 1 --warmup 1 ) > gpurun_out/bench_c5_full_n8.json 2> gpurun_out/bench_c5_full_n8.err
echo "exit $?" >> gpurun_out/bench_c5_full_n8.err
free -g >> gpurun_out/box_n8.txt
