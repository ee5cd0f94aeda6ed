#!/bin/bash
# 2 GPUs: config 3 with the directions of the scan sharded over the GPUs (every rank holds all frames)
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload C3 --shard vectors --steps 2 --warmup 3 --no-cpu ) > gpurun_out/bench_c3_vec_n2.json 2> gpurun_out/bench_c3_vec_n2.err
echo "exit $?" >> gpurun_out/bench_c3_vec_n2.err
