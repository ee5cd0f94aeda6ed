#!/bin/bash
# 4 GPUs: the default bench invocation (every workload, incl. the vector-sharded coherent one)
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 ) > gpurun_out/bench_all_n4.json 2> gpurun_out/bench_all_n4.err
echo "exit $?" >> gpurun_out/bench_all_n4.err
