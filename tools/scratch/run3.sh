#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload C2 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
echo "rc $?" >> gpurun_out/bench_c2_n1.err
