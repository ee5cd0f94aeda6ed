#!/bin/bash
# round-1 re-entry check: GPU parity, self split probe, default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_self.py 10000 2048 > gpurun_out/probe_self_10k.log 2>&1
timeout 300 python tools/probe_self.py 50000 1024 > gpurun_out/probe_self_50k.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo done
