#!/bin/bash
# round 2, fourth GPU call: all GPU tests, the default bench invocation at full size (timed), K3 probe
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 1500 python bench.py ) > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err
{
timeout 300 python tools/probe_k3.py
NM=25000 NF=1000 timeout 300 python tools/probe_k3.py
} > gpurun_out/k3_probe.log 2>&1
