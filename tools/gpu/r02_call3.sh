#!/bin/bash
# round 2, third GPU call: all GPU tests (no -x), the streamed sample under compute-sanitizer, K3 probe with the in-SM DSP
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( CUDA_LAUNCH_BLOCKING=1 timeout 600 python bench.py --workload C5s --atoms 96 --frames 9000 --wave-atoms 40 --steps 1 --warmup 1 --cpu-seconds 0.05 ) > gpurun_out/c5s_small.json 2> gpurun_out/c5s_small.err
echo "exit $?" >> gpurun_out/c5s_small.err
( timeout 900 compute-sanitizer --tool memcheck --report-api-errors all --error-exitcode 7 python bench.py --workload C5s --atoms 96 --frames 9000 --wave-atoms 40 --steps 1 --warmup 1 --no-cpu ) > gpurun_out/c5s_sanitizer.log 2>&1
echo "exit $?" >> gpurun_out/c5s_sanitizer.log
{
timeout 300 python tools/probe_k3.py
NM=25000 NF=1000 timeout 300 python tools/probe_k3.py
NM=1000 NF=100 timeout 300 python tools/probe_k3.py
NM=500 NF=10000 timeout 300 python tools/probe_k3.py
} > gpurun_out/k3_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"self_fused|sf_reduce" -c 8 -o gpurun_out/r02_k3_insm python tools/probe_k3.py > gpurun_out/ncu_k3_insm.log 2>&1
