#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "self" ) > gpurun_out/pytest_self.log 2>&1
timeout 300 python tools/probe_self.py 10000 2048 > gpurun_out/probe_self_10k_v2.log 2>&1
timeout 300 python tools/probe_self.py 50000 1024 > gpurun_out/probe_self_50k_v2.log 2>&1
timeout 300 python tools/probe_self.py 7000 2048 > gpurun_out/probe_self_7k_v2.log 2>&1
SELF_NA=96 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_self_v2.csv python tools/probe_self.py 10000 96 > /dev/null 2>&1
echo done
