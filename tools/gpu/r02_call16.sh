#!/bin/bash
# FP32-D corrected scan kernel: scan parity tests, default bench at full size, ncu launch list of the headline workload
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "scan" ) > gpurun_out/pytest_scan2.log 2>&1
( time timeout 1500 python bench.py ) > gpurun_out/bench_all_n1_c.json 2> gpurun_out/bench_all_n1_c.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r02_launches_c3.log 2>&1
