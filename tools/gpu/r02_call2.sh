#!/bin/bash
# round 2, second GPU call: GPU tests incl. the new bench contract, the default bench invocation at full size (timed), ncu of the
# symmetric scan kernels
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 1500 python bench.py ) > gpurun_out/bench_all_n1.json 2> gpurun_out/bench_all_n1.err
NQ=50 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_sym -c 2 -o gpurun_out/r02_scan_plain python tools/scan_probe.py > gpurun_out/ncu_scan_plain.log 2>&1
ROUNDED=1 NQ=50 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_sym -c 3 -o gpurun_out/r02_scan_corr python tools/scan_probe.py > gpurun_out/ncu_scan_corr.log 2>&1
