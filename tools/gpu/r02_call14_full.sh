#!/bin/bash
# round 2: all GPU tests, smoke, the default bench invocation at full size (timed), reference arm
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --durations=8 ) > gpurun_out/pytest_gpu_d.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_d.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke_d.log 2>&1
( time timeout 1500 python bench.py ) > gpurun_out/bench_all_n1_e.json 2> gpurun_out/bench_all_n1_e.err
