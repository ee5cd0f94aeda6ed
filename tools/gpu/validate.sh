#!/bin/bash
# What a round-end check on a B200 box runs (through gpurun): GPU parity tests, smoke, the two bench lines.
# Usage: gpurun --timeout 2400 -- 'bash tools/gpu/validate.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --workload C2 > gpurun_out/bench_c2_n1.json 2> gpurun_out/bench_c2_n1.err
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
