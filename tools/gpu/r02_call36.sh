#!/bin/bash
# kernel A coordinate prefetch, head/tail floats behind a warp-uniform test: parity, timings, racecheck
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -q -x -k "split or self" ) > gpurun_out/pytest_self12.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_self12.log
{
timeout 60 python tools/probe_self_ab.py 10000 2048 v2
timeout 60 python tools/probe_self_ab.py 50000 256 v2
timeout 60 python tools/probe_self_ab.py 26000 256 v2
timeout 60 python tools/probe_self_ab.py 10001 2048 v2
} > gpurun_out/self_ab12.log 2>&1
{
echo "--- racecheck: python tools/probe_self_ab.py 10001 2 v2"
timeout 100 compute-sanitizer --tool racecheck python tools/probe_self_ab.py 10001 2 v2 2>&1 | grep -E "=========|v2" | tail -6
} > gpurun_out/sanitizer_a3.log 2>&1
