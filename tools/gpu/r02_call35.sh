#!/bin/bash
# kernel A coordinate prefetch without the float-by-float loop: parity, timings at aligned (10000, 50000) and unaligned
# (26000 -> R = 14, 10001) sub-sequences, racecheck + memcheck on unaligned ones
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -q -x -k "split or self" ) > gpurun_out/pytest_self11.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_self11.log
{
timeout 100 python tools/probe_self_ab.py 26000 256 v2
timeout 100 python tools/probe_self_ab.py 10001 2048 v2
timeout 100 python tools/probe_self_ab.py 10000 2048 v2
timeout 100 python tools/probe_self_ab.py 50000 256 v2
} > gpurun_out/self_ab11.log 2>&1
{
echo "--- racecheck: python tools/probe_self_ab.py 10001 2 v2"
timeout 200 compute-sanitizer --tool racecheck python tools/probe_self_ab.py 10001 2 v2 2>&1 | grep -E "=========|v2" | tail -6
echo "--- memcheck: python tools/probe_self_ab.py 26001 3 v2"
timeout 200 compute-sanitizer --tool memcheck python tools/probe_self_ab.py 26001 3 v2 2>&1 | grep -E "=========|v2" | tail -6
} > gpurun_out/sanitizer_a2.log 2>&1
