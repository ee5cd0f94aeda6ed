#!/bin/bash
# kernel A with exclusive chunk ownership: racecheck + memcheck on a small case, parity tests of the self path, A/B timing
mkdir -p gpurun_out
( timeout 900 compute-sanitizer --tool racecheck python tools/probe_self_ab.py 10000 2 v2 ) > gpurun_out/r02_self_racecheck.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck python tools/probe_self_ab.py 10001 3 v2 ) > gpurun_out/r02_self_memcheck.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x -k "self or stage or stream" ) > gpurun_out/pytest_self7.log 2>&1
timeout 300 python tools/probe_self_ab.py 10000 2048 fused v2 > gpurun_out/self_ab7.log 2>&1
