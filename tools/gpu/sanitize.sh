#!/bin/bash
# compute-sanitizer memcheck + racecheck over the self split path and the multipole kernels (small parity cases).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu/sanitize.sh'
mkdir -p gpurun_out
( time timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_path and (10000 or 5000 or 8193 or 33000) or test_mpsphere or test_self_split_layout" ) > gpurun_out/memcheck.log 2>&1
echo "exit $?" >> gpurun_out/memcheck.log
( time timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split_path and (10000 or 5000) or test_mpsphere_batch" ) > gpurun_out/racecheck.log 2>&1
echo "exit $?" >> gpurun_out/racecheck.log
