#!/bin/bash
# round 2, sixth GPU call: second design of the split self path's kernel A: parity tests, memcheck, A/B timing, ncu capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "self or stage or stream" ) > gpurun_out/pytest_self.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_self.log
( timeout 600 compute-sanitizer --tool memcheck python tools/probe_self_ab.py 10000 4 v2 ) > gpurun_out/self_v2_memcheck.log 2>&1
{
timeout 300 python tools/probe_self_ab.py 10000 2048
timeout 300 python tools/probe_self_ab.py 50000 256 v1 v2
timeout 300 python tools/probe_self_ab.py 7001 1024
} > gpurun_out/self_ab.log 2>&1
SASSENA_SELF_PATH=split timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_split -c 4 -o gpurun_out/r02_self_split_v2 python tools/probe_self.py 10000 96 > gpurun_out/ncu_self_v2.log 2>&1
ncu -i gpurun_out/r02_self_split_v2.ncu-rep --page raw --csv > gpurun_out/r02_self_split_v2_raw.csv 2>/dev/null
