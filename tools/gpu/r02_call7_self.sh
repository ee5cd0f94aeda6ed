#!/bin/bash
# round 2, seventh GPU call: kernel A v2 after the spill / branch fixes: A/B timing, parity tests of the self path, ncu capture
mkdir -p gpurun_out
{
timeout 300 python tools/probe_self_ab.py 10000 2048 fused v1 v2
timeout 300 python tools/probe_self_ab.py 50000 256 v1 v2
timeout 300 python tools/probe_self_ab.py 4500 1024 fused v2
} > gpurun_out/self_ab2.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x -k "self or stage or stream" ) > gpurun_out/pytest_self2.log 2>&1
SASSENA_SELF_PATH=split timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_split_fft -c 1 -o gpurun_out/r02_self_split_v2b python tools/probe_self.py 10000 96 > gpurun_out/ncu_self_v2b.log 2>&1
ncu -i gpurun_out/r02_self_split_v2b.ncu-rep --page raw --csv > gpurun_out/r02_self_split_v2b_raw.csv 2>/dev/null
