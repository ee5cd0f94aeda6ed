#!/bin/bash
# two-stage combine kernel variants: parity tests, timing at NF = 50000 / 33000, launch times
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "split_path" ) > gpurun_out/pytest_self9.log 2>&1
{
timeout 300 python tools/probe_self_ab.py 50000 256 v2
} > gpurun_out/self_ab9.log 2>&1
SASSENA_SELF_PATH=split timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:combine_2s -c 2 python tools/probe_self.py 50000 64 2>&1 | grep -E "gpu__time" > gpurun_out/c5_launches3.log
