#!/bin/bash
# config-5 shaped slice (NF = 50000, R = 25): launch list and a full capture of the two-stage combine kernel
mkdir -p gpurun_out
SASSENA_SELF_PATH=split timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:self_split -c 12 python tools/probe_self.py 50000 64 2>&1 | grep -E "self_split|gpu__time" > gpurun_out/c5_launches.log
SASSENA_SELF_PATH=split timeout 600 ncu --set full --clock-control none --import-source on -k regex:combine_2s -c 1 -o gpurun_out/r02_combine_2s python tools/probe_self.py 50000 64 > gpurun_out/ncu_2s.log 2>&1
ncu -i gpurun_out/r02_combine_2s.ncu-rep --page raw --csv > gpurun_out/r02_combine_2s_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_combine_2s.ncu-rep --page source --csv > gpurun_out/r02_combine_2s_source.csv 2>/dev/null
