#!/bin/bash
# ncu launch list + full capture of the split self kernels on a config-2 shaped slice (96 atoms x 10000 frames x 200 q).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu/profile_self.sh'; read gpurun_out/self_split_raw.csv here
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_self.csv python tools/probe_self.py 10000 96 > gpurun_out/ncu_self_l.log 2>&1
SASSENA_SELF_PATH=split timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_split -c 4 -o gpurun_out/self_split python tools/probe_self.py 10000 96 > gpurun_out/ncu_self_f.log 2>&1
ncu -i gpurun_out/self_split.ncu-rep --page raw --csv > gpurun_out/self_split_raw.csv 2>/dev/null
