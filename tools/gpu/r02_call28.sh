#!/bin/bash
# final captures of the split self kernels (after the last change of kernel A) and the launch list of the config-2 workload
mkdir -p gpurun_out
SASSENA_SELF_PATH=split timeout 600 ncu --set full --clock-control none --import-source on -k regex:self_split -c 2 -o gpurun_out/r02_self_split_final python tools/probe_self.py 10000 96 > gpurun_out/ncu_self_final.log 2>&1
ncu -i gpurun_out/r02_self_split_final.ncu-rep --page raw --csv > gpurun_out/r02_self_split_final_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --workload C2 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/r02_launches_c2.log 2>&1
