#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py C1 C2 C4 C5 > gpurun_out/other_configs.jsonl 2> gpurun_out/other_configs.err
# launch list + full capture of the split self path on a small slice
SELF_NA=96 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_self.csv python tools/probe_paths.py self > gpurun_out/ncu_self_l.log 2>&1
SELF_NA=96 timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_split -c 6 -o gpurun_out/self_split python tools/probe_paths.py self > gpurun_out/ncu_self_f.log 2>&1
ncu -i gpurun_out/self_split.ncu-rep --page raw --csv > gpurun_out/self_split_raw.csv 2>/dev/null
echo done
