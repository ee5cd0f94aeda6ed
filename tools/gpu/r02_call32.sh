#!/bin/bash
# two-stage combine kernel after the 64-thread / cp.async rework: full capture, racecheck + memcheck at R = 25
mkdir -p gpurun_out
SASSENA_SELF_PATH=split timeout 300 ncu --set full --clock-control none --import-source on -k regex:combine_2s -c 1 -o gpurun_out/r02_combine_2s python tools/probe_self.py 50000 64 > gpurun_out/ncu_2s.log 2>&1
ncu -i gpurun_out/r02_combine_2s.ncu-rep --page raw --csv > gpurun_out/r02_combine_2s_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_combine_2s.ncu-rep --page source --csv > gpurun_out/r02_combine_2s_source.csv 2>/dev/null
rm -f gpurun_out/r02_combine_2s.ncu-rep
{
echo "--- racecheck: python tools/probe_self_ab.py 50000 2 v2"
timeout 300 compute-sanitizer --tool racecheck python tools/probe_self_ab.py 50000 2 v2 2>&1 | grep -E "=========|v2" | tail -8
echo "--- memcheck: python tools/probe_self_ab.py 50001 3 v2"
timeout 300 compute-sanitizer --tool memcheck python tools/probe_self_ab.py 50001 3 v2 2>&1 | grep -E "=========|v2" | tail -8
} > gpurun_out/sanitizer_2s.log 2>&1
