#!/bin/bash
mkdir -p gpurun_out
MP_NF=16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_mp.csv python tools/probe_paths.py mpbatch > /dev/null 2>&1
MP_NF=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:multipole_gemm -c 1 -o gpurun_out/mp_gemm_v2 python tools/probe_paths.py mpbatch > gpurun_out/ncu_mp.log 2>&1
echo done
