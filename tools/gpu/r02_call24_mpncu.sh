#!/bin/bash
mkdir -p gpurun_out
MP_NF=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:multipole_gemm --launch-skip 6 -c 1 -o gpurun_out/r02_mp_gemm_v3 python tools/probe_paths.py mpbatch > gpurun_out/ncu_mp_v3.log 2>&1
ncu -i gpurun_out/r02_mp_gemm_v3.ncu-rep --page raw --csv > gpurun_out/r02_mp_gemm_v3_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_mp_gemm_v3.ncu-rep --page source --csv > gpurun_out/r02_mp_gemm_v3_source.csv 2>/dev/null
