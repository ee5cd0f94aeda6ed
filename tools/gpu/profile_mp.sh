#!/bin/bash
# ncu full capture of the batched multipole kernel (Q = 8 pass; the probe launches six Q = 1 passes first).
# Usage: gpurun --timeout 1500 -- 'bash tools/gpu/profile_mp.sh'
mkdir -p gpurun_out
MP_NF=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:multipole_gemm --launch-skip 6 -c 1 -o gpurun_out/mp_gemm_q8 python tools/probe_paths.py mpbatch > gpurun_out/ncu_mp_q8.log 2>&1
ncu -i gpurun_out/mp_gemm_q8.ncu-rep --page raw --csv > gpurun_out/mp_gemm_q8_raw.csv 2>/dev/null
