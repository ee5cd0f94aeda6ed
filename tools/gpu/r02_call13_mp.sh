#!/bin/bash
# multipole gemm kernel as two 256-thread CTAs per SM: parity tests, bench C4, ncu capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "mpsphere or multipole or mp_" ) > gpurun_out/pytest_mp.log 2>&1
( timeout 600 python bench.py --workload C4 ) > gpurun_out/bench_c4_2cta.json 2> gpurun_out/bench_c4_2cta.err
MP_NF=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:multipole_gemm --launch-skip 6 -c 1 -o gpurun_out/r02_mp_gemm_2cta python tools/probe_paths.py mpbatch > gpurun_out/ncu_mp_2cta.log 2>&1
ncu -i gpurun_out/r02_mp_gemm_2cta.ncu-rep --page raw --csv > gpurun_out/r02_mp_gemm_2cta_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_mp_gemm_2cta.ncu-rep --page source --csv > gpurun_out/r02_mp_gemm_2cta_source.csv 2>/dev/null
