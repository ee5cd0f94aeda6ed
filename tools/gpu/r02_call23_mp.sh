#!/bin/bash
# multipole gemm kernel: B table with l fastest (one line per warp load), two-atom Legendre tasks, Bessel ladders first
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "mp or multipole or host_layer" ) > gpurun_out/pytest_mp5.log 2>&1
{ MP_NF=16 timeout 300 python tools/probe_paths.py mpbatch 2>&1 | grep "mpsphere"; } > gpurun_out/mp_probe5.log 2>&1
( timeout 600 python bench.py --workload C4 ) > gpurun_out/bench_c4_v3.json 2> gpurun_out/bench_c4_v3.err
