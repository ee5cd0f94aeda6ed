#!/bin/bash
# multipole: one 512-thread CTA per SM, 8 vs 16 |q| per pass
mkdir -p gpurun_out
{ MP_NQ=8 MP_NF=16 timeout 300 python tools/probe_paths.py mpbatch 2>&1 | grep "batch of";  MP_NQ=16 MP_NF=16 timeout 300 python tools/probe_paths.py mpbatch 2>&1 | grep "batch of"; } > gpurun_out/mp_tile3.log 2>&1
