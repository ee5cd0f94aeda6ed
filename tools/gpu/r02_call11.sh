#!/bin/bash
# kernel A v2 with pair-granular distribution: parity tests, A/B timing, ncu, bench C2 + C5s
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "self or stage or stream" ) > gpurun_out/pytest_self5.log 2>&1
{
timeout 300 python tools/probe_self_ab.py 10000 2048 fused v1 v2
timeout 300 python tools/probe_self_ab.py 50000 256 v1 v2
timeout 300 python tools/probe_self_ab.py 4500 1024 fused v2
timeout 300 python tools/probe_self_ab.py 10000 1 fused v2
} > gpurun_out/self_ab5.log 2>&1
SASSENA_SELF_PATH=split timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_split_fft -c 1 -o gpurun_out/r02_self_split_v2e python tools/probe_self.py 10000 96 > gpurun_out/ncu_self_v2e.log 2>&1
ncu -i gpurun_out/r02_self_split_v2e.ncu-rep --page raw --csv > gpurun_out/r02_self_split_v2e_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_self_split_v2e.ncu-rep --page source --csv > gpurun_out/r02_self_split_v2e_source.csv 2>/dev/null
( timeout 600 python bench.py --workload C2 ) > gpurun_out/bench_c2_v2e.json 2> gpurun_out/bench_c2_v2e.err
( timeout 600 python bench.py --workload C5s ) > gpurun_out/bench_c5s_v2e.json 2> gpurun_out/bench_c5s_v2e.err
