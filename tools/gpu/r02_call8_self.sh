#!/bin/bash
# round 2, eighth GPU call: kernel A v2 with warp-local prefetch (two barriers), interleaved stores: A/B timing, parity, ncu, bench C2
mkdir -p gpurun_out
{
timeout 300 python tools/probe_self_ab.py 10000 2048 fused v1 v2
timeout 300 python tools/probe_self_ab.py 50000 256 v1 v2
timeout 300 python tools/probe_self_ab.py 4500 1024 fused v2
timeout 300 python tools/probe_self_ab.py 9000 512 fused v2
} > gpurun_out/self_ab4.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x -k "self or stage or stream" ) > gpurun_out/pytest_self4.log 2>&1
( timeout 600 compute-sanitizer --tool racecheck python tools/probe_self_ab.py 10000 2 v2 ) > gpurun_out/self_v2_racecheck.log 2>&1
SASSENA_SELF_PATH=split timeout 900 ncu --set full --clock-control none --import-source on -k regex:self_split_fft -c 1 -o gpurun_out/r02_self_split_v2d python tools/probe_self.py 10000 96 > gpurun_out/ncu_self_v2d.log 2>&1
ncu -i gpurun_out/r02_self_split_v2d.ncu-rep --page raw --csv > gpurun_out/r02_self_split_v2d_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_self_split_v2d.ncu-rep --page source --csv > gpurun_out/r02_self_split_v2d_source.csv 2>/dev/null
( timeout 600 python bench.py --workload C2 ) > gpurun_out/bench_c2_v2d.json 2> gpurun_out/bench_c2_v2d.err
