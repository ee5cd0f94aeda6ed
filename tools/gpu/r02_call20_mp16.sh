#!/bin/bash
# multipole kernel with 16 |q| per pass: parity tests, bench C4 with 16 and (A/B) 8 |q| per pass
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "mp or multipole or host_layer or job or cli or bench" ) > gpurun_out/pytest_mp3.log 2>&1
( timeout 600 python bench.py --workload C4 ) > gpurun_out/bench_c4_q16.json 2> gpurun_out/bench_c4_q16.err
( SASSENA_BENCH_MP_BATCH=8 timeout 600 python bench.py --workload C4 --no-cpu ) > gpurun_out/bench_c4_q8.json 2> gpurun_out/bench_c4_q8.err
