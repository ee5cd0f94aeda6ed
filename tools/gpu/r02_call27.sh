#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -k "bench" ) > gpurun_out/pytest_bench.log 2>&1
( timeout 900 python bench.py --workload C3 ) > gpurun_out/bench_c3_parity.json 2> gpurun_out/bench_c3_parity.err
