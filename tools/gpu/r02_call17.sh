#!/bin/bash
# multipole: per-chunk spherical conversion on its own stream (e2e overlap); parity tests + bench C4
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "mp or multipole or stage or job or cli or host_layer" ) > gpurun_out/pytest_mp2.log 2>&1
( timeout 600 python bench.py --workload C4 ) > gpurun_out/bench_c4_conv.json 2> gpurun_out/bench_c4_conv.err
