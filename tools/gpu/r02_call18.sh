#!/bin/bash
# chunked asynchronous atom staging (self path): parity tests + bench C2 / C5s
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -x -k "self or stage or stream or job or cli or host_layer or bench" ) > gpurun_out/pytest_self6.log 2>&1
( timeout 600 python bench.py --workload C2 ) > gpurun_out/bench_c2_async.json 2> gpurun_out/bench_c2_async.err
