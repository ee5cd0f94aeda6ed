#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu2.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --workload C2 > gpurun_out/bench_c2_n1_v2.json 2> gpurun_out/bench_c2_n1_v2.err
timeout 600 python bench.py > gpurun_out/bench_n1_v2.json 2> gpurun_out/bench_n1_v2.err
timeout 600 python tools/bench_configs.py C5 > gpurun_out/c5_v2.jsonl 2> gpurun_out/c5_v2.err
echo done
