#!/bin/bash
# round 2, validation after the two-stage combine rework: all GPU tests, smoke, default bench; launch times of the register
# combine kernel at R = 13 / 15 (no BASELINE config uses them; to know where they stand)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --durations=5 ) > gpurun_out/pytest_gpu_f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_f.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke_f.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/bench_all_n1_f.json 2> gpurun_out/bench_all_n1_f.err
for nf in 26000 30000; do
SASSENA_SELF_PATH=split timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:self_split -c 4 python tools/probe_self.py $nf 64 2>&1 | grep -E "self_split|gpu__time"
done > gpurun_out/reg_launches.log 2>&1
