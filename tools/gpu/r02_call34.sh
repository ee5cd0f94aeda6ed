#!/bin/bash
# R = 9...16 through the two-stage combine kernel: parity (split_path incl. the new timeline lengths, generic combine, layout),
# launch times at R = 14 (NF = 26000) and 15 (NF = 30000), smoke
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q -x -k "split or self" ) > gpurun_out/pytest_self10.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_self10.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke_g.log 2>&1
for nf in 26000 30000; do
SASSENA_SELF_PATH=split timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:self_split -c 2 python tools/probe_self.py $nf 64 2>&1 | grep -E "self_split|gpu__time"
done > gpurun_out/reg_launches2.log 2>&1
timeout 200 python tools/probe_self_ab.py 26000 256 v2 > gpurun_out/self_ab10.log 2>&1
