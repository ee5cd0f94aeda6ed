#!/bin/bash
# round 2, first GPU call: parity of the symmetric scan kernels + their throughput, K3 standalone timing + ncu capture
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
{
for nq in 25 29 17 50; do NQ=$nq timeout 300 python tools/scan_probe.py; done
for nq in 17 15 50; do ROUNDED=1 NQ=$nq timeout 300 python tools/scan_probe.py; done
VARB=1 NQ=24 timeout 300 python tools/scan_probe.py
} > gpurun_out/scan_probe.log 2>&1
{
timeout 300 python tools/probe_k3.py
NM=500 NF=10000 timeout 300 python tools/probe_k3.py
NM=25000 NF=1000 timeout 300 python tools/probe_k3.py
} > gpurun_out/k3_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"corr_|reduce_" -c 12 -o gpurun_out/r02_k3 python tools/probe_k3.py > gpurun_out/ncu_k3.log 2>&1
timeout 600 python tools/probe_self.py > gpurun_out/self_probe.log 2>&1
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/box.txt; free -g >> gpurun_out/box.txt; nproc >> gpurun_out/box.txt
