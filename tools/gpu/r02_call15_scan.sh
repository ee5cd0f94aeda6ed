#!/bin/bash
# corrected scan kernel with the first-order sums in FP32 (CORR = 2, 25-|q| passes): accuracy against the per-|q| kernel for
# every |q|, speed against the FP64-D kernel, parity tests of the scan paths, ncu capture
mkdir -p gpurun_out
{
echo "== FP32-D (default)"; ROUNDED=1 NQ=50 NF=300 timeout 600 python tools/scan_probe.py
echo "== FP64-D (SASSENA_SCAN_FP64_CORR=1)"; SASSENA_SCAN_FP64_CORR=1 ROUNDED=1 NQ=50 NF=300 timeout 600 python tools/scan_probe.py
echo "== plain"; NQ=50 NF=300 timeout 600 python tools/scan_probe.py
} > gpurun_out/scan_probe_fp32d.log 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -x -k "scan or all_vectors or bench" ) > gpurun_out/pytest_scan.log 2>&1
ROUNDED=1 NQ=50 timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_sym -c 2 -o gpurun_out/r02_scan_corr_fp32d python tools/scan_probe.py > gpurun_out/ncu_scan_corr_fp32d.log 2>&1
ncu -i gpurun_out/r02_scan_corr_fp32d.ncu-rep --page raw --csv > gpurun_out/r02_scan_corr_fp32d_raw.csv 2>/dev/null
