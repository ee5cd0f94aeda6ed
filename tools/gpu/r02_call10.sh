#!/bin/bash
# experiment: skew between the two co-resident CTAs of kernel A (lock-step hypothesis)
mkdir -p gpurun_out
{
for skew in 0 2000 4000 6000 8000; do
echo "skew $skew"
SASSENA_SELF_PATH=split SASSENA_SELF_SKEW=$skew timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:self_split_fft -c 3 python tools/probe_self.py 10000 96 2>&1 | grep -E "gpu__time|rel.err"
done
} > gpurun_out/self_skew.log 2>&1
