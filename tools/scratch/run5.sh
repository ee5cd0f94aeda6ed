#!/bin/bash
mkdir -p gpurun_out
for nf in 6000 7000 10000 12000 16000 50000; do
  na=2048; [ $nf -ge 16000 ] && na=1024
  timeout 300 python tools/probe_self.py $nf $na > gpurun_out/probe_self_${nf}_v3.log 2>&1
done
echo done
