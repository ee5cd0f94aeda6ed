#!/bin/bash
# A/B on ONE box: kernel A with the old (loop) and the new (lanes 0..5) copy of the floats around a sub-sequence's ends
mkdir -p gpurun_out
for v in old new old new; do
  if [ $v = old ]; then export SASSENA_B200_LIB=$PWD/tools/ab/libsassena_b200_old.so; else unset SASSENA_B200_LIB; fi
  echo "== $v"; timeout 14 python tools/probe_self_ab.py 50000 256 v2 2>&1 | grep "^v2:"
done > gpurun_out/self_ab13.log 2>&1
