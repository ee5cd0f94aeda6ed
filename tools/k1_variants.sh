#!/bin/bash
# times the K1 variants (SASSENA_K1_VARIANT) on a cfg3-shaped slice; usage: tools/k1_variants.sh [frames]
F=${1:-1500}
for v in ${VARIANTS:-0 1 2 3 4 5 6 7}; do
  echo "variant $v: $(SASSENA_K1_VARIANT=$v python tools/first_light.py $F 2>&1 | tail -1)"
done
