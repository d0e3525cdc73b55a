#!/bin/bash
# A/B of build flags in ONE box: tools/ab_test.sh "<flags A>" "<flags B>" [microbench --only list]
only=${3:-fwd1,fwd3,bwd1,bwd3,gA_fwd,gA_fb,gB_fb,gC_fb}
for round in 1 2; do
for cfg in "$1" "$2"; do
  export SPNB_NVCC_EXTRA="$cfg"
  python -m smoothparticlenets_b200.build > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  echo "== [$round] flags: '$cfg'"
  python tools/microbench.py --graph --iters 10 --only $only 2>&1 | grep -E "^(g|f|b|c|s|r|w)[A-Za-z0-9_]* " | awk '{printf "%s %s ms; ", $1, $2} END {print ""}'
done
done
