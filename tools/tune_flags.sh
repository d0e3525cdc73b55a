#!/bin/bash
# Generic build-flag sweep: tools/tune_flags.sh "<flags A>" "<flags B>" ...  (times the fused group ops)
for cfg in "$@"; do
  export SPNB_NVCC_EXTRA="$cfg"
  python -m smoothparticlenets_b200.build > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  echo "== $cfg"
  python tools/microbench.py --graph --iters 5 --only gA_fwd,gA_fb,gB_fb,gC_fb 2>&1 | grep -E "^g"
done
