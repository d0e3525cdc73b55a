#!/bin/bash
# Where does the list-walk time go?  Builds the group forward kernel with (a) no gathers/math,
# (b) gathers but no math, (c) everything, and times group A/C forward with CUDA-graph replay.
for mode in "-DSPNB_DEBUG_WALK_ONLY=1" "-DSPNB_DEBUG_NO_MATH=1" ""; do
  export SPNB_NVCC_EXTRA="$mode $1"
  python -m smoothparticlenets_b200.build > /dev/null 2>&1 || { echo "build failed"; continue; }
  echo "== mode: ${mode:-full} $1"
  python tools/microbench.py --graph --iters 5 --only gA_fwd,gC_fwd 2>&1 | grep -E "^g"
done
