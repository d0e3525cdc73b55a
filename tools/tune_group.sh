#!/bin/bash
# Sweep lanes-per-query (G) and entries-in-flight (U) of the fused group kernels on the GPU box.
# usage: tools/tune_group.sh "1:4:1:2 2:4:2:2 4:4:4:2 8:4:8:2"   (FWD_G:FWD_U:BWD_G:BWD_U)
for cfg in $1; do
  IFS=: read fg fu bg bu <<< "$cfg"
  export SPNB_NVCC_EXTRA="-DSPNB_GROUP_FWD_G=$fg -DSPNB_GROUP_FWD_U=$fu -DSPNB_GROUP_BWD_G=$bg -DSPNB_GROUP_BWD_U=$bu"
  python -m smoothparticlenets_b200.build > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo "== FWD G=$fg U=$fu  BWD G=$bg U=$bu"
  python tools/microbench.py --graph --iters 5 --only gA_fwd,gA_fb,gB_fwd,gB_fb,gC_fwd,gC_fb 2>&1 | grep -E "^g"
done
