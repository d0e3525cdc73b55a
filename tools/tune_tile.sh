#!/bin/bash
# Sweep the tile-list kernels' build parameters on the GPU box.
# usage: tools/tune_tile.sh "1:2:1024 2:2:1024 2:4:1024"   (FWD_G:BWD_G:TILE_CAP)
for cfg in $1; do
  IFS=: read fg bg cap <<< "$cfg"
  export SPNB_NVCC_EXTRA="-DSPNB_TILE_FWD_G=$fg -DSPNB_TILE_BWD_G=$bg -DSPNB_TILE_CAP=$cap"
  python -m smoothparticlenets_b200.build > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo "== FWD_G=$fg BWD_G=$bg CAP=$cap"
  python tools/microbench.py --graph --iters 10 --only kA_f,kA_b,kB_f,kB_b,kC_f,kC_b,kV_f,kV_b 2>&1 | grep -E "^k" | awk '{printf "%s %s ms; ", $1, $2} END {print ""}'
done
