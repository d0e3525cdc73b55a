#!/bin/bash
# Time the group ops with every prebuilt library variant under smoothparticlenets_b200/_variants/
# (built locally with SPNB_LIB=... SPNB_NVCC_EXTRA=... python -m smoothparticlenets_b200.build).
only=${1:-kA_f,kA_b,kB_f,kB_b,kC_f,kC_b,kV_f,kV_b}
for lib in smoothparticlenets_b200/_variants/*.so; do
  echo "== $(basename $lib)"
  SPNB_LIB=$PWD/$lib python tools/microbench.py --graph --iters 10 --only $only 2>&1 | grep -E "^[a-zA-Z_0-9]+ +[0-9.]+ ms" | awk '{printf "%s %s ms; ", $1, $2} END {print ""}'
done
