#!/bin/bash
# A/B of prebuilt variant libraries on one box: tools/gpu_call9.sh tag v1 v2 ...
tag=$1; shift
summ() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "step %.3f e2e %.3f per_layer %.3f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d.get('per_layer',{}).get('ms_per_step',0)))
print({k:round(v['ms']*1000,1) for k,v in d.get('kernels',{}).items()})
PY
}
for round in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_main_$round.json 2> gpurun_out/${tag}_main.err; summ gpurun_out/${tag}_main_$round.json
for v in "$@"; do
  SPNB_NO_BUILD=1 SPNB_LIB=$PWD/smoothparticlenets_b200/_variants/libspnb_$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${v}_$round.json 2> gpurun_out/${tag}_$v.err; summ gpurun_out/${tag}_${v}_$round.json
done
done
