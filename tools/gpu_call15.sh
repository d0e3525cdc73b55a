#!/bin/bash
# sort variants: bit-exactness, c2 collision chain, c5 step
timeout 600 python -m pytest tests/test_gpu_hashgrid.py -q -x 2>&1 | tail -2
c2() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'step %.3f coll %.1f us' % (d['ms_per_step'], d['kernels']['particle_collision']['ms']*1000))" $1; }
c5() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1]); print(sys.argv[1], 'c5 %.2f ms' % d['ms_per_step'])" $1; }
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t1_main.json 2>/dev/null; c2 gpurun_out/t1_main.json
timeout 600 python bench.py --workload c5 --gpus 1 --steps 5 --warmup 3 > gpurun_out/t1_c5_main.json 2>/dev/null; c5 gpurun_out/t1_c5_main.json
for v in "$@"; do
export SPNB_NO_BUILD=1 SPNB_LIB=$PWD/smoothparticlenets_b200/_variants/libspnb_$v.so
timeout 600 python -m pytest tests/test_gpu_hashgrid.py -q -x 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/t1_$v.json 2>/dev/null; c2 gpurun_out/t1_$v.json
timeout 600 python bench.py --workload c5 --gpus 1 --steps 5 --warmup 3 > gpurun_out/t1_c5_$v.json 2>/dev/null; c5 gpurun_out/t1_c5_$v.json
done
unset SPNB_NO_BUILD SPNB_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_onesweep -c 12 --csv --log-file gpurun_out/t1_sort.csv python bench.py --workload c5 --gpus 1 --steps 1 --warmup 1 > /dev/null 2>&1
grep -c k_onesweep gpurun_out/t1_sort.csv; grep k_onesweep gpurun_out/t1_sort.csv | tail -3 | cut -d, -f 13-
