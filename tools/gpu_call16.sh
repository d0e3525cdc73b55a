#!/bin/bash
for v in "$@"; do
export SPNB_NO_BUILD=1 SPNB_LIB=$PWD/smoothparticlenets_b200/_variants/libspnb_$v.so
timeout 600 python -m pytest tests/test_gpu_hashgrid.py -q -x 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_onesweep -c 12 --csv --log-file gpurun_out/t3_sort_$v.csv python bench.py --workload c5 --gpus 1 --steps 1 --warmup 1 > /dev/null 2>&1
echo $v; grep k_onesweep gpurun_out/t3_sort_$v.csv | tail -3 | awk -F'","' '{print $(NF)}'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_onesweep -c 12 --csv --log-file gpurun_out/t3_sort_c2_$v.csv python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
grep k_onesweep gpurun_out/t3_sort_c2_$v.csv | tail -3 | awk -F'","' '{print $(NF)}'
done
