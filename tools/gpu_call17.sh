#!/bin/bash
run() { ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_onesweep -c 12 --csv --log-file gpurun_out/y_$1.csv python bench.py --workload c5 --gpus 1 --steps 1 --warmup 1 > /dev/null 2>&1; echo $1 c5; grep k_onesweep gpurun_out/y_$1.csv | tail -4 | awk -F'","' '{print $(NF)}' | tr '\n' ' '; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_onesweep -c 12 --csv --log-file gpurun_out/y2_$1.csv python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1; echo $1 c2; grep k_onesweep gpurun_out/y2_$1.csv | tail -4 | awk -F'","' '{print $(NF)}' | tr '\n' ' '; echo; }
run main
for v in "$@"; do export SPNB_NO_BUILD=1 SPNB_LIB=$PWD/smoothparticlenets_b200/_variants/libspnb_$v.so; run $v; done
