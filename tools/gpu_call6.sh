#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/c6_c3.json 2> gpurun_out/c6_c3.err; tail -c 1500 gpurun_out/c6_c3.json; tail -3 gpurun_out/c6_c3.err
SPNB_WIDE_NO_MMA=1 timeout 600 python bench.py --workload c3 --steps 3 --warmup 3 --queries 16384 > gpurun_out/c6_c3_nomma.json 2> gpurun_out/c6_c3_nomma.err; python -c "
import json; d=json.load(open('gpurun_out/c6_c3_nomma.json')); print('no-mma (CUDA-core factored kernel):', d['value'], d['ms_per_step'])"
timeout 600 python bench.py --workload c5 --gpus 1 --check > gpurun_out/c6_c5_check1.json 2> gpurun_out/c6_c5_check1.err; tail -c 800 gpurun_out/c6_c5_check1.json; tail -3 gpurun_out/c6_c5_check1.err
timeout 900 python bench.py --workload c5 --gpus 1 --steps 3 --warmup 3 > gpurun_out/c6_c5_n1.json 2> gpurun_out/c6_c5_n1.err; tail -c 1800 gpurun_out/c6_c5_n1.json; tail -3 gpurun_out/c6_c5_n1.err
