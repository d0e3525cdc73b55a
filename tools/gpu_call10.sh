#!/bin/bash
# round-2 ncu evidence: collide, wide gather (final), wide GEMMs, ConvSDF at the c4 shape
N="ncu --set full --import-source on --clock-control none"
$N -k regex:k_collide_tiles -c 1 -o gpurun_out/r2_collide_b python bench.py --steps 1 --warmup 1 --lite > gpurun_out/f1_a.log 2>&1
$N -k regex:k_wide_gather -s 2 -c 1 -o gpurun_out/r2_wide_gather_b python bench.py --workload c3 --steps 1 --warmup 1 --queries 37888 > gpurun_out/f1_b.log 2>&1
$N -k regex:k_wide_gemm -s 2 -c 1 -o gpurun_out/r2_wide_gemm python bench.py --workload c3 --steps 1 --warmup 1 --queries 37888 > gpurun_out/f1_c.log 2>&1
$N -k regex:k_wide_d -s 2 -c 2 -o gpurun_out/r2_wide_dgemms python bench.py --workload c3 --steps 1 --warmup 1 --queries 37888 > gpurun_out/f1_d.log 2>&1
$N -k regex:k_convsdf -s 2 -c 2 -o gpurun_out/r2_convsdf python tools/config_bench.py c4 > gpurun_out/f1_e.log 2>&1
tail -2 gpurun_out/f1_a.log gpurun_out/f1_e.log
ls -la gpurun_out/*.ncu-rep | tail -6
