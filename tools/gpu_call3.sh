#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_tile_bwd -s 3 -c 1 -f -o gpurun_out/r2_bwdA python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_tile_fwd -s 1 -c 1 -f -o gpurun_out/r2_fwdB python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
ls -la gpurun_out/r2*.ncu-rep
