#!/bin/bash
# final evidence of the round: tests, smoke, bench lines, launch list, ncu of the dominant tile kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/z_tests.txt; cat gpurun_out/z_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; tail -c 600 gpurun_out/z_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; cat gpurun_out/z_bench_ref.json | cut -c1-600
timeout 600 python bench.py --workload c3 --steps 6 --warmup 3 > gpurun_out/z_c3.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 500 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
N="ncu --set full --import-source on --clock-control none"
$N -k regex:k_tile_bwd -s 9 -c 1 -o gpurun_out/r2_tile_bwdA_final python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
$N -k regex:k_tile_fwd -s 10 -c 1 -o gpurun_out/r2_tile_fwdB_final python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/z_c3_launches.csv python bench.py --workload c3 --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/z_* gpurun_out/r2_tile_*final*
