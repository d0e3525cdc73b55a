#!/bin/bash
p() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'fwd %.3f ms  bwd %.3f ms' % (d['ms_per_step'], d['backward']['ms_per_step']))" $1; }
for r in 1 2; do
timeout 300 python bench.py --workload c3 --steps 6 --warmup 3 > gpurun_out/j1_main.json 2>/dev/null; p gpurun_out/j1_main.json
SPNB_NO_BUILD=1 SPNB_LIB=$PWD/smoothparticlenets_b200/_variants/libspnb_$1.so timeout 300 python bench.py --workload c3 --steps 6 --warmup 3 > gpurun_out/j1_$1.json 2>/dev/null; p gpurun_out/j1_$1.json
done
SPNB_NO_BUILD=1 SPNB_LIB=$PWD/smoothparticlenets_b200/_variants/libspnb_$1.so timeout 300 python -m pytest tests/test_gpu_convsp.py -x -q -k "wide" 2>&1 | tail -3
timeout 600 python bench.py --impl reference > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; cut -c1-300 gpurun_out/z_bench_ref.json
