#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_convsp.py -m gpu -q -x -k "wide" 2>&1 | tail -15 > gpurun_out/c8_wide.txt
tail -6 gpurun_out/c8_wide.txt
timeout 300 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -x -s -k "c3" 2>&1 | tail -30 > gpurun_out/c8_c3.txt
tail -6 gpurun_out/c8_c3.txt
timeout 600 python bench.py --workload c3 --steps 5 --warmup 3 > gpurun_out/c8_c3.json 2> gpurun_out/c8_c3.err; tail -c 1300 gpurun_out/c8_c3.json; tail -3 gpurun_out/c8_c3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c8_c3_launches.csv python bench.py --workload c3 --steps 1 --warmup 1 --queries 37888 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c8_c3_launches.csv 2>/dev/null | head -12
