#!/bin/bash
# first GPU call of round 2: new parity tests on the round-1 kernels (+ADVICE fixes), f32x2 micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1

timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -150 > gpurun_out/c1_tests.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/c1_smoke.txt 2>&1
tail -5 gpurun_out/c1_tests.txt; cat gpurun_out/c1_f32x2.txt; tail -2 gpurun_out/c1_smoke.txt
