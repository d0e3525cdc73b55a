#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_convsp.py -m gpu -q -x -k "wide" 2>&1 | tail -40 > gpurun_out/c5_wide.txt
tail -25 gpurun_out/c5_wide.txt
timeout 300 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -x -s -k "c3" 2>&1 | tail -30 > gpurun_out/c5_c3.txt
tail -12 gpurun_out/c5_c3.txt
nvidia-smi --query-gpu=name,memory.used --format=csv
