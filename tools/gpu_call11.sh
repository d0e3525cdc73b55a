#!/bin/bash
# c3 A/B: tensor-core gather vs the CUDA-core gather (env knob), then the parity tests
p() { python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'fwd %.3f ms  bwd %.3f ms' % (d['ms_per_step'], d['backward']['ms_per_step']))" $1; }
timeout 300 python bench.py --workload c3 --steps 6 --warmup 3 > gpurun_out/h4_tc.json 2>/dev/null; p gpurun_out/h4_tc.json
SPNB_WIDE_NO_TC=1 timeout 300 python bench.py --workload c3 --steps 6 --warmup 3 > gpurun_out/h4_notc.json 2>/dev/null; p gpurun_out/h4_notc.json
timeout 300 python -m pytest tests/test_gpu_convsp.py tests/test_gpu_parity_configs.py -x -q -s -k "wide or c3" 2>&1 | tail -6
SPNB_WIDE_NO_TC=1 timeout 300 python -m pytest tests/test_gpu_convsp.py tests/test_gpu_parity_configs.py -x -q -s -k "wide or c3" 2>&1 | tail -6
