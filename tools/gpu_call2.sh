#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -300 > gpurun_out/c2_tests.txt
grep -E "passed|failed|^FAILED|^\[" gpurun_out/c2_tests.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c2_bench.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['per_layer']['ms_per_step'])
print({k:v['ms'] for k,v in d['kernels'].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/c2_launches.csv python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c2_launches.csv 2>/dev/null | head -40
