#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -300 > gpurun_out/c4_tests.txt
grep -E "passed|failed|^FAILED|^\[" gpurun_out/c4_tests.txt
summ() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "step %.3f e2e %.3f per_layer %.3f" % (d['ms_per_step'], d['e2e']['ms_per_step'], d['per_layer']['ms_per_step']))
print({k:v['ms'] for k,v in d['kernels'].items()})
PY
}
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; summ gpurun_out/c4_bench.json
for v in ; do
  SPNB_NO_BUILD=1 SPNB_LIB=$PWD/smoothparticlenets_b200/_variants/libspnb_$v.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench_$v.json 2> gpurun_out/c4_bench_$v.err; summ gpurun_out/c4_bench_$v.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/c4_launches.csv python bench.py --steps 1 --warmup 1 --lite > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/c4_launches.csv 2>/dev/null | head -16
