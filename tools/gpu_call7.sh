#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --workload c5 --gpus $N --check > gpurun_out/c7_c5_check$N.json 2> gpurun_out/c7_c5_check$N.err; tail -c 900 gpurun_out/c7_c5_check$N.json; tail -3 gpurun_out/c7_c5_check$N.err
timeout 900 $TR bench.py --workload c5 --gpus $N --steps 3 --warmup 3 > gpurun_out/c7_c5_n$N.json 2> gpurun_out/c7_c5_n$N.err; tail -c 1900 gpurun_out/c7_c5_n$N.json; tail -3 gpurun_out/c7_c5_n$N.err
timeout 600 python -m pytest tests/test_gpu_scene_sharding.py -m gpu -q -x 2>&1 | tail -5
