#!/bin/bash
# c5 on a 2-GPU box: parity leg and timing at 1 and 2 GPUs
for n in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --workload c5 --gpus $n --check > gpurun_out/q1_check$n.json 2> gpurun_out/q1_check$n.err; head -c 330 gpurun_out/q1_check$n.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --workload c5 --gpus $n --steps 5 --warmup 3 > gpurun_out/q1_c5_n$n.json 2> gpurun_out/q1_c5_n$n.err; python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().split('\n')[-1]); print(sys.argv[1], d['ms_per_step'], d['config'].get('peak_memory_bytes_per_rank_max'))" gpurun_out/q1_c5_n$n.json
done
tail -3 gpurun_out/q1_check2.err
