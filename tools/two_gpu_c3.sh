#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu_C3.json 2> gpurun_out/bench_2gpu_C3.err
python -c "
import json; j=json.load(open('gpurun_out/bench_2gpu_C3.json')); print('C3 x2', j['value'], 'e2e', j['e2e']['value'], j['clocks'], j['roofline']['frac'])"
