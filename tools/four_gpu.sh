#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_4gpu_C3.json 2> gpurun_out/bench_4gpu_C3.err
tail -2 gpurun_out/bench_4gpu_C3.err | cut -c1-300
python -c "
import json; j=json.load(open('gpurun_out/bench_4gpu_C3.json')); print('C3 x4', j['value'], j['unit'], 'e2e', j['e2e']['value'], j['n_gpus'], j['clocks'], j.get('argmax_check_fp64_full_pool'))"
