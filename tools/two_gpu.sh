#!/bin/bash
P=29511
for c in C3 C1; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --config $c --gpus 2 --steps $([ $c = C1 ] && echo 5 || echo 3) --warmup 3 > gpurun_out/bench_2gpu_$c.json 2> gpurun_out/bench_2gpu_$c.err
  tail -2 gpurun_out/bench_2gpu_$c.err | cut -c1-300
  python -c "
import json; j=json.load(open('gpurun_out/bench_2gpu_$c.json')); print('$c', j['value'], j['unit'], 'e2e', j['e2e']['value'], j['n_gpus'], j['clocks'])"
  P=$((P+1))
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_2gpu_reference.json 2> gpurun_out/bench_2gpu_reference.err; cut -c1-200 gpurun_out/bench_2gpu_reference.json
