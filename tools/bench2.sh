#!/bin/bash
# 2-GPU bench run (the driver's launch line at N=2)
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo rc=$?; tail -5 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['value'], d['e2e'], d.get('strong'), d.get('parity'))"
