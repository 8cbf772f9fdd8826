#!/bin/bash
# GPU batch S (8 GPUs): the default bench of the final build exactly as the driver's scaling run launches it
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 8 ) > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err
tail -5 gpurun_out/r2s_bench_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_n8.json')); print(d['value'], json.dumps(d['spmv']), json.dumps(d['solve']), json.dumps(d['parity']), json.dumps(d['e2e'])[:700]); print(json.dumps(d.get('c4'))[:1800]); print(json.dumps(d.get('c5'))[:1200])"
