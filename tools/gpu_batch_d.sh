#!/bin/bash
# GPU batch D (round 2, 8 GPUs): the default bench exactly as the driver's scaling run launches it (peer transport),
# then the C3 part again over NCCL for the A/B
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 ) > gpurun_out/r2d_bench_n8_peer.json 2> gpurun_out/r2d_bench_n8_peer.err
tail -5 gpurun_out/r2d_bench_n8_peer.err
PG_TRANSPORT=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --extras none --no-cpu --no-tts --jacobi-seconds 3 > gpurun_out/r2d_bench_n8_nccl.json 2> gpurun_out/r2d_bench_n8_nccl.err
tail -3 gpurun_out/r2d_bench_n8_nccl.err
for tr in peer nccl; do python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_n8_$tr.json')); print('$tr', d['value'], json.dumps(d['spmv']), json.dumps(d['solve']), json.dumps(d['parity']), json.dumps(d['e2e'])[:900]); print(json.dumps(d.get('c4'))[:1200]); print(json.dumps(d.get('c5'))[:800])"; done
