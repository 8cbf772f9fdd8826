#!/bin/bash
mkdir -p gpurun_out
for ov in 1 0; do
PG_HALO_OVERLAP=$ov timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2958$ov bench.py --gpus 8 --extras none --no-cpu --no-tts --jacobi-seconds 2 --solve-maxit 500 > gpurun_out/r2x_bench_n8_ov$ov.json 2> gpurun_out/r2x_bench_n8_ov$ov.err
python -c "
import json; d=json.load(open('gpurun_out/r2x_bench_n8_ov$ov.json')); s=d['solve']; print('overlap=$ov spmv', d['spmv']['ms'], 'spmm4', d['spmv']['four_rhs']['ms'], {k: round(v['ms_per_iteration'],4) for k,v in s.items()}, d['parity']['pass'])"
done
