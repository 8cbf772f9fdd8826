#!/bin/bash
# GPU batch C (round 2, 2 GPUs): 2-rank CLI parity test, C3 bench with the peer transport and with NCCL
mkdir -p gpurun_out
( time python -m pytest tests/test_glue.py -m gpu -q -x -k two_ranks --timeout 900 -p no:cacheprovider ) > gpurun_out/r2c_pytest.log 2>&1; tail -4 gpurun_out/r2c_pytest.log
for tr in peer nccl; do
  PG_TRANSPORT=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --extras none --no-cpu --no-tts > gpurun_out/r2c_bench_n2_$tr.json 2> gpurun_out/r2c_bench_n2_$tr.err
  tail -3 gpurun_out/r2c_bench_n2_$tr.err
  python -c "
import json; d=json.load(open('gpurun_out/r2c_bench_n2_$tr.json')); print('$tr', d['value'], json.dumps(d['spmv']), json.dumps(d['solve']), json.dumps(d['parity']))"
done
