#!/bin/bash
# GPU batch O (8 GPUs): default bench as the driver launches it (current build), per-entry-point timeline of a COCR iteration
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 ) > gpurun_out/r2o_bench_n8.json 2> gpurun_out/r2o_bench_n8.err
tail -5 gpurun_out/r2o_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 tools/iteration_timeline.py > gpurun_out/r2o_timeline_n8.json 2> gpurun_out/r2o_timeline_n8.err
tail -3 gpurun_out/r2o_timeline_n8.err
python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_n8.json')); print(d['value'], json.dumps(d['spmv']), json.dumps(d['solve']), json.dumps(d['parity'])); print(json.dumps(d.get('c4'))[:1800]); print(json.dumps(d.get('c5'))[:1200])
t=json.load(open('gpurun_out/r2o_timeline_n8.json'))
for pc,v in t['per_iteration_us'].items():
    print(pc, v['sum_us'], v['wall_us_per_iteration_eager'])
    for k,e in v['entry_points'].items(): print('   %-28s %5.2f calls %9.2f us' % (k, e['calls_per_iteration'], e['us_per_iteration']))
"
