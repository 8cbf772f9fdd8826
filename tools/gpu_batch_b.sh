#!/bin/bash
# GPU batch B (round 2): re-validate the restored tree: GPU tests, smoke, full default bench, launch list
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --timeout 1500 -p no:cacheprovider ) > gpurun_out/r2b_pytest.log 2>&1; tail -4 gpurun_out/r2b_pytest.log
python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; tail -1 gpurun_out/r2b_smoke.log
( time python bench.py ) > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -4 gpurun_out/r2b_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(d['value'], d['roofline']['frac'], json.dumps(d['spmv']), json.dumps(d['solve'])[:1500], json.dumps(d['e2e']), d['clocks'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-solve --no-cpu --no-tts --extras none > gpurun_out/r2b_ncu_bench.log 2>&1
tail -2 gpurun_out/r2b_ncu_bench.log
