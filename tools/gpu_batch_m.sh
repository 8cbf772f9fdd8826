#!/bin/bash
# GPU batch M: GPU tests (fused dots, peer path), then the C3 bench with and without the fused dots
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_path.py tests/test_peer_gpu.py tests/test_glue.py -m gpu -q -x --timeout 1500 -p no:cacheprovider ) > gpurun_out/r2m_pytest.log 2>&1; tail -5 gpurun_out/r2m_pytest.log
for f in 1 0; do
PG_FUSED_DOTS=$f python bench.py --extras none --no-cpu --no-tts --jacobi-seconds 6 > gpurun_out/r2m_bench_fused$f.json 2> gpurun_out/r2m_bench_fused$f.err; tail -2 gpurun_out/r2m_bench_fused$f.err
python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_fused$f.json')); print('fused=$f', json.dumps(d['spmv'])[:200]); print(json.dumps(d['solve']))"
done
