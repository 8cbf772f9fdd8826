#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py -m gpu -q -x -k "fused_dot or krylov or lockstep" --timeout 900 -p no:cacheprovider 2>&1 | tail -3
for f in 1 0 1 0; do
PG_FUSED_DOTS=$f python bench.py --extras none --no-cpu --no-tts --jacobi-seconds 4 --solve-maxit 600 > gpurun_out/r2n_bench_fused$f.json 2> gpurun_out/r2n_bench_fused$f.err; tail -2 gpurun_out/r2n_bench_fused$f.err
python -c "
import json; d=json.load(open('gpurun_out/r2n_bench_fused$f.json')); s=d['solve']; print('fused=$f', {k: round(v['ms_per_iteration'],3) for k,v in s.items()})"
done
