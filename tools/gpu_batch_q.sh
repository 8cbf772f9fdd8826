#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py -m gpu -q -x -k "multi_rhs or spmv or lockstep or krylov" --timeout 900 -p no:cacheprovider 2>&1 | tail -3
python bench.py --m 40 --extras c4 --no-cpu --no-tts --no-solve > gpurun_out/r2q_bench_c4.json 2> gpurun_out/r2q_bench_c4.err; tail -2 gpurun_out/r2q_bench_c4.err
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench_c4.json')); print(json.dumps(d['c4'])[:1800])"
