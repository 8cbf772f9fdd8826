#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_path.py -m gpu -q -x -k "multi_rhs or spmv or lockstep or krylov or petsc_fixture" --timeout 900 -p no:cacheprovider 2>&1 | tail -3
python bench.py --m 40 --extras none --no-cpu --solve-maxit 50 --jacobi-seconds 1 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; tail -2 gpurun_out/r2r_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); t=d['tts']; print({k:(round(v['seconds'],4), v['iterations']) for k,v in t.items() if isinstance(v,dict) and 'seconds' in v})"
