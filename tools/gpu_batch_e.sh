#!/bin/bash
# GPU batch E (round 2, 1 GPU): full GPU test suite (new C-ABI tests), SpMV L2 probe
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x --timeout 1500 -p no:cacheprovider ) > gpurun_out/r2e_pytest.log 2>&1; tail -4 gpurun_out/r2e_pytest.log
timeout 900 python tools/spmv_probe.py > gpurun_out/r2e_spmv_probe.jsonl 2> gpurun_out/r2e_spmv_probe.err; tail -3 gpurun_out/r2e_spmv_probe.err
cat gpurun_out/r2e_spmv_probe.jsonl | cut -c1-400
