#!/bin/bash
# GPU batch F (round 2, 1 GPU): SpMV probe (fixed floor), ncu of the blocked SpMV (row block and whole), DMMA A/B
mkdir -p gpurun_out
timeout 900 python tools/spmv_probe.py > gpurun_out/r2f_spmv_probe.jsonl 2> gpurun_out/r2f_spmv_probe.err; tail -3 gpurun_out/r2f_spmv_probe.err
cut -c1-330 gpurun_out/r2f_spmv_probe.jsonl
for which in block whole; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_blocked2 --launch-skip 3 -c 1 -o gpurun_out/r2f_spmv_$which -f python tools/spmv_probe.py --only $which --quick > gpurun_out/r2f_ncu_$which.log 2>&1
  ncu -i gpurun_out/r2f_spmv_$which.ncu-rep --page raw --csv > gpurun_out/r2f_spmv_${which}_raw.csv 2>/dev/null
done
timeout 600 python tools/dmma_ab.py > gpurun_out/r2f_dmma_ab.json 2> gpurun_out/r2f_dmma_ab.err; tail -3 gpurun_out/r2f_dmma_ab.err; cut -c1-3000 gpurun_out/r2f_dmma_ab.json
