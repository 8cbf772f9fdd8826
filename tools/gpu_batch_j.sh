#!/bin/bash
# GPU batch J: GPU tests of the SpMV paths after the 256-bit change, ncu of the new default kernel
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_path.py tests/test_peer_gpu.py -m gpu -q -x --timeout 1500 -p no:cacheprovider ) > gpurun_out/r2j_pytest.log 2>&1; tail -4 gpurun_out/r2j_pytest.log
for which in block whole; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_blocked2 --launch-skip 3 -c 1 -o gpurun_out/r2j_spmv_$which -f python tools/spmv_probe.py --only $which --quick > gpurun_out/r2j_ncu_$which.log 2>&1
ncu -i gpurun_out/r2j_spmv_$which.ncu-rep --page raw --csv > gpurun_out/r2j_spmv_${which}_raw.csv 2>/dev/null
done
