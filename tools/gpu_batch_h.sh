#!/bin/bash
# GPU batch H: stream reference points, ncu of the x-free SpMV (DRAM bytes of the streams alone)
mkdir -p gpurun_out
timeout 600 python tools/spmv_probe.py --only whole > gpurun_out/r2h_spmv_probe.jsonl 2> gpurun_out/r2h_spmv_probe.err; tail -3 gpurun_out/r2h_spmv_probe.err
head -4 gpurun_out/r2h_spmv_probe.jsonl | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_ltcfabric.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:spmv_blocked2 --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2h_ncu_floor.csv python tools/spmv_probe.py --only whole --quick --quick-floor > gpurun_out/r2h_ncu_floor.log 2>&1
cat gpurun_out/r2h_ncu_floor.csv | tail -14 | cut -d, -f 5,13-16
