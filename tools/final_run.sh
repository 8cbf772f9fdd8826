# Round-end validation on one B200: GPU tests, smoke, the default bench (both arms), launch list, ncu captures.
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1; tail -2 gpurun_out/final_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches_m40.csv python bench.py --steps 3 --warmup 3 --m 40 --solve-maxit 60 --no-tts --no-cpu > gpurun_out/final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"assemble_small" -s 3 -c 1 -o gpurun_out/prof_final_asm_m40 python bench.py --steps 3 --warmup 3 --m 40 --no-cpu --no-solve --no-tts > gpurun_out/final_ncu_asm.log 2>&1
for k in assemble_small spmv_kernel spmm_blocked2; do
  ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -o gpurun_out/prof_final_c3_$k python bench.py --steps 3 --warmup 3 --no-cpu --no-solve --no-tts > gpurun_out/final_ncu_c3_$k.log 2>&1
done
ls -la gpurun_out/prof_final_*
