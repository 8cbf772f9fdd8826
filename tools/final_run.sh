#!/bin/bash
# Round-end validation on one B200: GPU tests, smoke, the default bench (both arms), the ncu launch list of the
# bench command (kernel shares of the step) and one full capture of the dominant kernel (DRAM traffic per launch).
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -q --timeout 1500 -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1; tail -2 gpurun_out/final_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
( time python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -4 gpurun_out/final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['roofline']['frac'], json.dumps(d['spmv']), json.dumps(d['solve']), json.dumps(d['tts'])[:1500], json.dumps(d['e2e'])[:600], json.dumps(d['cpu_baseline'])[:600], d['clocks']); print(json.dumps(d['c4'])[:1500]); print(json.dumps(d['c5'])[:1000])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 400 gpurun_out/final_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu --no-tts --extras none > gpurun_out/final_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_small -c 1 --launch-skip 2 -o gpurun_out/final_assemble_small -f python bench.py --steps 2 --warmup 3 --no-solve --no-cpu --no-tts --extras none > gpurun_out/final_ncu_asm.log 2>&1
ncu -i gpurun_out/final_assemble_small.ncu-rep --page raw --csv > gpurun_out/final_assemble_small_raw.csv 2>/dev/null
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/final_launches.csv")) if len(r)>6 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split("(")[0][:64]; t=float(r[-1])
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]:
    print("%-66s n=%4d total %9.3f ms avg %9.1f us %5.1f%%" % (k, v[0], v[1]/1e6, v[1]/v[0]/1e3, 100*v[1]/tot))
rows=list(csv.reader(open("gpurun_out/final_assemble_small_raw.csv")))
d=dict(zip(rows[0], rows[2]))
for k in ("gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed","sm__warps_active.avg.pct_of_peak_sustained_active"):
    print(k, d.get(k))
PY
