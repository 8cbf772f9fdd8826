#!/bin/bash
# GPU batch L: launch list of the Krylov section (per-kernel device times of a COCR + Hiptmair iteration at C3)
mkdir -p gpurun_out
PG_CUDA_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 60 -c 1500 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --solve-maxit 40 --jacobi-seconds 1 --no-tts --no-cpu --extras none > gpurun_out/r2l_bench.log 2>&1
tail -2 gpurun_out/r2l_bench.log | cut -c1-300
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2l_launches.csv")) if len(r)>6 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split("(")[0][:60]; t=float(r[-1])
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=t
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:32]:
    print("%-62s n=%5d total %9.3f ms  avg %8.1f us  %4.1f%%" % (k, v[0], v[1]/1e6, v[1]/v[0]/1e3, 100*v[1]/tot))
PY
