#!/bin/bash
# GPU batch A (round 2): full GPU test suite, default bench (C3 + extras), bulk-store A/B, ncu launch list + full capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_path.py::test_receiver_fields_match_reference_pipeline_high_order[3]" 2>&1 | tail -25 > gpurun_out/r2a_pytest.log
( time python bench.py ) > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
PG_ASM_BULK=0 python bench.py --no-solve --no-cpu --extras none > gpurun_out/r2a_bench_nobulk.json 2> gpurun_out/r2a_bench_nobulk.err
PG_ASM_BULK=1 python bench.py --no-solve --no-cpu --extras none > gpurun_out/r2a_bench_bulk.json 2> gpurun_out/r2a_bench_bulk.err
ncu --set full --clock-control none --import-source on -k regex:assemble_small -c 1 -o gpurun_out/r2a_asm_bulk -f python bench.py --no-solve --no-cpu --extras none --steps 2 --warmup 3 > gpurun_out/r2a_ncu_asm.log 2>&1
ncu -i gpurun_out/r2a_asm_bulk.ncu-rep --page raw --csv > gpurun_out/r2a_asm_bulk_raw.csv 2>/dev/null
tail -5 gpurun_out/r2a_pytest.log
python - <<'PY'
import json
for f in ("r2a_bench","r2a_bench_nobulk","r2a_bench_bulk"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f))
        print(f, "el/s %.3e"%d["value"], "frac %.3f"%d["roofline"]["frac"], "kernel_ms %.3f"%d["roofline"]["kernel_ms"], "spmv", d["spmv"]["ms"])
        if d.get("solve"): print("  solve", json.dumps(d["solve"].get("cocr+hiptmair")))
        for k in ("c4","c5"):
            if k in d: print("  ",k, json.dumps(d[k])[:600])
        if d.get("cpu_baseline"): print("  cpu", json.dumps(d["cpu_baseline"])[:900])
        if d.get("tts"): print("  tts", json.dumps(d["tts"])[:1500])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 gpurun_out/r2a_bench.err
