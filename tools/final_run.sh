# Round-end validation on one B200: GPU tests, smoke, the default bench (both arms).
mkdir -p gpurun_out
python -m pytest tests/ -m gpu -q --timeout 1200 -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1; tail -2 gpurun_out/final_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['roofline']['frac'], json.dumps(d['spmv']), json.dumps(d['solve']), json.dumps(d['tts']), json.dumps(d['e2e']), json.dumps(d['cpu_baseline']), d['clocks'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -c 300 gpurun_out/final_bench_ref.json
