python -m pytest tests/test_gpu_path.py -m gpu -q --timeout 900 -p no:cacheprovider -x 2>&1 | tail -3
for cfg in "3 40" "4 24" "5 16" "6 12"; do set -- $cfg; python bench.py --steps 5 --warmup 3 --p $1 --m $2 --no-cpu --no-solve --no-tts > gpurun_out/bench_p$1c.json 2> gpurun_out/bench_p$1c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_p$1c.json')); print('p$1', d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])"; done
