mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/tests.log; echo "rc=$?" >> gpurun_out/tests.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -5 gpurun_out/tests.log
python -c "
import json
d=json.load(open('gpurun_out/bench_a.json')); print(d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})
"
