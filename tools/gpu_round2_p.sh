mkdir -p gpurun_out
python -m pytest tests/test_gpu_fullsize_parity.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -q -s 2>&1 | grep -v "^   d" | tail -30 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -22 gpurun_out/t_a.log
python -c "
import json
d=json.load(open('gpurun_out/bench_a.json')); print(d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()}, d['clocks'])
"
