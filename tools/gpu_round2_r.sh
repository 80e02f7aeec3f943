# projection kernel variants: parity + bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py tests/test_gpu_adapter.py tests/test_gpu_loss_fusion.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r_tests.log; echo "rc=$?" >> gpurun_out/r_tests.log
for i in 1 2; do
python bench.py --steps 20 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r_bench.json')); print(d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['clocks']['sm_mhz'])
"
done
python tools/sweep.py C1 C4 C2 C3 > gpurun_out/r_sweep.jsonl 2> gpurun_out/r_sweep.err; python -c "
import json
for l in open('gpurun_out/r_sweep.jsonl'):
    d=json.loads(l); print(d['config'], d['ms_per_call'], d['stages_ms'])
"
tail -4 gpurun_out/r_tests.log
