# A/B: register budget of the projection forward; parity of the new projection backward budget; small scenes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_adapter.py tests/test_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/q_tests.log; echo "rc=$?" >> gpurun_out/q_tests.log
for kn in "" "--knob 3=3" ""; do
  tag=$(echo "$kn" | tr -d ' -' | tr '=' '_'); tag=${tag:-base}
  python bench.py --steps 20 --warmup 3 --no-cpu --no-gpu-baseline $kn > gpurun_out/q_bench_$tag.json 2> gpurun_out/q_bench_$tag.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/q_bench_$tag.json')); print('$tag', d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['clocks']['sm_mhz'])
"
done
python tools/sweep.py C1 C4 > gpurun_out/q_sweep.jsonl 2> gpurun_out/q_sweep.err; python -c "
import json
for l in open('gpurun_out/q_sweep.jsonl'):
    x=json.loads(l); print(x['config'], x['ms_per_call'], x['stages_ms'])
"
python tools/adapter_bench.py 2>/dev/null | cut -c1-200
cat gpurun_out/q_tests.log
