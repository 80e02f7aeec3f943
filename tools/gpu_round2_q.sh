# A/B of load placement in the compositing kernels
mkdir -p gpurun_out
B200S_KNOBS=0=3,1=5 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_loss_fusion.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/q_tests.log; echo "rc=$?" >> gpurun_out/q_tests.log

for kn in "" "--knob 0=3" "--knob 1=4" "--knob 1=5" "--knob 0=3 --knob 1=5"; do
  tag=$(echo "$kn" | tr -d ' -' | tr '=' '_'); tag=${tag:-base}
  python bench.py --steps 20 --warmup 3 --no-cpu --no-gpu-baseline $kn > gpurun_out/q_bench_$tag.json 2> gpurun_out/q_bench_$tag.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/q_bench_$tag.json')); print('$tag', d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['clocks']['sm_mhz'])
"
done
cat gpurun_out/q_tests.log
