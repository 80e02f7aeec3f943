# A/B of the half-warp compositing variants: parity with the variants on, then bench lines per variant
mkdir -p gpurun_out
B200S_KNOBS=2=1,3=2 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py tests/test_gpu_loss_fusion.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/q_tests.log; echo "rc=$?" >> gpurun_out/q_tests.log
B200S_KNOBS=2=1,3=1 python -m pytest "tests/test_gpu_fullsize_parity.py" -m gpu -q -x -s -k "C2T or C1" 2>&1 | grep -v "^   d" | tail -15 > gpurun_out/q_tests_full.log; echo "rc=$?" >> gpurun_out/q_tests_full.log
for kn in "" "--knob 2=1" "--knob 3=1" "--knob 3=2" "--knob 2=1 --knob 3=2"; do
  tag=$(echo "$kn" | tr -d ' -' | tr '=' '_'); tag=${tag:-base}
  python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline $kn > gpurun_out/q_bench_$tag.json 2> gpurun_out/q_bench_$tag.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/q_bench_$tag.json')); print('$tag', d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['clocks']['sm_mhz'])
"
done
tail -6 gpurun_out/q_tests.log; tail -8 gpurun_out/q_tests_full.log
