mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_loss_fusion.py tests/test_gpu_edge_cases.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/q_tests.log; echo "rc=$?" >> gpurun_out/q_tests.log
python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -x -k "C1 or C2T or render_depth" 2>&1 | tail -2 >> gpurun_out/q_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/q_bench_final.json 2> gpurun_out/q_bench_final.err
python -c "
import json,sys
d=json.load(open('gpurun_out/q_bench_final.json')); print(d['ms_per_step'], d['value'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['roofline']['frac'])
"
cat gpurun_out/q_tests.log
