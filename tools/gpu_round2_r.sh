# projection kernels without per-view barriers: parity + bench
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py tests/test_gpu_adapter.py tests/test_gpu_loss_fusion.py tests/test_gpu_binsort.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r_tests.log; echo "rc=$?" >> gpurun_out/r_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r_bench.json')); print(d['ms_per_step'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['clocks']['sm_mhz'])
"
python tools/sweep.py C1 C4 C2 > gpurun_out/r_sweep.jsonl 2> gpurun_out/r_sweep.err; cut -c1-250 gpurun_out/r_sweep.jsonl
tail -6 gpurun_out/r_tests.log
