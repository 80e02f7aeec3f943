set -o pipefail
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline --knob 3=2 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
python tools/sweep.py C5 > gpurun_out/sweep_c5_auto.jsonl 2> gpurun_out/sweep_c5_auto.err
python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_binsort.py tests/test_gpu_adapter.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"project_kernel|bin_walk_kernel|bin_scan_kernel|bucket_sort_kernel|lsd_sort_kernel|composite_fwd_kernel|composite_bwd_kernel|preprocess_bwd_kernel" -s 22 -c 11 -f -o gpurun_out/prof_r2l $NCU_CMD > gpurun_out/ncu_full.log 2>&1
python -c "
import json
for f in ('a','b'):
    d=json.load(open('gpurun_out/bench_%s.json'%f)); print(f, d['ms_per_step'], {k:v['ms'] for k,v in d['stages'].items()})
"
cat gpurun_out/sweep_c5_auto.jsonl; tail -3 gpurun_out/t_a.log; tail -3 gpurun_out/ncu_full.log
