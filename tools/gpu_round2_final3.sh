# re-validation after the one-hit-ahead compositing backward: all GPU tests, smoke, bench line, launch list, ncu capture of the kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/h_tests.log; echo "rc=$?" >> gpurun_out/h_tests.log
python __graft_entry__.py smoke > gpurun_out/h_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/h_smoke.log
python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s -k "full_size" 2>&1 | grep "^\[" > gpurun_out/h_parity_report.txt
python bench.py > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "rc=$?" >> gpurun_out/h_bench.err
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/h_launches.csv $NCU_CMD > gpurun_out/h_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"project_kernel|bin_walk_kernel|bin_scan_kernel|bucket_sort_kernel|lsd_sort_kernel|composite_fwd_kernel|composite_bwd_kernel|preprocess_bwd_kernel" -s 22 -c 11 -f -o gpurun_out/h_prof $NCU_CMD > gpurun_out/h_ncu_full.log 2>&1
python tools/sweep.py C1 C2T C4 > gpurun_out/h_sweep.jsonl 2> gpurun_out/h_sweep.err
tail -3 gpurun_out/h_tests.log; tail -2 gpurun_out/h_smoke.log; python -c "
import json
d=json.load(open('gpurun_out/h_bench.json')); print(d['ms_per_step'], d['value'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['e2e']['value'], d['roofline']['frac'])
for l in open('gpurun_out/h_sweep.jsonl'):
    x=json.loads(l); print(x['config'], x['ms_per_call'], x['stages_ms'])
"; cat gpurun_out/h_parity_report.txt
