# last validation of the round: all GPU tests, smoke, bench line, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/i_tests.log; echo "rc=$?" >> gpurun_out/i_tests.log
python __graft_entry__.py smoke > gpurun_out/i_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/i_smoke.log
python bench.py > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "rc=$?" >> gpurun_out/i_bench.err
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/i_launches.csv $NCU_CMD > gpurun_out/i_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"preprocess_bwd_kernel" -s 2 -c 1 -f -o gpurun_out/i_prof_prebwd $NCU_CMD > gpurun_out/i_ncu_full.log 2>&1
tail -3 gpurun_out/i_tests.log; tail -2 gpurun_out/i_smoke.log; python -c "
import json
d=json.load(open('gpurun_out/i_bench.json')); print(d['ms_per_step'], d['value'], {k:round(v['ms'],3) for k,v in d['stages'].items()}, d['e2e']['value'], d['roofline']['frac'])
"
