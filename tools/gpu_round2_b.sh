set -o pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/t_b.log; echo "rc=$?" >> gpurun_out/t_b.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_binned.json 2> gpurun_out/bench_binned.err; echo "rc=$?" >> gpurun_out/bench_binned.err
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bin_sort_kernel|bin_walk_kernel|composite_bwd" -s 30 -c 7 -f -o gpurun_out/prof_r2b $NCU_CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/t_a.log gpurun_out/t_b.log; cat gpurun_out/bench_binned.json
