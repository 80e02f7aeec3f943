set -o pipefail
mkdir -p gpurun_out
python tools/sweep.py C5 > gpurun_out/sweep_c5_auto.jsonl 2> gpurun_out/sweep_c5_auto.err
python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_binsort.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"project_kernel|bin_walk_kernel|bin_scan_kernel|bucket_sort_kernel|lsd_sort_kernel|composite_fwd_kernel|composite_bwd_kernel|preprocess_bwd_kernel" -s 22 -c 11 -f -o gpurun_out/prof_r2l $NCU_CMD > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/sweep_c5_auto.jsonl; tail -3 gpurun_out/t_a.log; tail -3 gpurun_out/ncu_full.log
