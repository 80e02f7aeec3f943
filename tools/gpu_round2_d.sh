set -o pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_binsort.py tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s -k "C2T or C3 or C5b" 2>&1 | tail -40 > gpurun_out/t_b.log; echo "rc=$?" >> gpurun_out/t_b.log
python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/bench_binned.json 2> gpurun_out/bench_binned.err; echo "rc=$?" >> gpurun_out/bench_binned.err
python tools/sweep.py C1 C2 C3 C4 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"sort_kernel|bin_walk_kernel|project_kernel" -s 24 -c 7 -f -o gpurun_out/prof_r2d $NCU_CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/t_a.log gpurun_out/t_b.log; cat gpurun_out/bench_binned.json; cat gpurun_out/sweep.jsonl
