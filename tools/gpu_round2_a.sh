set -o pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_binsort.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s 2>&1 | tail -80 > gpurun_out/t_b.log; echo "rc=$?" >> gpurun_out/t_b.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize_parity.py --deselect tests/test_gpu_binsort.py --deselect tests/test_gpu_parity.py 2>&1 | tail -40 > gpurun_out/t_c.log; echo "rc=$?" >> gpurun_out/t_c.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_binned.json 2> gpurun_out/bench_binned.err; echo "rc=$?" >> gpurun_out/bench_binned.err
B200S_SORT_MODE=global python bench.py --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/bench_global.json 2> gpurun_out/bench_global.err; echo "rc=$?" >> gpurun_out/bench_global.err
tail -3 gpurun_out/t_a.log gpurun_out/t_b.log gpurun_out/t_c.log; cat gpurun_out/bench_binned.json
