# rematerialising split calls (C5 memory bound), clip frames against the oracle
mkdir -p gpurun_out
python -m pytest tests/test_gpu_clip.py tests/test_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/s_tests.log; echo "rc=$?" >> gpurun_out/s_tests.log
tail -5 gpurun_out/s_tests.log
timeout 600 python tools/sweep.py C5 > gpurun_out/s_sweep_c5.jsonl 2> gpurun_out/s_sweep_c5.err; cut -c1-600 gpurun_out/s_sweep_c5.jsonl; tail -3 gpurun_out/s_sweep_c5.err
B200S_SORT_MODE=binned timeout 600 python tools/sweep.py C5 C2T > gpurun_out/s_sweep_c5b.jsonl 2> gpurun_out/s_sweep_c5b.err; cut -c1-600 gpurun_out/s_sweep_c5b.jsonl; tail -3 gpurun_out/s_sweep_c5b.err
