# final one-GPU validation of round 2: all GPU tests, smoke, bench (own arm + reference arm), ncu launch list + full capture, sweeps
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/f_tests.log; echo "rc=$?" >> gpurun_out/f_tests.log
python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/f_smoke.log
python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q -s -k "full_size" 2>&1 | grep "^\[" > gpurun_out/f_parity_report.txt
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "rc=$?" >> gpurun_out/f_bench.err
python bench.py --impl reference > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "rc=$?" >> gpurun_out/f_bench_ref.err
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/f_launches.csv $NCU_CMD > gpurun_out/f_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"project_kernel|bin_walk_kernel|bin_scan_kernel|bucket_sort_kernel|lsd_sort_kernel|composite_fwd_kernel|composite_bwd_kernel|preprocess_bwd_kernel" -s 22 -c 11 -f -o gpurun_out/f_prof $NCU_CMD > gpurun_out/f_ncu_full.log 2>&1
python tools/sweep.py C1 C2 C2T C3 C4 C5 > gpurun_out/f_sweep.jsonl 2> gpurun_out/f_sweep.err
python tools/clip_bench.py > gpurun_out/f_clip.jsonl 2> gpurun_out/f_clip.err
python tools/adapter_bench.py > gpurun_out/f_adapter.jsonl 2> gpurun_out/f_adapter.err
tail -4 gpurun_out/f_tests.log; tail -2 gpurun_out/f_smoke.log; cut -c1-300 gpurun_out/f_bench.json; cut -c1-300 gpurun_out/f_bench_ref.json; cut -c1-260 gpurun_out/f_sweep.jsonl; cut -c1-300 gpurun_out/f_clip.jsonl; cut -c1-300 gpurun_out/f_adapter.jsonl; ls -la gpurun_out/f_prof.ncu-rep
