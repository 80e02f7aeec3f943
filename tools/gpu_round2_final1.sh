mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/tests.log; echo "rc=$?" >> gpurun_out/tests.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"project_kernel|bin_walk_kernel|bin_scan_kernel|bucket_sort_kernel|lsd_sort_kernel|composite_fwd_kernel|composite_bwd_kernel|preprocess_bwd_kernel" -s 22 -c 11 -f -o gpurun_out/prof_final $NCU_CMD > gpurun_out/ncu_full.log 2>&1
python tools/sweep.py C1 C2 C2T C3 C4 > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err
python tools/clip_bench.py > gpurun_out/clip.jsonl 2> gpurun_out/clip.err
tail -4 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/sweep.jsonl | cut -c1-260; cat gpurun_out/clip.jsonl | cut -c1-300
