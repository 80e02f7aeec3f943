set -o pipefail
mkdir -p gpurun_out
B200S_SORT_MODE=global python tools/sweep.py C5 > gpurun_out/sweep_c5_global.jsonl 2> gpurun_out/sweep_c5_global.err
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_full.json 2> gpurun_out/bench_n1_full.err; echo "rc=$?" >> gpurun_out/bench_n1_full.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/bench_ref.err
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 150 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"b200s" -s 20 -c 16 -f -o gpurun_out/prof_r2k $NCU_CMD > gpurun_out/ncu_full.log 2>&1
python tools/clip_bench.py > gpurun_out/clip.jsonl 2> gpurun_out/clip.err
cat gpurun_out/sweep_c5_global.jsonl; cut -c1-300 gpurun_out/bench_n1_full.json; cat gpurun_out/bench_ref.json | cut -c1-300; cat gpurun_out/clip.jsonl; tail -3 gpurun_out/ncu_full.log
