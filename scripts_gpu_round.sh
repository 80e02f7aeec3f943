#!/bin/bash
# one GPU round: parity tests, smoke, bench, then (only if the plain run exited 0) ncu passes
set -o pipefail
mkdir -p gpurun_out
if [ "$1" != "noTests" ]; then
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/tests.log; echo "pytest rc=$?" >> gpurun_out/tests.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
fi
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -3 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
# ncu: same command line run plain first, directly before, no pipe
NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu"
$NCU_CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
$NCU_CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'onesweep_pass|composite_bwd|composite_fwd|preprocess_bin|preprocess_bwd|digit_histogram' -s 55 -c 14 -o gpurun_out/prof_r1a $NCU_CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu done rc=$?"; ls -la gpurun_out | tail -12
