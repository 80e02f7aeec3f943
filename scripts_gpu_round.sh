#!/bin/bash
# One GPU round on the box.  Usage: ./scripts_gpu_round.sh [tests] [bench] [ncu-list] [ncu-full <kernel-regex>]
set -o pipefail
mkdir -p gpurun_out
while [ $# -gt 0 ]; do
  case "$1" in
    tests)
      python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/tests.log; echo "pytest rc=$?" >> gpurun_out/tests.log
      python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
      tail -4 gpurun_out/tests.log; tail -2 gpurun_out/smoke.log ;;
    bench)
      python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
      cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
    benchquick)
      python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
      cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
    baseline)
      python -m pytest tests/test_gpu_baseline.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/baseline_tests.log; echo "pytest rc=$?" >> gpurun_out/baseline_tests.log; tail -15 gpurun_out/baseline_tests.log ;;
    ncu-list)
      # same command line run plain first, directly before, no pipe
      NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
      $NCU_CMD > gpurun_out/ncu_plain.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 130 --csv --log-file gpurun_out/launches.csv $NCU_CMD > gpurun_out/ncu_launches.log 2>&1
      echo "ncu-list rc=$?" ;;
    ncu-full)
      shift; REGEX="$1"
      NCU_CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-gpu-baseline"
      $NCU_CMD > gpurun_out/ncu_plain.log 2>&1 &&
      ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s ${NCU_SKIP:-9} -c ${NCU_COUNT:-6} -f -o gpurun_out/prof $NCU_CMD > gpurun_out/ncu_full.log 2>&1
      echo "ncu-full rc=$?" ;;
  esac
  shift
done
