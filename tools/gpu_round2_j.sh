set -o pipefail
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/tests.log; echo "rc=$?" >> gpurun_out/tests.log
python tools/adapter_bench.py C1 > gpurun_out/adapter_bench.jsonl 2> gpurun_out/adapter_bench.err
python tools/adapter_bench.py C2T >> gpurun_out/adapter_bench.jsonl 2>> gpurun_out/adapter_bench.err
python tools/sweep.py C5 > gpurun_out/sweep_c5.jsonl 2> gpurun_out/sweep_c5.err
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -6 gpurun_out/tests.log; cat gpurun_out/adapter_bench.jsonl; tail -2 gpurun_out/adapter_bench.err; cat gpurun_out/sweep_c5.jsonl; tail -2 gpurun_out/sweep_c5.err; tail -2 gpurun_out/smoke.log
