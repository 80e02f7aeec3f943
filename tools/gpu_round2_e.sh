set -o pipefail
mkdir -p gpurun_out
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_n2.log
for mode in views ranges; do
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --shard $mode --no-cpu --no-gpu-baseline > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "rc=$?" >> gpurun_out/bench_n2_$mode.err
done
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --shard ranges --nccl-scatter --no-cpu --no-gpu-baseline > gpurun_out/bench_n2_ranges_nccl.json 2> gpurun_out/bench_n2_ranges_nccl.err
tail -12 gpurun_out/dist_check_n2.log; for f in gpurun_out/bench_n2_*.json; do echo $f; cut -c1-400 $f; done; tail -3 gpurun_out/bench_n2_ranges.err
