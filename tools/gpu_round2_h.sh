set -o pipefail
mkdir -p gpurun_out
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_n8.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_n8.log
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/bench_n8_default.json 2> gpurun_out/bench_n8_default.err; echo "rc=$?" >> gpurun_out/bench_n8_default.err
torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --grads allreduce --e2e-replicated --no-cpu --no-gpu-baseline > gpurun_out/bench_n8_allreduce.json 2> gpurun_out/bench_n8_allreduce.err
grep -c OK gpurun_out/dist_check_n8.log; tail -2 gpurun_out/dist_check_n8.log; cut -c1-250 gpurun_out/bench_n8_default.json; tail -3 gpurun_out/bench_n8_default.err
