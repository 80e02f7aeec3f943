set -o pipefail
mkdir -p gpurun_out
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-gpu-baseline > gpurun_out/bench_n2_default.json 2> gpurun_out/bench_n2_default.err; echo "rc=$?" >> gpurun_out/bench_n2_default.err
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --pieces 8 --no-cpu --no-gpu-baseline > gpurun_out/bench_n2_p8.json 2> gpurun_out/bench_n2_p8.err
cut -c1-300 gpurun_out/bench_n2_default.json; tail -4 gpurun_out/bench_n2_default.err
