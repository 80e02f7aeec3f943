set -o pipefail
mkdir -p gpurun_out
python -m pytest tests/test_gpu_loss_fusion.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/t_a.log; echo "rc=$?" >> gpurun_out/t_a.log
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/dist_check_n2.log
for g in allreduce scatter; do
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --grads $g --no-cpu --no-gpu-baseline > gpurun_out/bench_n2_$g.json 2> gpurun_out/bench_n2_$g.err; echo "rc=$?" >> gpurun_out/bench_n2_$g.err
done
tail -5 gpurun_out/t_a.log; grep -c OK gpurun_out/dist_check_n2.log; tail -3 gpurun_out/dist_check_n2.log; for f in gpurun_out/bench_n2_allreduce.json gpurun_out/bench_n2_scatter.json; do echo $f; cut -c1-300 $f; done; tail -5 gpurun_out/bench_n2_scatter.err
