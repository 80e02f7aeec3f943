# N=2: every reducer against single-GPU gradients with the new projection backward, bench lines per piece count
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/dist_check.py > gpurun_out/t_dist_check_n2.log 2>&1; echo "rc=$?" >> gpurun_out/t_dist_check_n2.log
grep -c OK gpurun_out/t_dist_check_n2.log; tail -3 gpurun_out/t_dist_check_n2.log
for pc in 4 8 2; do
  $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-gpu-baseline --pieces $pc > gpurun_out/t_bench_n2_p$pc.json 2> gpurun_out/t_bench_n2_p$pc.err
  python -c "
import json
d=json.load(open('gpurun_out/t_bench_n2_p$pc.json')); print('pieces $pc', d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('parallelism'))
"
done
